"""Known-answer vectors (tests/golden/secp256k1_n16.json, produced by tests/golden/make_golden.py from
Python big integers only) against the CPU oracle and — on a GPU box — the CUDA path through the C ABI."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "secp256k1_n16.json")) as f:
        doc = json.load(f)
    return {k: ([int(v, 16) for v in doc[k]] if isinstance(doc[k], list) else doc[k]) for k in doc}


def _check(tree, g):
    n = g["n"]
    assert O.from_mont(tree.eval_domain() if hasattr(tree, "eval_domain") else tree.table("f")[n:]) == g["leaves"]
    assert O.from_mont(tree.enter(O.to_mont(g["coeffs"]))) == g["enter"]
    assert O.from_mont(tree.exit(O.to_mont(g["enter"]))) == g["coeffs"]
    assert O.from_mont(tree.extend(O.to_mont(g["half_on_s0"]), 1)) == g["half_on_s1"]
    assert O.from_mont(tree.extend(O.to_mont(g["half_on_s1"]), 0)) == g["half_on_s0"]
    assert tree.degree(O.to_mont(g["enter"])) == max(i for i, c in enumerate(g["coeffs"]) if c)
    # vanish(domain) evaluated on the 2k-leaf tree (k = 8 -> the 16-leaf tree), src/fftree.rs:291-316
    assert O.from_mont(tree.vanish(O.to_mont(g["vanish_domain"]))) == g["vanish"]


def test_oracle_reproduces_the_golden_vectors(golden):
    _check(O.OracleTree.build(golden["n"]), golden)


@pytest.mark.gpu
def test_cuda_path_reproduces_the_golden_vectors(golden):
    import ecfft_b200
    _check(ecfft_b200.build_fftree(golden["n"]), golden)
