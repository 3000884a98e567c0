"""Consumes golden vectors produced by the REAL reference crate (tools/rust_golden, arkworks 0.4) when they
are present: tests/golden/arkworks_n64.txt.  No Rust toolchain exists in this repository's image, so the file
cannot be generated here; a maintainer with cargo runs tools/rust_golden once (see its Cargo.toml) and these
tests then pin the oracle (CPU) and the CUDA path (GPU) to arkworks bit for bit — the serialised bytes
included.  Until then they are reported as skipped.  The parsing / comparison code itself is exercised by
`test_consumer_against_oracle_written_file`, which writes a file of the same format from the oracle."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.environ.get("ECFFT_ARKWORKS_GOLDEN") or os.path.join(ROOT, "tests", "golden", "arkworks_n64.txt")


def parse(path):
    out = {}
    with open(path) as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            if parts[0] == "n":
                out["n"] = int(parts[1])
            elif parts[0] == "num":
                out[parts[1]] = int(parts[2])
            elif parts[0] == "bytes":
                out[parts[1]] = bytes.fromhex(parts[2]) if len(parts) > 2 else b""
            elif parts[0] == "vec":
                raw = bytes.fromhex(parts[2]) if len(parts) > 2 else b""
                out[parts[1]] = np.frombuffer(raw, dtype="<u8").reshape(-1, 4).astype(np.uint64)
    return out


def write_from_oracle(O, path, n=64):
    """the same file as tools/rust_golden/src/main.rs writes, produced by the oracle (inputs from the seeded
    generator instead of Rust's StdRng: the file carries its own inputs)"""
    tree = O.OracleTree.build(n)
    lines = [f"n {n}"]

    def vec(name, a):
        lines.append(f"vec {name} {np.ascontiguousarray(a, dtype='<u8').tobytes().hex()}")

    lines.append(f"bytes tree_compressed {tree.serialize(True).hex()}")
    lines.append(f"bytes tree_uncompressed {tree.serialize(False).hex()}")
    vec("leaves", tree.leaves())
    vec("xnn_s", tree.table("xnn_s"))
    vec("z0z0_rem_xnn_s", tree.table("z0z0_rem_xnn_s"))
    coeffs, arb, half = O.random_elements(n, 1), O.random_elements(n, 2), O.random_elements(n // 2, 3)
    a, c = O.random_elements(n, 4), O.random_elements(n, 5)
    evals = tree.enter(coeffs)
    vec("enter.in", coeffs); vec("enter.out", evals)
    vec("exit.in", arb); vec("exit.out", tree.exit(arb))
    vec("extend.in", half)
    vec("extend_s1.out", tree.extend(half, 1)); vec("extend_s0.out", tree.extend(half, 0))
    vec("mextend_s1.out", tree.mextend(half, 1)); vec("mextend_s0.out", tree.mextend(half, 0))
    vec("redc.in", arb)
    vec("redc_z0.out", tree.redc_z0(arb, tree.table("xnn_s"))); vec("redc_z1.out", tree.redc_z1(arb, tree.table("xnn_s")))
    vec("mod.out", tree.modular_reduce(arb, tree.table("xnn_s"), tree.table("z0z0_rem_xnn_s")))
    vec("redc_a.in", a); vec("mod_c.in", c)
    vec("redc_z0_a.out", tree.redc_z0(arb, a)); vec("mod_ac.out", tree.modular_reduce(arb, a, c))
    vec("vanish.in", half); vec("vanish.out", tree.vanish(half))
    low = coeffs.copy()
    low[41:] = 0
    le = tree.enter(low)
    vec("degree.in", le)
    lines.append(f"num degree.out {tree.degree(le)}")
    lines.append(f"num degree_full.out {tree.degree(evals)}")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def check(g, tree, serialize):
    """every vector of the file against `tree` (an oracle tree or the CUDA FFTree: same method names)"""
    def eq(name, got):
        want = g[name]
        got = np.asarray(got, dtype=np.uint64).reshape(-1, 4)
        assert got.shape == want.shape and (got == want).all(), f"{name} differs from arkworks"

    eq("enter.out", tree.enter(g["enter.in"]))
    eq("exit.out", tree.exit(g["exit.in"]))
    eq("extend_s1.out", tree.extend(g["extend.in"], 1))
    eq("extend_s0.out", tree.extend(g["extend.in"], 0))
    eq("mextend_s1.out", tree.mextend(g["extend.in"], 1))
    eq("mextend_s0.out", tree.mextend(g["extend.in"], 0))
    eq("redc_z0.out", tree.redc_z0(g["redc.in"], g["xnn_s"]))
    eq("redc_z1.out", tree.redc_z1(g["redc.in"], g["xnn_s"]))
    eq("mod.out", tree.modular_reduce(g["redc.in"], g["xnn_s"], g["z0z0_rem_xnn_s"]))
    eq("redc_z0_a.out", tree.redc_z0(g["redc.in"], g["redc_a.in"]))
    eq("mod_ac.out", tree.modular_reduce(g["redc.in"], g["redc_a.in"], g["mod_c.in"]))
    eq("vanish.out", tree.vanish(g["vanish.in"]))
    assert tree.degree(g["degree.in"]) == g["degree.out"]
    assert tree.degree(g["enter.out"]) == g["degree_full.out"]
    assert serialize(True) == g["tree_compressed"], "compressed FFTree bytes differ from ark-serialize"
    assert serialize(False) == g["tree_uncompressed"], "uncompressed FFTree bytes differ from ark-serialize"


def test_consumer_against_oracle_written_file(oracle_mod, tmp_path):
    """not a parity check: keeps the reader and the comparison alive while no arkworks file exists"""
    path = str(tmp_path / "oracle_n64.txt")
    write_from_oracle(oracle_mod, path)
    g = parse(path)
    tree = oracle_mod.OracleTree.build(g["n"])
    assert (tree.leaves() == g["leaves"]).all()
    check(g, tree, tree.serialize)
    g["enter.out"] = g["enter.out"].copy()
    g["enter.out"][5, 0] ^= 1
    with pytest.raises(AssertionError):
        check(g, tree, tree.serialize)


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="tests/golden/arkworks_n64.txt not generated (needs cargo: tools/rust_golden)")
def test_oracle_matches_arkworks_vectors(oracle_mod):
    g = parse(GOLDEN)
    tree = oracle_mod.OracleTree.build(g["n"])
    assert (tree.leaves() == g["leaves"]).all()
    check(g, tree, tree.serialize)
    # and the real crate's bytes load into the oracle and behave
    for compressed in (True, False):
        t2 = oracle_mod.OracleTree.deserialize(g["tree_compressed" if compressed else "tree_uncompressed"], compressed)
        assert (t2.enter(g["enter.in"]) == g["enter.out"]).all()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="tests/golden/arkworks_n64.txt not generated (needs cargo: tools/rust_golden)")
def test_cuda_path_matches_arkworks_vectors():
    import ecfft_b200
    g = parse(GOLDEN)
    tree = ecfft_b200.build_fftree(g["n"])
    assert (tree.eval_domain() == g["leaves"]).all()
    check(g, tree, tree.serialize)
    for compressed in (True, False):
        t2 = ecfft_b200.FFTree.deserialize(g["tree_compressed" if compressed else "tree_uncompressed"], compressed)
        assert (t2.enter(g["enter.in"]) == g["enter.out"]).all()


@pytest.mark.gpu
def test_cuda_path_against_oracle_written_file(oracle_mod, tmp_path):
    """the same consumer, CUDA path against the oracle-written file (always runs)"""
    import ecfft_b200
    path = str(tmp_path / "oracle_n64.txt")
    write_from_oracle(oracle_mod, path)
    g = parse(path)
    tree = ecfft_b200.build_fftree(g["n"])
    check(g, tree, tree.serialize)
