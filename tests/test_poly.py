"""Caller-side polynomial multiplication through the tree: ENTER, element-wise product, EXIT.
The reference has no such function; it is what its users write around `enter` / `exit` (README.md:60-63) and
what benches/comparison.rs:37-43 times as evaluate / interpolate.  Checked against schoolbook multiplication
in Python big integers (oracle/pyref.py)."""
import numpy as np
import pytest


def _mont_mul_np(O, a, b):
    """element-wise Montgomery product with the oracle's field routine (ark-ff `*`)"""
    import ctypes
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(a)
    L = O.lib()
    for i in range(len(a)):
        L.orc_fe_mul(a[i].ctypes.data_as(ctypes.c_void_p), b[i].ctypes.data_as(ctypes.c_void_p),
                     out[i].ctypes.data_as(ctypes.c_void_p))
    return out


def test_enter_multiply_exit_is_schoolbook_on_the_oracle(oracle_mod):
    """the identity itself, on the CPU oracle: exit(enter(a) * enter(b)) == a * b for deg a + deg b < n"""
    from oracle import pyref
    O = oracle_mod
    n = 64
    tree = O.OracleTree.build(n)
    rng = np.random.default_rng(7)
    for la, lb in ((1, 1), (5, 9), (32, 33), (40, 25)):
        a = [int(v) % pyref.P for v in rng.integers(0, 2**63, la)]
        b = [int(v) * 2**190 % pyref.P for v in rng.integers(0, 2**63, lb)]
        pa, pb = np.zeros((n, 4), dtype=np.uint64), np.zeros((n, 4), dtype=np.uint64)
        pa[:la], pb[:lb] = O.to_mont(a), O.to_mont(b)
        got = tree.exit(_mont_mul_np(O, tree.enter(pa), tree.enter(pb)))
        want = pyref.poly_mul(a, b)
        assert O.from_mont(got[: la + lb - 1]) == want
        assert not got[la + lb - 1:].any()


@pytest.mark.gpu
def test_pointwise_mul_matches_field_product(oracle_mod):
    import torch
    import ecfft_b200
    O = oracle_mod
    tree = ecfft_b200.build_fftree(16, parts=ecfft_b200.PARTS_ENTER_ONLY)
    for n in (1, 3, 1000, 4097):
        a, b = O.random_elements(n, seed=n), O.random_elements(n, seed=n + 1)
        want = _mont_mul_np(O, a, b)
        assert (tree.pointwise_mul(a, b) == want).all()
        da, db = (torch.from_numpy(v.view(np.int64)).cuda() for v in (a, b))
        assert (tree.pointwise_mul(da, db).cpu().numpy().view(np.uint64) == want).all()
    # edge values: 0, 1, p - 1 (Montgomery forms)
    P = O.P
    vals = O.to_mont([0, 1, P - 1, 2, P - 2])
    other = O.to_mont([P - 1, P - 1, P - 1, (P + 1) // 2, 5])
    assert O.from_mont(tree.pointwise_mul(vals, other)) == [0, P - 1, 1, 1, (P - 10) % P]
    with pytest.raises(ecfft_b200.EcfftError):
        tree.pointwise_mul(da, db[:5])


@pytest.mark.gpu
@pytest.mark.parametrize("la,lb", [(1, 1), (7, 2), (64, 64), (300, 213), (1024, 1025)])
def test_poly_mul_is_schoolbook(oracle_mod, la, lb):
    import ecfft_b200
    from oracle import pyref
    O = oracle_mod
    tree = ecfft_b200.build_fftree(4096)
    rng = np.random.default_rng(la * 1000 + lb)
    a = [int.from_bytes(rng.bytes(32), "little") % pyref.P for _ in range(la)]
    b = [int.from_bytes(rng.bytes(32), "little") % pyref.P for _ in range(lb)]
    got = ecfft_b200.poly_mul(tree, O.to_mont(a), O.to_mont(b))
    assert len(got) == la + lb - 1
    assert O.from_mont(got) == pyref.poly_mul(a, b)


@pytest.mark.gpu
def test_poly_mul_large_on_device_by_evaluation(oracle_mod):
    """2^19 x 2^19 coefficients on device tensors: the product evaluated at random points equals the product of
    the factors' evaluations (Horner in Python big integers)"""
    import torch
    import ecfft_b200
    from oracle import pyref
    O = oracle_mod
    tree = ecfft_b200.build_fftree(1 << 20)
    la = lb = 1 << 19
    a, b = O.random_elements(la, seed=91), O.random_elements(lb, seed=92)
    da, db = (torch.from_numpy(v.view(np.int64)).cuda() for v in (a, b))
    prod = ecfft_b200.poly_mul(tree, da, db).cpu().numpy().view(np.uint64)
    assert len(prod) == la + lb - 1
    ai, bi, ci = O.from_mont(a), O.from_mont(b), O.from_mont(prod)
    for x in (2, 3**100 % pyref.P, pyref.P - 5):
        assert pyref.horner(ci, x) == pyref.horner(ai, x) * pyref.horner(bi, x) % pyref.P
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.poly_mul(tree, da, torch.cat([db, db, db]))
    assert e.value.code == ecfft_b200._lib.ERR_TREE_TOO_SMALL
