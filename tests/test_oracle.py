"""Pins the CPU oracle (oracle/ecfft_oracle.c) — the checker everything else is compared with.

The reference holds no golden vectors for this path; what pins it are its own property tests
(src/lib.rs:108-186) re-run here with an INDEPENDENT Python big-integer evaluator
(oracle/pyref.py: Horner, schoolbook polynomial arithmetic, affine curve arithmetic), the curve
constants of src/lib.rs:45-59, and definition-level checks for the operations the reference
leaves untested on secp256k1 (SURVEY.md 8c properties 3-8).
"""
import random

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref

P = pyref.P


@pytest.fixture(scope="module")
def tree64():
    return O.OracleTree.build(64)


@pytest.fixture(scope="module")
def tree256():
    return O.OracleTree.build(256)


def rand_coeffs(n, seed):
    rnd = random.Random(seed)
    return [rnd.randrange(P) for _ in range(n)]


def test_curve_constants():
    """src/lib.rs:45-59: both points lie on y^2 = x^3 + a x^2 + bb x, the generator has order 2^36"""
    on_curve = lambda x, y: (y * y - (x ** 3 + pyref.A * x * x + pyref.BB * x)) % P == 0
    assert on_curve(*pyref.OFFSET) and on_curve(*pyref.GEN)
    g = pyref.GEN
    for _ in range(35):
        g = pyref.ec_add(g, g)
    assert g is not None and g[1] == 0          # order-2 point (0,0) shared by all Good Curves (src/ec.rs:26)
    assert pyref.ec_add(g, g) is None
    # the hex literals in the C oracle / CUDA builder are these decimals
    assert "%064x" % pyref.A == "44eae664a07c69e1c7d7821cacf2a3ccca446568bd32b2a48166309c5c4297e5"
    assert "%064x" % pyref.BB == "649cd342698de65c9bc86f1ece3beb99197d6715a53bdb5609cc937a16154ca8"
    assert "%064x" % pyref.OFFSET[0] == "e9850041b13ea03fadc4bee2afd2959604bf64c290bf3fc15165f15163fd5431"
    assert "%064x" % pyref.OFFSET[1] == "110b996c1374482d0a6b9055a21dc8af9a098495b902b3663322f53ee416d65f"
    assert "%064x" % pyref.GEN[0] == "5b4b3e43cd5d95fba244389bb8655539cf8d527f331697e2e93ea60ef50ad5c4"
    assert "%064x" % pyref.GEN[1] == "a30fcedca51e68850478e0905816b86d88b79d7b549f4a340016e31de71ded06"


def test_field_arithmetic_vs_python():
    rnd = random.Random(7)
    vals = [0, 1, P - 1, P - 2, 2 ** 255, 977] + [rnd.randrange(P) for _ in range(200)]
    a = O.to_mont(vals)
    b = O.to_mont(list(reversed(vals)))
    out = np.empty_like(a)
    L = O.lib()
    for i in range(len(vals)):
        L.orc_fe_mul(a[i].ctypes.data, b[i].ctypes.data, out[i].ctypes.data)
    assert O.from_mont(out) == [x * y % P for x, y in zip(vals, reversed(vals))]
    inv = a.copy()
    L.orc_batch_inversion(inv.ctypes.data, len(inv))
    assert O.from_mont(inv) == [pow(x, -1, P) if x else 0 for x in vals]  # zeros untouched


@pytest.mark.parametrize("n", [1, 2, 4, 64, 256])
def test_leaves_are_coset_x_coordinates(n):
    """src/lib.rs:66-78 against an independent affine-arithmetic derivation"""
    t = O.OracleTree.build(n)
    assert O.from_mont(t.leaves()) == pyref.leaves(n)


def test_rational_maps_are_two_to_one(tree64):
    """debug_assert at src/fftree.rs:65 + derive_subtree's even-index layers (:465-482)"""
    f = O.from_mont(tree64.table("f"))
    n = 64
    a, bb = pyref.A, pyref.BB
    size = n
    while size > 1:
        b = pow(bb, (P + 1) // 4, P)
        layer, nxt = f[size:2 * size], f[size // 2:size]
        for i in range(size // 2):
            for x in (layer[i], layer[i + size // 2]):
                assert (x * x - 2 * b * x + b * b) * pow(x, -1, P) % P == nxt[i]
        a, bb = (a + 6 * b) % P, (4 * a * b + 8 * b * b) % P
        size //= 2
    sub = tree64.subtree_with_size(32)
    fs = O.from_mont(sub.table("f"))
    assert fs[32:] == f[64::2] and fs[16:32] == f[32:64:2]


def test_evaluates_polynomial(tree64):
    """src/lib.rs:108-120"""
    c = rand_coeffs(64, 1)
    xs = O.from_mont(tree64.leaves())
    assert O.from_mont(tree64.enter(O.to_mont(c))) == [pyref.horner(c, x) for x in xs]


def test_enter_on_subtree_sizes(tree64):
    for n in (1, 2, 4, 8, 16, 32):
        c = rand_coeffs(n, 10 + n)
        xs = O.from_mont(tree64.subtree_with_size(n).leaves())
        assert O.from_mont(tree64.enter(O.to_mont(c))) == [pyref.horner(c, x) for x in xs]


@pytest.mark.parametrize("target", [0, 1])
def test_extends_evaluations(tree64, target):
    """src/lib.rs:122-152 (both directions)"""
    c = rand_coeffs(32, 2)
    xs = O.from_mont(tree64.leaves())
    src, dst = (xs[1::2], xs[0::2]) if target == 0 else (xs[0::2], xs[1::2])
    got = tree64.extend(O.to_mont([pyref.horner(c, x) for x in src]), target)
    assert O.from_mont(got) == [pyref.horner(c, x) for x in dst]


@pytest.mark.parametrize("compressed", [True, False])
def test_deserialized_tree_works(tree64, compressed):
    """src/lib.rs:154-186"""
    blob = tree64.serialize(compressed)
    again = O.OracleTree.deserialize(blob, compressed)
    c = rand_coeffs(64, 3)
    xs = O.from_mont(tree64.leaves())
    assert O.from_mont(again.enter(O.to_mont(c))) == [pyref.horner(c, x) for x in xs]
    assert again.serialize(compressed) == blob
    for name in ("xnn_s_inv", "z0_inv_s1", "z1_inv_s0"):
        assert (again.table(name) == tree64.table(name)).all()


def test_serialized_layout_structure(tree64):
    """ark-serialize conventions restated in SURVEY.md section 5 (no golden bytes exist upstream)"""
    blob = tree64.serialize(False)
    n = 64
    u64 = lambda off: int.from_bytes(blob[off:off + 8], "little")
    assert u64(0) == 2 * n
    leaves = O.from_mont(tree64.leaves())
    off_leaf0 = 8 + 32 * n
    assert int.from_bytes(blob[off_leaf0:off_leaf0 + 32], "little") == leaves[0]  # canonical, not Montgomery
    off = 8 + 32 * 2 * n
    assert u64(off) == n                      # recombine_matrices: n entries of 4 Fp
    off += 8 + 128 * n
    assert u64(off) == n                      # decompose_matrices
    off += 8 + 128 * n
    assert u64(off) == 6                      # rational_maps: log2 n
    assert u64(off + 8) == 3                  # numerator b^2 - 2b x + x^2
    assert u64(off + 8 + 8 + 96) == 2         # denominator x
    sizes = {c: len(tree64.serialize(c)) for c in (True, False)}
    # per level: compressed drops xnn_s_inv (N), z0_inv_s1, z1_inv_s0 (N/2 each) and their 3 length words
    expect = sum((2 * N if N > 1 else 1) * 32 + 24 for N in (64, 32, 16, 8, 4, 2, 1))
    assert sizes[False] - sizes[True] == expect


def test_exit_inverts_enter(tree64):
    """pattern of src/lib.rs:253-264 (m31) applied to secp256k1"""
    for n in (1, 2, 8, 64):
        c = rand_coeffs(n, 20 + n)
        assert O.from_mont(tree64.exit(tree64.enter(O.to_mont(c)))) == c


def test_vanishing_tables(tree64):
    """z0_s1[i] = prod_{s in S0}(S1[i]-s), z1_s0[i] = prod_{s in S1}(S0[i]-s)  (src/fftree.rs:381-410)"""
    for n in (2, 4, 16, 64):
        st = tree64.subtree_with_size(n)
        xs = O.from_mont(st.leaves())
        s0, s1 = xs[0::2], xs[1::2]
        z0 = [pyref.horner(pyref.poly_from_roots(s0), x) for x in s1]
        z1 = [pyref.horner(pyref.poly_from_roots(s1), x) for x in s0]
        assert O.from_mont(st.table("z0_s1")) == z0
        assert O.from_mont(st.table("z1_s0")) == z1
        assert O.from_mont(st.table("z0_inv_s1")) == [pow(v, -1, P) for v in z0]


def test_z0z0_tables(tree64):
    """<Z_0^2 mod X^(n/2) on S> and <Z_1^2 mod X^(n/2) on S> (src/fftree.rs:412-460)"""
    for n in (2, 4, 16, 64):
        st = tree64.subtree_with_size(n)
        xs = O.from_mont(st.leaves())
        for name, roots in (("z0z0_rem_xnn_s", xs[0::2]), ("z1z1_rem_xnn_s", xs[1::2])):
            z = pyref.poly_from_roots(roots)
            rem = pyref.poly_mul(z, z)[: n // 2]
            assert O.from_mont(st.table(name)) == [pyref.horner(rem, x) for x in xs]


def test_vanish(tree64):
    """vanish(dom)[i] = prod_j (dom_j - leaf_i) on the 2k-leaf tree (src/fftree.rs:291-316).
    The base case is `alpha - leaf` (:297), so for k = 1 the result is MINUS the monic vanishing
    polynomial the doc comment promises; for even k the two agree."""
    for k in (1, 2, 8, 32):
        dom = rand_coeffs(k, 30 + k)
        xs = O.from_mont(tree64.subtree_with_size(2 * k).leaves())
        z = pyref.poly_from_roots(dom)
        sign = -1 if k == 1 else 1
        assert O.from_mont(tree64.vanish(O.to_mont(dom))) == [sign * pyref.horner(z, x) % P for x in xs]


def test_mextend(tree64):
    """MEXTEND extends MONIC polynomials of degree exactly h (src/fftree.rs:128-141)"""
    for h in (1, 2, 8, 32):
        c = rand_coeffs(h, 40 + h) + [1]
        xs = O.from_mont(tree64.subtree_with_size(2 * h).leaves())
        s0, s1 = xs[0::2], xs[1::2]
        got = tree64.mextend(O.to_mont([pyref.horner(c, x) for x in s0]), 1)
        assert O.from_mont(got) == [pyref.horner(c, x) for x in s1]
        got = tree64.mextend(O.to_mont([pyref.horner(c, x) for x in s1]), 0)
        assert O.from_mont(got) == [pyref.horner(c, x) for x in s0]


def test_degree(tree64):
    """src/lib.rs:266-278 pattern"""
    n = 64
    for d in (0, 1, n // 2 - 1, n // 2, n - 3, n - 1):
        c = rand_coeffs(d, 50 + d) + [1 + d] + [0] * (n - d - 1)
        assert tree64.degree(tree64.enter(O.to_mont(c))) == d


def test_redc_definition(tree64):
    """h = redc_z0(P, A): deg h < n/2 and h * Z_0 = P (mod A) for deg A = n/2, deg P < n.

    redc_z1 is NOT checked against its doc comment: as written (src/fftree.rs:232-259 with
    moiety = S1) it still divides the S_0 samples by a's S_0 samples and then extends them as
    if they lived on S_1, so it does not compute P * Z_1^-1 mod A.  Nothing in the reference
    tests or calls redc_z1; the oracle (and the CUDA path) restate the code literally so that
    results stay identical to the reference's, and only determinism is asserted here."""
    n = 32
    st = tree64.subtree_with_size(n)
    xs = O.from_mont(st.leaves())
    pc = rand_coeffs(n, 60)
    ac = rand_coeffs(n // 2, 61) + [1]
    ev = O.to_mont([pyref.horner(pc, x) for x in xs])
    av = O.to_mont([pyref.horner(ac, x) for x in xs])
    h = st.redc_z0(ev, av)
    hc = O.from_mont(tree64.exit(h))
    assert all(v == 0 for v in hc[n // 2:])
    z = pyref.poly_from_roots(xs[0::2])
    lhs = pyref.poly_divmod(pyref.poly_mul(hc[: n // 2], z), ac)[1]
    rhs = pyref.poly_divmod(pc, ac)[1]
    assert lhs == rhs
    assert (st.redc_z1(ev, av) == st.redc_z1(ev, av)).all()


def test_modular_reduce_is_polynomial_remainder(tree256):
    """MOD(P, A, <Z_0^2 mod A>) = P mod A on S (src/fftree.rs:277-289); also the identity EXIT relies on"""
    n = 64
    st = tree256.subtree_with_size(n)
    xs = O.from_mont(st.leaves())
    pc = rand_coeffs(n, 70)
    ac = rand_coeffs(n // 2, 71) + [1]
    z0 = pyref.poly_from_roots(xs[0::2])
    cc = pyref.poly_divmod(pyref.poly_mul(z0, z0), ac)[1]
    ev = lambda poly: O.to_mont([pyref.horner(poly, x) for x in xs])
    got = O.from_mont(st.modular_reduce(ev(pc), ev(ac), ev(cc)))
    rem = pyref.poly_divmod(pc, ac)[1]
    assert got == [pyref.horner(rem, x) for x in xs]
    # with the tree's own tables: P mod X^(n/2) = low half of the coefficients
    got = st.modular_reduce(ev(pc), st.table("xnn_s"), st.table("z0z0_rem_xnn_s"))
    assert O.from_mont(got) == [pyref.horner(pc[: n // 2], x) for x in xs]


def test_threaded_enter_is_identical():
    t = O.OracleTree.build(4096, parts=1)
    x = O.random_elements(4096, seed=3)
    assert (t.enter(x, threads=8) == t.enter(x)).all()


def test_errors_where_reference_panics(tree64):
    with pytest.raises(O.OracleError):
        tree64.enter(O.random_elements(128))        # "FFTree is too small"
    with pytest.raises(O.OracleError):
        tree64.enter(O.random_elements(3))          # not a power of two
    assert O.lib().orc_build_fftree(1 << 36, 1) is None
    with pytest.raises(O.OracleError):
        O.OracleTree.deserialize(tree64.serialize(True)[:-1], True)


def test_random_elements_are_canonical_and_seeded():
    a = O.random_elements(1000, seed=1)
    b = O.random_elements(1000, seed=1)
    assert (a == b).all() and not (a == O.random_elements(1000, seed=2)).all()
    for row in a[:50].tolist():
        assert row[0] | (row[1] << 64) | (row[2] << 128) | (row[3] << 192) < P
