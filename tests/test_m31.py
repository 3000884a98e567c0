"""FFTree<m31::Fp> (reference src/lib.rs:190-280).

CPU: the Python restatement (oracle/m31_ref.py) against the reference's own m31 tests (src/lib.rs:230-278) and the
definition-level checks the secp256k1 oracle gets (tests/test_oracle.py).
GPU: the CUDA path through the C ABI against that restatement, bit for bit: every table of every chain level,
all nine methods at sizes that cross the shared-memory tile, error codes; larger sizes through Horner at sampled
leaves and EXIT(ENTER(c)) == c."""
import random

import numpy as np
import pytest

from oracle import m31_ref as M

P = M.P


@pytest.fixture(scope="module")
def ref64():
    return M.FFTree.build(64)


def test_reference_evaluates_polynomial(ref64):
    """src/lib.rs:230-243: ENTER of a degree n-1 polynomial == its evaluations on the eval domain (n = 64)"""
    rng = random.Random(1)
    c = [rng.randrange(P) for _ in range(64)]
    assert ref64.enter(c) == [M.horner(c, x) for x in ref64.subtree_with_size(64).eval_domain()]


def test_reference_interpolates_evaluations(ref64):
    """src/lib.rs:245-257"""
    coeffs = [1, 1, 5, 0, 0, 1, 0, 0]
    assert ref64.exit(ref64.enter(coeffs)) == coeffs


def test_reference_determines_degree(ref64):
    """src/lib.rs:259-271"""
    assert ref64.degree(ref64.enter([1, 1, 1, 0, 0, 1, 0, 0])) == 5


def test_reference_definitions(ref64):
    """what the reference leaves untested: EXTEND both ways, VANISH, the Z tables, MOD = remainder, REDC"""
    rng = random.Random(2)
    t = ref64
    dom = t.eval_domain()
    s0, s1 = dom[0::2], dom[1::2]
    half = [rng.randrange(P) for _ in range(32)]
    assert t.extend([M.horner(half, x) for x in s0], 1) == [M.horner(half, x) for x in s1]
    assert t.extend([M.horner(half, x) for x in s1], 0) == [M.horner(half, x) for x in s0]

    def prod(x, pts):
        r = 1
        for a in pts:
            r = r * (x - a) % P
        return r
    assert t.z0_s1 == [prod(x, s0) for x in s1]
    assert t.z1_s0 == [prod(x, s1) for x in s0]
    d = [rng.randrange(P) for _ in range(32)]
    assert t.vanish(d) == [prod(x, d) for x in dom]
    c = [rng.randrange(P) for _ in range(64)]
    assert t.modular_reduce(t.enter(c), t.xnn_s, t.z0z0) == t.enter(c[:32] + [0] * 32)
    for deg in (0, 1, 31, 32, 61):
        cc = [rng.randrange(1, P) for _ in range(deg + 1)] + [0] * (63 - deg)
        assert t.degree(t.enter(cc)) == deg
    monic = [rng.randrange(P) for _ in range(32)]
    assert t.mextend([(M.horner(monic, x) + pow(x, 32, P)) % P for x in s0], 1) == \
        [(M.horner(monic, x) + pow(x, 32, P)) % P for x in s1]
    h = t.redc_z0(t.enter(c), t.xnn_s)
    assert t.degree(h) < 32
    assert len(M.cubic_roots(M.CURVE_A, M.CURVE_B)) >= 1


def test_symmetric_butterflies_equal_the_matrix_network(ref64):
    """The one-product butterflies of m31.cu (k31_extend<true>): in the coordinate y = x - x0 every map of the chain is
    y + t/y + x0 with t a square, so the two nodes of a pair are y and beta^2/y and g = (y - beta)/(y + beta) takes
    opposite values on them; EXTEND = Gamma^tgt . prod [[1, g],[1, -g]] . prod ([[1, g],[1, -g]]^src)^-1 . 2^-L / Gamma^src
    with Gamma_p = prod_j (y_j(p) + beta_j) y_j(p)^(2^j - 1) — bit-identical to the reference's 2x2 matrix network."""
    rng = random.Random(3)

    def sqrt_p(a):
        r = pow(a, (P + 1) // 4, P)
        assert r * r % P == a % P, "t must be a square for the symmetric form"
        return r

    for N in (4, 8, 64):
        t = ref64.subtree_with_size(N)
        h, L, logN = N // 2, (N // 2).bit_length() - 1, N.bit_length() - 1
        x0 = [(-t.maps[logN - 2 - j][1][0]) % P for j in range(L)]
        beta = [sqrt_p(t.maps[logN - 2 - j][0][0]) for j in range(L)]

        def y(mu, j, i, bit):
            return (t.f[(4 << j) + 2 * i + mu + bit * (2 << j)] - x0[j]) % P

        def g(mu, j, i):
            return (y(mu, j, i, 0) - beta[j]) * pow(y(mu, j, i, 0) + beta[j], -1, P) % P

        def gamma(mu, p):
            acc = 1
            for j in range(L):
                yy = y(mu, j, p & ((1 << j) - 1), (p >> j) & 1)
                acc = acc * (yy + beta[j]) * pow(yy, (1 << j) - 1, P) % P
            return acc

        for mu in (0, 1):
            for j in range(L):
                for i in range(1 << j):
                    assert y(mu, j, i, 0) * y(mu, j, i, 1) % P == beta[j] ** 2 % P
        x = [rng.randrange(P) for _ in range(h)]
        for target in (0, 1):
            src = 1 - target
            v = [xi * pow(2, -L, P) * pow(gamma(src, p), -1, P) % P for p, xi in enumerate(x)]
            for j in range(L - 1, -1, -1):
                for p in range(h):
                    if not (p >> j) & 1:
                        q, gi = p + (1 << j), pow(g(src, j, p & ((1 << j) - 1)), -1, P)
                        v[p], v[q] = (v[p] + v[q]) % P, (v[p] - v[q]) * gi % P
            for j in range(L):
                for p in range(h):
                    if not (p >> j) & 1:
                        q, tt = p + (1 << j), g(target, j, p & ((1 << j) - 1)) * v[p + (1 << j)] % P
                        v[p], v[q] = (v[p] + tt) % P, (v[p] - tt) % P
            assert [vi * gamma(target, p) % P for p, vi in enumerate(v)] == t.extend(x, target), (N, target)


def test_chain_constants():
    """src/lib.rs:199-206: both points are on y^2 = x^3 + x and the generator has order exactly 2^28"""
    for (x, y) in (M.COSET_OFFSET, M.SUBGROUP_GENERATOR):
        assert (y * y - x * x * x - x) % P == 0
    assert M.two_adicity(M.SUBGROUP_GENERATOR, M.CURVE_A, M.CURVE_B) == M.SUBGROUP_TWO_ADICITY


# ---- golden vectors from the real crate, when a maintainer has produced them (tools/rust_golden) ------------------------
import os

_GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = os.environ.get("ECFFT_ARKWORKS_M31_GOLDEN") or os.path.join(_GOLDEN_DIR, "arkworks_m31_n64.txt")
if not os.path.exists(GOLDEN):   # tools/rust_golden prints both fields to one stream: its whole output may sit in arkworks_n64.txt
    _both = os.path.join(_GOLDEN_DIR, "arkworks_n64.txt")
    if os.path.exists(_both) and "vec32 m31." in open(_both).read():
        GOLDEN = _both


def parse_golden(path):
    out = {}
    with open(path) as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 2 and parts[0] == "vec32" and parts[1].startswith("m31."):
                raw = bytes.fromhex(parts[2]) if len(parts) > 2 else b""
                out[parts[1][4:]] = np.frombuffer(raw, dtype="<u4").astype(np.uint32).tolist()
            elif len(parts) == 3 and parts[0] == "num" and parts[1].startswith("m31."):
                out[parts[1][4:]] = int(parts[2])
    return out


def check_against_golden(tree, g):
    """every vector of the file against an FFTree-like object (the restatement, or the CUDA path through lists)"""
    n = len(g["enter.in"])
    assert tree.eval_domain() == g["leaves"]
    assert tree.enter(g["enter.in"]) == g["enter.out"]
    assert tree.exit(g["exit.in"]) == g["exit.out"]
    assert tree.extend(g["extend.in"], 1) == g["extend_s1.out"]
    assert tree.extend(g["extend.in"], 0) == g["extend_s0.out"]
    assert tree.mextend(g["extend.in"], 1) == g["mextend_s1.out"]
    assert tree.redc_z0(g["exit.in"], g["xnn_s"]) == g["redc_z0.out"]
    assert tree.modular_reduce(g["exit.in"], g["xnn_s"], g["z0z0_rem_xnn_s"]) == g["mod.out"]
    assert tree.vanish(g["extend.in"]) == g["vanish.out"]
    assert tree.degree(g["enter.out"]) == g["degree.out"]
    assert n == 64


def write_golden_from_restatement(path):
    rng = random.Random(11)
    t = M.FFTree.build(64)
    vec = {"leaves": t.eval_domain(), "xnn_s": t.xnn_s, "z0z0_rem_xnn_s": t.z0z0}
    vec["enter.in"] = [rng.randrange(P) for _ in range(64)]
    vec["enter.out"] = t.enter(vec["enter.in"])
    vec["exit.in"] = [rng.randrange(P) for _ in range(64)]
    vec["exit.out"] = t.exit(vec["exit.in"])
    vec["extend.in"] = [rng.randrange(P) for _ in range(32)]
    vec["extend_s1.out"], vec["extend_s0.out"] = t.extend(vec["extend.in"], 1), t.extend(vec["extend.in"], 0)
    vec["mextend_s1.out"] = t.mextend(vec["extend.in"], 1)
    vec["redc_z0.out"] = t.redc_z0(vec["exit.in"], t.xnn_s)
    vec["mod.out"] = t.modular_reduce(vec["exit.in"], t.xnn_s, t.z0z0)
    vec["vanish.out"] = t.vanish(vec["extend.in"])
    with open(path, "w") as f:
        for k, v in vec.items():
            f.write(f"vec32 m31.{k} {np.asarray(v, dtype='<u4').tobytes().hex()}\n")
        f.write(f"num m31.degree.out {t.degree(vec['enter.out'])}\n")


def test_golden_consumer_on_a_file_written_by_the_restatement(tmp_path):
    """exercises the parser and the comparison (the file format of tools/rust_golden's m31 section)"""
    path = str(tmp_path / "m31.txt")
    write_golden_from_restatement(path)
    check_against_golden(M.FFTree.build(64), parse_golden(path))


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="no arkworks m31 golden file (tools/rust_golden needs cargo; not available in this image)")
def test_restatement_against_arkworks_golden():
    check_against_golden(M.FFTree.build(64), parse_golden(GOLDEN))


class _GpuAsLists:
    """the CUDA tree behind the list-based interface check_against_golden uses"""
    def __init__(self, t):
        self.t = t

    def eval_domain(self):
        return self.t.eval_domain(64).tolist()

    def __getattr__(self, name):
        fn = getattr(self.t, name)

        def call(*args):
            conv = [np.asarray(a, dtype=np.uint32) if isinstance(a, list) else a for a in args]
            r = fn(*conv)
            return r.tolist() if hasattr(r, "tolist") else r
        return call


@pytest.mark.gpu
def test_cuda_path_against_golden_file(tmp_path):
    """the CUDA path against the arkworks file when present, else against the file the restatement writes"""
    import ecfft_b200
    path = GOLDEN
    if not os.path.exists(path):
        path = str(tmp_path / "m31.txt")
        write_golden_from_restatement(path)
    check_against_golden(_GpuAsLists(ecfft_b200.m31.build_fftree(64)), parse_golden(path))


# ---- GPU ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def trees():
    import ecfft_b200
    n = 1 << 10
    return ecfft_b200.m31.build_fftree(n), M.FFTree.build(n)


def _u32(v):
    return np.asarray(v, dtype=np.uint32)


@pytest.mark.gpu
def test_tables_equal_the_restatement(trees):
    gpu, ref = trees
    t = ref
    while t is not None and t.n >= 1:
        n = t.n
        assert gpu.table("f", n).tolist() == t.f
        assert gpu.table("recombine_matrices", n).reshape(-1, 4).tolist() == [list(m) for m in t.rmat]
        assert gpu.table("decompose_matrices", n).reshape(-1, 4).tolist() == [list(m) for m in t.dmat]
        for name, want in (("xnn_s", t.xnn_s), ("xnn_s_inv", t.xnn_s_inv), ("z0_s1", t.z0_s1), ("z1_s0", t.z1_s0),
                           ("z0_inv_s1", t.z0_inv_s1), ("z1_inv_s0", t.z1_inv_s0), ("z0z0_rem_xnn_s", t.z0z0),
                           ("z1z1_rem_xnn_s", t.z1z1)):
            assert gpu.table(name, n).tolist() == list(want), (name, n)
        t = t.subtree


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 6, 9, 10])
def test_algorithms_equal_the_restatement(trees, log_n):
    gpu, ref = trees
    rng = random.Random(100 + log_n)
    n = 1 << log_n
    c = [rng.randrange(P) for _ in range(n)]
    ev = ref.enter(c)
    assert gpu.enter(_u32(c)).tolist() == ev
    assert gpu.exit(_u32(ev)).tolist() == c
    assert gpu.degree(_u32(ev)) == ref.degree(ev)
    if n >= 2:
        h = n // 2
        half = [rng.randrange(P) for _ in range(h)]
        for mo in (0, 1):
            assert gpu.extend(_u32(half), mo).tolist() == ref.extend(half, mo)
            assert gpu.mextend(_u32(half), mo).tolist() == ref.mextend(half, mo)
        a = [rng.randrange(1, P) for _ in range(n)]
        cc = [rng.randrange(P) for _ in range(n)]
        assert gpu.redc_z0(_u32(ev), _u32(a)).tolist() == ref.redc_z0(ev, a)
        assert gpu.redc_z1(_u32(ev), _u32(a)).tolist() == ref.redc_z1(ev, a)
        assert gpu.modular_reduce(_u32(ev), _u32(a), _u32(cc)).tolist() == ref.modular_reduce(ev, a, cc)
        t = ref.subtree_with_size(n)
        assert gpu.modular_reduce(_u32(ev), _u32(t.xnn_s), _u32(t.z0z0)).tolist() == ref.modular_reduce(ev, t.xnn_s, t.z0z0)
        assert gpu.vanish(_u32(half)).tolist() == ref.vanish(half)
        low = c[:h] + [0] * h
        assert gpu.degree(gpu.enter(_u32(low))) == ref.degree(ref.enter(low))


@pytest.mark.gpu
def test_reference_tests_on_the_gpu():
    """src/lib.rs:230-271 through the CUDA path (tree of 64 leaves, as the reference builds)"""
    import ecfft_b200
    t = ecfft_b200.m31.build_fftree(64)
    rng = random.Random(1)
    c = [rng.randrange(P) for _ in range(64)]
    dom = t.eval_domain(64).tolist()
    assert t.enter(_u32(c)).tolist() == [M.horner(c, x) for x in dom]
    coeffs = [1, 1, 5, 0, 0, 1, 0, 0]
    assert t.exit(t.enter(_u32(coeffs))).tolist() == coeffs
    assert t.degree(t.enter(_u32([1, 1, 1, 0, 0, 1, 0, 0]))) == 5


@pytest.mark.gpu
def test_large_sizes_by_properties():
    """n = 2^18 (multi-pass EXTEND: strided tiles): Horner at sampled leaves, EXIT(ENTER) = id, EXTEND both ways,
    device-tensor calls equal host-buffer calls"""
    import torch
    import ecfft_b200
    n = 1 << 18
    t = ecfft_b200.m31.build_fftree(n)
    rng = np.random.default_rng(7)
    c = rng.integers(0, P, size=n, dtype=np.uint32)
    ev = t.enter(c)
    dom = t.eval_domain()
    cl = c.tolist()
    for i in (0, 1, 2, n // 2 - 1, n // 2, n - 1, 12345):
        assert int(ev[i]) == M.horner(cl, int(dom[i]))
    assert (t.exit(ev) == c).all()
    lowdeg = c.copy()
    lowdeg[n // 2:] = 0
    e2 = t.enter(lowdeg)
    assert (t.extend(e2[0::2].copy(), 1) == e2[1::2]).all()
    assert (t.extend(e2[1::2].copy(), 0) == e2[0::2]).all()
    assert t.degree(e2) == int(np.nonzero(lowdeg)[0].max())
    d = t.enter(torch.from_numpy(c.view(np.int32)).cuda())
    assert (d.cpu().numpy().view(np.uint32) == ev).all()
    assert (t.exit(d).cpu().numpy().view(np.uint32) == c).all()


@pytest.mark.gpu
def test_error_codes():
    import ecfft_b200
    from ecfft_b200 import _lib
    t = ecfft_b200.m31.build_fftree(16)
    with pytest.raises(ecfft_b200.EcfftError) as e:
        t.enter(np.zeros(12, dtype=np.uint32))
    assert e.value.code == _lib.ERR_NOT_POW2          # src/fftree.rs:490
    with pytest.raises(ecfft_b200.EcfftError) as e:
        t.enter(np.zeros(32, dtype=np.uint32))
    assert e.value.code == _lib.ERR_TREE_TOO_SMALL    # src/fftree.rs:494
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.m31.build_fftree(1 << 29)
    assert e.value.code == _lib.ERR_TOO_LARGE         # src/ec.rs:513-515: None


def test_m31_arguments_are_checked_without_a_device():
    """argument errors come before anything that needs a GPU"""
    import ecfft_b200
    from ecfft_b200 import _lib
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.m31.build_fftree(12)
    assert e.value.code == _lib.ERR_NOT_POW2
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.m31.build_fftree(1 << 29)
    assert e.value.code == _lib.ERR_TOO_LARGE
