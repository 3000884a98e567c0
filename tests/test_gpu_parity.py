"""GPU parity: the CUDA engine, called through the C ABI (ecfft_b200.FFTree -> ctypes ->
libecfft_b200.so), against the CPU oracle on the same seeded inputs — bit-exact.

Mirrors the reference's own tests (src/lib.rs:108-186) and adds the operations it leaves
untested on secp256k1.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES_FULL = [2, 4, 8, 64, 1024, 8192]      # full trees (all tables), oracle builds in seconds
TABLES = ["f", "recombine_matrices", "decompose_matrices", "xnn_s", "xnn_s_inv", "z0_s1", "z1_s0",
          "z0_inv_s1", "z1_inv_s0", "z0z0_rem_xnn_s", "z1z1_rem_xnn_s"]
ORACLE_NAMES = {"recombine_matrices": "recombine", "decompose_matrices": "decompose"}


@pytest.fixture(scope="module")
def trees(oracle_mod):
    import ecfft_b200
    n = 1 << 14
    gpu = ecfft_b200.build_fftree(n)
    cpu = oracle_mod.OracleTree.build(n)
    return gpu, cpu


def eq(a, b):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.asarray(b, dtype=np.uint64).reshape(-1, 4)
    assert a.shape == b.shape
    bad = np.nonzero((a != b).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} of {len(a)} elements differ, first at {bad[:8]}"


@pytest.mark.parametrize("sub", [1, 2, 4, 16, 256, 4096, 1 << 14])
def test_tree_tables_match_oracle(trees, sub):
    """GPU-built FFTree (src/fftree.rs:318-463 on device) == oracle's, every pub field, every chain level"""
    gpu, cpu = trees
    st = cpu.subtree_with_size(sub)
    for name in TABLES:
        eq(gpu.table(name, sub), st.table(ORACLE_NAMES.get(name, name)))


@pytest.mark.parametrize("n", [1, 2, 4, 8, 64, 1024, 4096, 8192, 1 << 14])
def test_enter(trees, oracle_mod, n):
    gpu, cpu = trees
    x = oracle_mod.random_elements(n, seed=n)
    eq(gpu.enter(x), cpu.enter(x))


@pytest.mark.parametrize("n", [1, 2, 4, 64, 1024, 4096, 1 << 13])
@pytest.mark.parametrize("moiety", [0, 1])
def test_extend(trees, oracle_mod, n, moiety):
    gpu, cpu = trees
    x = oracle_mod.random_elements(n, seed=100 + n)
    eq(gpu.extend(x, moiety), cpu.extend(x, moiety))


@pytest.mark.parametrize("n", [1, 2, 4, 64, 1024, 1 << 13])
@pytest.mark.parametrize("moiety", [0, 1])
def test_mextend(trees, oracle_mod, n, moiety):
    gpu, cpu = trees
    x = oracle_mod.random_elements(n, seed=200 + n)
    eq(gpu.mextend(x, moiety), cpu.mextend(x, moiety))


@pytest.mark.parametrize("n", [1, 2, 4, 8, 64, 1024, 8192, 1 << 14])
def test_exit(trees, oracle_mod, n):
    gpu, cpu = trees
    x = oracle_mod.random_elements(n, seed=300 + n)
    eq(gpu.exit(x), cpu.exit(x))


@pytest.mark.parametrize("n", [2, 8, 64, 4096, 1 << 14])
def test_exit_inverts_enter(trees, oracle_mod, n):
    gpu, _ = trees
    x = oracle_mod.random_elements(n, seed=400 + n)
    eq(gpu.exit(gpu.enter(x)), x)


@pytest.mark.parametrize("n", [2, 4, 64, 1024, 1 << 14])
def test_redc_and_mod(trees, oracle_mod, n):
    gpu, cpu = trees
    x = oracle_mod.random_elements(n, seed=500 + n)
    a = oracle_mod.random_elements(n, seed=600 + n)
    c = oracle_mod.random_elements(n, seed=700 + n)
    eq(gpu.redc_z0(x, a), cpu.redc_z0(x, a))
    eq(gpu.redc_z1(x, a), cpu.redc_z1(x, a))
    eq(gpu.modular_reduce(x, a, c), cpu.modular_reduce(x, a, c))
    # the reference's bench arguments (benches/fftree.rs:48-54): tree tables as a and c
    st = cpu.subtree_with_size(n)
    xa, zz = st.table("xnn_s"), st.table("z0z0_rem_xnn_s")
    eq(gpu.redc_z0(x, xa), cpu.redc_z0(x, xa))
    eq(gpu.modular_reduce(x, xa, zz), cpu.modular_reduce(x, xa, zz))


def test_redc_with_zero_in_a(trees, oracle_mod):
    """ark_ff::batch_inversion leaves zeros untouched"""
    gpu, cpu = trees
    n = 64
    x = oracle_mod.random_elements(n, seed=1)
    a = oracle_mod.random_elements(n, seed=2)
    a[0] = 0
    a[10] = 0
    eq(gpu.redc_z0(x, a), cpu.redc_z0(x, a))


@pytest.mark.parametrize("n", [1, 2, 4, 64, 1024, 1 << 13])
def test_vanish(trees, oracle_mod, n):
    gpu, cpu = trees
    x = oracle_mod.random_elements(n, seed=800 + n)
    eq(gpu.vanish(x), cpu.vanish(x))


@pytest.mark.parametrize("n", [1, 2, 64, 1024, 1 << 14])
def test_degree(trees, oracle_mod, n):
    gpu, cpu = trees
    for d in sorted({0, 1, n // 2 - 1, n // 2, n - 3, n - 1}):
        if d < 0 or d >= n:
            continue
        c = oracle_mod.random_elements(n, seed=900 + n + d)
        c[d + 1:] = 0
        ev = cpu.enter(c)
        assert gpu.degree(ev) == cpu.degree(ev) == d


def test_serialization_bytes_match_oracle_and_round_trip(oracle_mod):
    """src/lib.rs:154-186: deserialised trees (both modes) still evaluate polynomials; the bytes
    themselves equal the oracle's restatement of the arkworks layout"""
    import ecfft_b200
    n = 64
    gpu = ecfft_b200.build_fftree(n)
    cpu = oracle_mod.OracleTree.build(n)
    x = oracle_mod.random_elements(n, seed=5)
    want = cpu.enter(x)
    for compressed in (True, False):
        blob = gpu.serialize(compressed)
        assert blob == cpu.serialize(compressed)
        again = ecfft_b200.FFTree.deserialize(blob, compressed)
        eq(again.enter(x), want)
        eq(again.exit(want), x)
        assert again.serialize(compressed) == blob


def test_deserialize_rejects_bad_bytes(oracle_mod):
    import ecfft_b200
    from ecfft_b200 import _lib
    cpu = oracle_mod.OracleTree.build(16)
    blob = bytearray(cpu.serialize(False))
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.FFTree.deserialize(bytes(blob[:-5]), False)
    assert e.value.code == _lib.ERR_BAD_BYTES
    blob[8 + 32:8 + 64] = b"\xff" * 32  # f[1] >= p
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.FFTree.deserialize(bytes(blob), False)
    assert e.value.code == _lib.ERR_BAD_BYTES


def test_errors_mirror_reference_panics(trees, oracle_mod):
    import ecfft_b200
    from ecfft_b200 import _lib
    gpu, _ = trees
    with pytest.raises(ecfft_b200.EcfftError) as e:
        gpu.enter(oracle_mod.random_elements(3))
    assert e.value.code == _lib.ERR_NOT_POW2
    with pytest.raises(ecfft_b200.EcfftError) as e:
        gpu.enter(oracle_mod.random_elements(1 << 15))
    assert e.value.code == _lib.ERR_TREE_TOO_SMALL
    with pytest.raises(ecfft_b200.EcfftError) as e:
        gpu.extend(oracle_mod.random_elements(1 << 14), 1)
    assert e.value.code == _lib.ERR_TREE_TOO_SMALL
    assert ecfft_b200.build_fftree(1 << 36) is None  # src/lib.rs:61-64


def test_enter_only_tree_and_device_tensors(oracle_mod):
    import torch
    import ecfft_b200
    from ecfft_b200 import _lib
    n = 1 << 13
    gpu = ecfft_b200.build_fftree(n, parts=ecfft_b200.PARTS_ENTER_ONLY)
    cpu = oracle_mod.OracleTree.build(n, parts=1)
    x = oracle_mod.random_elements(n, seed=77)
    want = cpu.enter(x)
    eq(gpu.enter(x), want)
    xd = torch.from_numpy(x.view(np.int64)).cuda()
    out = gpu.enter(xd)
    torch.cuda.synchronize()
    eq(out.cpu().numpy().view(np.uint64), want)
    # sharded schedule == single call (the multi-GPU building block)
    G = 4
    parts = [gpu.enter_range(xd[g * (n // G):(g + 1) * (n // G)], 1, n // G) for g in range(G)]
    out2 = gpu.enter_range(torch.cat(parts), n // G, n)
    torch.cuda.synchronize()
    eq(out2.cpu().numpy().view(np.uint64), want)
    with pytest.raises(ecfft_b200.EcfftError) as e:
        gpu.exit(x)
    assert e.value.code == _lib.ERR_MISSING_TABLES


def test_tree_new_from_leaves(oracle_mod):
    """FFTree::new(leaves, rational_maps) (src/fftree.rs:42-70) reproduces build_fftree's tree"""
    import ecfft_b200
    from oracle import pyref
    n = 64
    ref = oracle_mod.OracleTree.build(n)
    leaves = ref.leaves()
    # rational maps of the Good-Curve chain: r = (x^2 - 2b x + b^2) / x  (src/ec.rs:84)
    p = pyref.P
    a, bb = pyref.A, pyref.BB
    maps = []
    for _ in range(6):
        b = pow(bb, (p + 1) // 4, p)
        maps.append((oracle_mod.to_mont([bb, (-2 * b) % p, 1]), oracle_mod.to_mont([0, 1])))
        a, bb = (a + 6 * b) % p, (4 * a * b + 8 * b * b) % p
    t = ecfft_b200.FFTree.new(leaves, maps)
    for name in TABLES:
        eq(t.table(name), ref.table(ORACLE_NAMES.get(name, name)))


@pytest.mark.parametrize("mode", ["matrix", "normalised"])
def test_other_butterfly_modes_match_too(mode):
    """ECFFT_B200_BUTTERFLY=matrix selects the reference-shaped 2x2 mat-vec butterflies (the measured
    'phase 1' kernel), =normalised the two-product twiddle form (also the fallback for trees whose maps are
    not (x^2 + c1 x + beta^2)/x); the default is the one-product symmetric form.  All give the oracle's bits."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, ecfft_b200\n"
        "from oracle import oracle as O\n"
        "n = 1 << 13\n"
        "g = ecfft_b200.build_fftree(n); c = O.OracleTree.build(n)\n"
        "x = O.random_elements(n, seed=3)\n"
        "assert (g.enter(x) == c.enter(x)).all()\n"
        "assert (g.exit(x) == c.exit(x)).all()\n"
        "assert (g.extend(x[:n//2], 0) == c.extend(x[:n//2], 0)).all()\n"
        "assert (g.mextend(x[:n//2], 1) == c.mextend(x[:n//2], 1)).all()\n"
        "print('matrix mode ok')\n")
    env = dict(os.environ, ECFFT_B200_BUTTERFLY=mode)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "matrix mode ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("variant", ["0", "1"])
def test_flow_launch_matches_oracle(variant):
    """ECFFT_B200_FLOW=1 runs all passes of an ENTER as ONE persistent launch with per-block dependency counters
    instead of kernel boundaries (csrc/sym_kernel.cu k_sym_flow; off by default, measured slower).  Same bits:
    sizes that take packed tiles only, the combine-only pass (vector = tile), one, two and three outer strided
    passes, a batch that is not a power of two (enter_range over three blocks), and the host-buffer pipeline."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, ecfft_b200\n"
        "from oracle import oracle as O\n"
        "n = 1 << 18\n"
        "g = ecfft_b200.build_fftree(n, parts=ecfft_b200.PARTS_ENTER_ONLY); c = O.OracleTree.build(n, parts=1, threads=8)\n"
        "L = ecfft_b200._lib.load(); l0 = L.ecfft_launch_count()\n"
        "for k in (10, 11, 12, 14, 16, 17, 18):\n"
        "    x = O.random_elements(1 << k, seed=k)\n"
        "    xd = torch.from_numpy(x.view(np.int64)).cuda()\n"
        "    assert (g.enter(xd).cpu().numpy().view(np.uint64) == c.enter(x, threads=8)).all(), k\n"
        "x = O.random_elements(3 << 12, seed=5)\n"
        "xd = torch.from_numpy(x.view(np.int64)).cuda()\n"
        "got = g.enter_range(xd, 1, 1 << 12).cpu().numpy().view(np.uint64)\n"
        "for b in range(3):\n"
        "    assert (got[b << 12:(b + 1) << 12] == c.enter(x[b << 12:(b + 1) << 12])).all()\n"
        "x = O.random_elements(n, seed=6)\n"
        "assert (g.enter(x) == c.enter(x, threads=8)).all()\n"
        "assert L.ecfft_launch_count() - l0 < 40, 'flow path was not taken'\n"
        "print('flow ok')\n")
    env = dict(os.environ, ECFFT_B200_FLOW="1", ECFFT_B200_SYM_VARIANT=variant)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "flow ok" in r.stdout, r.stdout + r.stderr


# ---- BASELINE.json full sizes: size-independent properties (the oracle needs minutes there) ----
@pytest.fixture(scope="module")
def tree22():
    import ecfft_b200
    return ecfft_b200.build_fftree(1 << 22)          # full tree: every table of every chain level


def _to_ints(a):
    return [r[0] | (r[1] << 64) | (r[2] << 128) | (r[3] << 192) for r in np.asarray(a, dtype=np.uint64).tolist()]


def test_full_size_enter_spot_checks_and_round_trip(tree22, oracle_mod):
    """ENTER n=2^22 (BASELINE configs[2]): Horner at sampled leaves with Python big ints, linearity,
    EXIT(ENTER(c)) == c, and the low-degree input cross-checked against the oracle on the subtree."""
    from oracle import pyref
    P, R = pyref.P, 2**256 % pyref.P
    n = 1 << 22
    c = oracle_mod.random_elements(n, seed=22)
    ev = tree22.enter(c)
    # (1) Horner at three leaves (values are Montgomery limbs: v~ = v*R)
    leaves = tree22.eval_domain()
    rinv = pow(R, -1, P)
    coeffs = [v * rinv % P for v in _to_ints(c)]
    for idx in (0, 1234567, n - 1):
        x = _to_ints(leaves[idx:idx + 1])[0] * rinv % P
        want = pyref.horner(coeffs, x)
        assert _to_ints(ev[idx:idx + 1])[0] * rinv % P == want
    # (2) linearity on a sample: enter(c + d) == enter(c) + enter(d)
    d = oracle_mod.random_elements(n, seed=23)
    cd = np.array([[(s >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)] for s in
                   [(a + b) % P for a, b in zip(_to_ints(c[:4096]), _to_ints(d[:4096]))]], dtype=np.uint64)
    c2 = c.copy()
    c2[4096:] = 0
    d2 = d.copy()
    d2[4096:] = 0
    cd_full = np.zeros_like(c)
    cd_full[:4096] = cd
    e1, e2, e3 = tree22.enter(c2), tree22.enter(d2), tree22.enter(cd_full)
    sample = np.random.default_rng(1).integers(0, n, 2000)
    for i in sample.tolist():
        a, b, s = (_to_ints(t[i:i + 1])[0] for t in (e1, e2, e3))
        assert (a + b) % P == s
    # (3) round trip through EXIT at full size
    eq(tree22.exit(ev), c)
    # (4) degree < 2^14 input: its evaluations on the 2^14-leaf subtree's domain come from the oracle
    small = oracle_mod.OracleTree.build(1 << 14, parts=1)
    c14 = oracle_mod.random_elements(1 << 14, seed=24)
    padded = np.zeros_like(c)
    padded[: 1 << 14] = c14
    big = tree22.enter(padded)
    eq(big[:: 1 << 8], small.enter(c14))            # subtree leaves = every 2^8-th leaf (src/fftree.rs:465-482)


def test_full_size_extend_redc_mod(tree22, oracle_mod):
    """EXTEND n=2^20 (configs[1]) and REDC+MOD n=2^20 (configs[4]) through their defining identities"""
    n = 1 << 20
    c = oracle_mod.random_elements(n, seed=30)
    ev = tree22.subtree_with_size(2 * n).enter(np.concatenate([c, np.zeros_like(c)]))   # deg < n on 2n leaves
    eq(tree22.extend(ev[0::2], 1), ev[1::2])
    eq(tree22.extend(ev[1::2], 0), ev[0::2])
    # MOD by X^(n/2) with the tree's own tables keeps the low half of the coefficients (fftree.rs:206-210)
    full = oracle_mod.random_elements(n, seed=31)
    evn = tree22.enter(full)
    xnn = tree22.table("xnn_s", n)
    zz = tree22.table("z0z0_rem_xnn_s", n)
    low = full.copy()
    low[n // 2:] = 0
    eq(tree22.modular_reduce(evn, xnn, zz), tree22.enter(low))
    # REDC: h = redc_z0(P, X^(n/2)) has degree < n/2, and redc(redc(P) * c) == MOD
    h = tree22.redc_z0(evn, xnn)
    assert tree22.degree(h) < n // 2


# ---- oracle-exact checks at sizes that take the multi-outer-pass path of the EXTEND kernel (log_h >= 16) ----
def test_large_oracle_exact_enter_2p18(tree22, oracle_mod):
    """ENTER n = 2^18 bit for bit against the oracle (its top depths run EXTENDs of 2^16 and 2^17 elements: two
    outer strided passes + the inner pass + the fused combine pass)."""
    import os
    n = 1 << 18
    cpu = oracle_mod.OracleTree.build(n, parts=1, threads=os.cpu_count() or 1)
    x = oracle_mod.random_elements(n, seed=218)
    eq(tree22.enter(x), cpu.enter(x, threads=os.cpu_count() or 1))


def test_large_oracle_exact_exit_extend_redc_2p17(tree22, oracle_mod):
    """EXIT, EXTEND (both moieties), REDC and MOD at n = 2^17 bit for bit against the oracle (full tables)."""
    import os
    n = 1 << 17
    cpu = oracle_mod.OracleTree.build(n, threads=os.cpu_count() or 1)
    x = oracle_mod.random_elements(n, seed=217)
    ev = cpu.enter(x, threads=os.cpu_count() or 1)
    eq(tree22.enter(x), ev)
    eq(tree22.exit(ev), x)
    y = oracle_mod.random_elements(n, seed=2170)
    eq(tree22.exit(y), cpu.exit(y))                  # evaluations of no low-degree polynomial in particular
    half = oracle_mod.random_elements(n // 2, seed=2171)
    for moiety in (0, 1):
        eq(tree22.subtree_with_size(n).extend(half, moiety), cpu.extend(half, moiety))
    a = cpu.table("xnn_s")
    c = cpu.table("z0z0_rem_xnn_s")
    eq(tree22.redc_z0(y, a), cpu.redc_z0(y, a))
    eq(tree22.modular_reduce(y, a, c), cpu.modular_reduce(y, a, c))


def test_device_calls_reject_operands_of_different_length(trees, oracle_mod):
    """redc / modular_reduce with a shorter `a` or `c` CUDA tensor must be refused (the C entry points take
    ONE n), and degree() validates its tensor like the other calls."""
    import torch
    import ecfft_b200
    gpu, _ = trees
    n = 1024
    x = torch.from_numpy(oracle_mod.random_elements(n, seed=5).view(np.int64)).cuda()
    short = x[: n // 2].contiguous()
    for call in (lambda: gpu.redc_z0(x, short), lambda: gpu.redc_z1(x, short),
                 lambda: gpu.modular_reduce(x, x, short), lambda: gpu.modular_reduce(x, short, x)):
        with pytest.raises(ecfft_b200.EcfftError) as e:
            call()
        assert e.value.code == ecfft_b200._lib.ERR_INVALID_ARG
    with pytest.raises(ecfft_b200.EcfftError):
        gpu.degree(x.to(torch.int32))
    with pytest.raises(ecfft_b200.EcfftError):
        gpu.degree(x[:, :2])


def test_tree_new_rejects_degenerate_and_noncanonical_input(trees, oracle_mod):
    """FFTree::new panics in the reference on a zero denominator / singular matrix (unwrap at src/fftree.rs:57-58,
    :361); here the build fails with ERR_INVALID_ARG, and so do limbs >= p."""
    import ecfft_b200
    from oracle import pyref
    n = 16
    leaves = oracle_mod.OracleTree.build(n).leaves().copy()
    zero = np.zeros(4, dtype=np.uint64)
    p = pyref.P
    a, bb = pyref.A, pyref.BB
    maps = []
    for _ in range(4):                               # r = (x^2 - 2b x + b^2) / x  (src/ec.rs:84)
        b = pow(bb, (p + 1) // 4, p)
        maps.append((oracle_mod.to_mont([bb, (-2 * b) % p, 1]), oracle_mod.to_mont([0, 1])))
        a, bb = (a + 6 * b) % p, (4 * a * b + 8 * b * b) % p
    ok = ecfft_b200.FFTree.new(leaves, maps)
    assert ok.leaves_count == n
    bad = leaves.copy()
    bad[3] = zero                                    # x = 0 is the pole of every map (x - b)^2 / x
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.FFTree.new(bad, maps)
    assert e.value.code == ecfft_b200._lib.ERR_INVALID_ARG
    big = leaves.copy()
    big[0] = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)   # >= p
    with pytest.raises(ecfft_b200.EcfftError) as e:
        ecfft_b200.FFTree.new(big, maps)
    assert e.value.code == ecfft_b200._lib.ERR_INVALID_ARG


def test_host_buffer_enter_pipeline_equals_device_path(tree22, oracle_mod):
    """ecfft_enter uploads three chunks on a copy stream and runs the low recursion depths on each while the
    next is in flight; every element must equal the single device-resident ENTER
    (n = 2^17 and 2^20: pipelined path)"""
    import torch
    for log_n, seed in ((17, 60), (20, 61)):
        n = 1 << log_n
        x = oracle_mod.random_elements(n, seed=seed)
        host = tree22.enter(x)                                                  # numpy in -> host-buffer ABI
        dev = tree22.enter(torch.from_numpy(x.view(np.int64)).cuda())           # device tensor -> _dev ABI
        eq(host, dev.cpu().numpy().view(np.uint64))


def test_enter_many_equals_single_calls(trees, tree22, oracle_mod):
    """ecfft_enter_many pipelines uploads, kernels and downloads of consecutive vectors over two buffer sets:
    every vector must equal its own ecfft_enter (oracle-exact at 2^12; odd and even counts, count 1, count 0)"""
    import ecfft_b200
    gpu, cpu = trees
    xs = np.stack([oracle_mod.random_elements(1 << 12, seed=70 + i) for i in range(5)])
    ys = gpu.enter_many(xs)
    for i in range(5):
        eq(ys[i], cpu.enter(xs[i]))
    eq(gpu.enter_many(xs[:1])[0], ys[0])
    assert gpu.enter_many(xs[:0]).shape == (0, 1 << 12, 4)
    big = np.stack([oracle_mod.random_elements(1 << 18, seed=80 + i) for i in range(4)])
    out = tree22.enter_many(big)
    for i in range(4):
        eq(out[i], tree22.enter(big[i]))
    with pytest.raises(ecfft_b200.EcfftError) as e:
        gpu.enter_many(np.zeros((2, 3, 4), dtype=np.uint64))      # src/fftree.rs:490: not a power of two
    assert e.value.code == ecfft_b200._lib.ERR_NOT_POW2


def test_folded_combines_equal_the_plain_schedule(tree22, oracle_mod):
    """ENTER stores the data between two depths pre-multiplied by the next EXTEND's pre-scale (Engine::enter_range_serial);
    enter_range starts and ends in plain form wherever the range is cut, so every cut of the depth range must
    give the same result as the whole ENTER — and that is oracle-exact on the 2^14 sub-lattice (other tests)."""
    import torch
    n = 1 << 16
    x = torch.from_numpy(oracle_mod.random_elements(n, seed=90).view(np.int64)).cuda()
    whole = tree22.enter(x)
    for cut in (2, 1 << 5, 1 << 10, 1 << 11, 1 << 15):
        part = tree22.enter_range(tree22.enter_range(x, 1, cut), cut, n)
        assert torch.equal(part, whole), cut


def test_device_field_selftest_drives_the_rare_carry_paths():
    """fp_add_lazy_f / fp_sub_lazy2_f branch on a carry out of the low two limbs (probability ~2^-31 on random
    data, i.e. about once per ENTER(2^22)): directed operands on the device itself, against canonical
    arithmetic — and proof that the rare paths were taken."""
    import ctypes
    from ecfft_b200 import _lib
    counters = (ctypes.c_ulonglong * 3)()
    _lib.check(_lib.load().ecfft_selftest_field(0, 6 << 20, counters))
    mismatches, add_ripples, sub_ripples = counters[0], counters[1], counters[2]
    assert mismatches == 0
    assert add_ripples > 100000 and sub_ripples > 100000


class _ThreadComm:
    """virtual ranks as threads on one device: FIFO mailboxes per (src, dst) pair"""

    def __init__(self, rank, world, boxes, slots, barrier):
        self.rank, self.world, self.boxes, self.slots, self.barrier = rank, world, boxes, slots, barrier

    def sendrecv(self, sends, recvs):
        for tensor, dst in sends:
            self.boxes[(self.rank, dst)].put(tensor.clone())
        return [self.boxes[(src, self.rank)].get(timeout=120) for _, _, src in recvs]

    def all_gather(self, tensor):
        import torch
        self.slots[self.rank] = tensor
        self.barrier.wait()
        out = torch.cat([self.slots[r] for r in range(self.world)])
        self.barrier.wait()
        return out


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_schedule_with_virtual_ranks(trees, oracle_mod, world):
    """The fully sharded multi-GPU ENTER (ecfft_b200/dist.py) with `world` virtual ranks on ONE GPU: every
    building block of the C ABI (ecfft_mg_*_dev, ecfft_enter_range_dev) in composition == the oracle"""
    import queue
    import threading
    import torch
    from ecfft_b200.dist import enter_sharded, enter_sharded_allgather
    gpu, cpu = trees
    n = 1 << 14
    x = oracle_mod.random_elements(n, seed=world)
    want = cpu.enter(x)
    xd = torch.from_numpy(x.view(np.int64)).cuda()
    c = n // world
    boxes = {(a, b): queue.Queue() for a in range(world) for b in range(world)}
    slots, barrier = [None] * world, threading.Barrier(world)
    results, errors = [None] * world, []

    def run(rank):
        try:
            comm = _ThreadComm(rank, world, boxes, slots, barrier)
            full = enter_sharded(gpu, xd[rank * c:(rank + 1) * c], n, comm=comm)
            part = enter_sharded(gpu, xd[rank * c:(rank + 1) * c], n, comm=comm, gather=False)
            ag = enter_sharded_allgather(gpu, xd[rank * c:(rank + 1) * c], n, comm=comm)
            torch.cuda.synchronize()
            results[rank] = (full.cpu().numpy().view(np.uint64), part.cpu().numpy().view(np.uint64), ag.cpu().numpy().view(np.uint64))
        except Exception as e:  # surface failures instead of deadlocking the other ranks
            errors.append(repr(e))
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    for rank in range(world):
        full, part, ag = results[rank]
        eq(full, want)
        eq(ag, want)
        eq(part, want[rank * c:(rank + 1) * c])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_schedule_with_virtual_ranks(trees, oracle_mod, world):
    """ecfft_b200.dist.enter_sharded_peer — butterfly / combine kernels reading the partner's buffers
    through arena pointers, ordered by ecfft_mg_signal_dev / ecfft_mg_wait_dev flags — with `world` virtual
    ranks (threads, one CUDA stream each) whose arenas live on ONE GPU.  Two calls per rank exercise the
    epoch counter and slot reuse; both the in-library schedule (ecfft_enter_peer_dev) and the step-by-step
    one run.  (Across GPUs the same pointers are CUDA-IPC mappings; bench.py --gpus N.)"""
    import threading
    import torch
    from ecfft_b200.dist import PeerArena, enter_sharded_peer
    gpu, cpu = trees
    n = 1 << 14
    x = oracle_mod.random_elements(n, seed=40 + world)
    want = cpu.enter(x)
    xd = torch.from_numpy(x.view(np.int64)).cuda()
    c = n // world
    arenas = PeerArena.local_group(n, world, device=0)
    slots, barrier = [None] * world, threading.Barrier(world)
    results, errors = [None] * world, []
    torch.cuda.synchronize()

    def run(rank):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                def all_gather(t):
                    stream.synchronize()
                    slots[rank] = t
                    barrier.wait()
                    out = torch.cat([slots[r] for r in range(world)])
                    stream.synchronize()
                    barrier.wait()
                    return out

                chunk = xd[rank * c:(rank + 1) * c]
                enter_sharded_peer(gpu, chunk, n, arenas[rank], all_gather=all_gather)
                full = enter_sharded_peer(gpu, chunk, n, arenas[rank], all_gather=all_gather)
                part = enter_sharded_peer(gpu, chunk, n, arenas[rank], gather=False, all_gather=all_gather)
                step = enter_sharded_peer(gpu, chunk, n, arenas[rank], all_gather=all_gather, native=False)
                stream.synchronize()
                results[rank] = (full.cpu().numpy().view(np.uint64), part.cpu().numpy().view(np.uint64),
                                 step.cpu().numpy().view(np.uint64))
        except Exception as e:  # surface failures instead of deadlocking the other ranks
            errors.append(repr(e))
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    for rank in range(world):
        full, part, step = results[rank]
        eq(full, want)
        eq(step, want)
        eq(part, want[rank * c:(rank + 1) * c])
    for a in arenas:
        a.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_exit_with_virtual_ranks(trees, oracle_mod, world):
    """ecfft_b200.dist.exit_sharded_peer (csrc/sharded.cu exit_peer; reference src/fftree.rs:200-224): the top
    log2(world) depths run MOD on vectors spread over `world` virtual ranks (threads, one stream each, arenas on
    ONE GPU), cross-rank butterfly levels reading the partner's arena, then the (u0 | v0) split moves each half
    to half of the ranks.  Against the oracle's EXIT, for evaluations of a polynomial and for arbitrary values;
    ENTER and EXIT calls interleave on the same arenas (shared epoch counter)."""
    import threading
    import torch
    from ecfft_b200.dist import PeerArena, enter_sharded_peer, exit_sharded_peer
    gpu, cpu = trees
    n = 1 << 14
    coeffs = oracle_mod.random_elements(n, seed=70 + world)
    evals = cpu.enter(coeffs)
    arbitrary = oracle_mod.random_elements(n, seed=80 + world)
    want_arb = cpu.exit(arbitrary)
    ed = torch.from_numpy(evals.view(np.int64)).cuda()
    ad = torch.from_numpy(arbitrary.view(np.int64)).cuda()
    cd = torch.from_numpy(coeffs.view(np.int64)).cuda()
    c = n // world
    arenas = PeerArena.local_group(n, world, device=0)
    results, errors = [None] * world, []
    torch.cuda.synchronize()

    def run(rank):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                sl = slice(rank * c, (rank + 1) * c)
                back = exit_sharded_peer(gpu, ed[sl], n, arenas[rank], gather=False)
                again = enter_sharded_peer(gpu, cd[sl], n, arenas[rank], gather=False)      # ENTER on the same arenas
                arb = exit_sharded_peer(gpu, ad[sl], n, arenas[rank], gather=False)
                stream.synchronize()
                results[rank] = tuple(t.cpu().numpy().view(np.uint64) for t in (back, again, arb))
        except Exception as e:
            errors.append(repr(e))

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    for rank in range(world):
        sl = slice(rank * c, (rank + 1) * c)
        back, again, arb = results[rank]
        eq(back, coeffs[sl])
        eq(again, evals[sl])
        eq(arb, want_arb[sl])
    for a in arenas:
        a.close()
