"""Host-side logic of the multi-GPU ENTER (ecfft_b200/dist.py) on CPU: world_size 2 and 4 over gloo.
Each rank enters its coefficient chunk, one all-gather exchanges the evaluation chunks, the top
recursion depths finish on the gathered vector.  The per-rank compute is played by the CPU oracle's
level-range restatement, so what is tested is the sharding/exchange schedule itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


P = 2**256 - 2**32 - 977


def _ints(t):
    return [r[0] | (r[1] << 64) | (r[2] << 128) | (r[3] << 192) for r in t.numpy().view(np.uint64).tolist()]


def _tensor(vals):
    a = np.array([[(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)] for v in vals], dtype=np.uint64)
    return torch.from_numpy(a.reshape(-1, 4).view(np.int64))


class OracleBackedTree:
    """Stands in for ecfft_b200.FFTree on CPU tensors: enter_range through the oracle's level-range
    restatement, the multi-GPU building blocks as Python big-integer restatements of the normalised
    butterflies with constants derived from the oracle's f and recombine tables."""

    def __init__(self, n):
        from oracle import oracle as O
        self.O = O
        self.t = O.OracleTree.build(n, parts=1)
        self._cache = {}

    def enter_range(self, data, m_lo, m_hi):
        arr = data.numpy().view(np.uint64)
        return torch.from_numpy(self.t.enter_range(arr, m_lo, m_hi).view(np.int64))

    def _level(self, m):
        if m in self._cache:
            return self._cache[m]
        O = self.O
        st = self.t.subtree_with_size(m)
        f = O.from_mont(st.table("f"))
        rm = O.from_mont(st.table("recombine"))
        xnn = O.from_mont(st.table("xnn_s"))
        h = m // 2
        tw = {}
        for mu in (0, 1):
            j = 0
            while (1 << j) < h:
                B = 2 << j
                for i in range(1 << j):
                    s0, s1 = f[2 * B + 2 * i + mu], f[2 * B + 2 * i + mu + B]
                    tw[(mu, j, i)] = (s0, s1, pow(s1 - s0, -1, P))
                j += 1
        gam = {}
        for mu in (0, 1):
            g = []
            for p in range(h):
                acc, j = 1, 0
                while (1 << j) < h:
                    i, b = p & ((1 << j) - 1), (p >> j) & 1
                    acc = acc * rm[4 * ((2 << j) + 2 * i + mu) + 2 * b] % P
                    j += 1
                g.append(acc)
            gam[mu] = g
        self._cache[m] = (tw, gam, xnn)
        return self._cache[m]

    def mg_prescale(self, m, pos0, x):
        _, gam, _ = self._level(m)
        return _tensor([v * pow(gam[0][pos0 + e], -1, P) % P for e, v in enumerate(_ints(x))])

    def mg_cross(self, m, phase, j, role, p_pos0, own, partner):
        tw, _, _ = self._level(m)
        out = []
        for e, (xo, xr) in enumerate(zip(_ints(own), _ints(partner))):
            xp, xq = (xo, xr) if role == 0 else (xr, xo)
            i = (p_pos0 + e) & ((1 << j) - 1)
            if phase == 0:
                s0, _, c = tw[(0, j, i)]
                yq = c * (xq - xp) % P
                out.append(yq if role == 1 else (xp - s0 * yq) % P)
            else:
                s0, s1, _ = tw[(1, j, i)]
                out.append((xp + (s0 if role == 0 else s1) * xq) % P)
        return _tensor(out)

    def mg_local(self, m, x):
        tw, _, _ = self._level(m)
        v = _ints(x)
        L = len(v).bit_length() - 1
        for j in range(L - 1, -1, -1):
            for p in range(len(v)):
                if not (p >> j) & 1:
                    s0, _, c = tw[(0, j, p & ((1 << j) - 1))]
                    q = p + (1 << j)
                    yq = c * (v[q] - v[p]) % P
                    v[p], v[q] = (v[p] - s0 * yq) % P, yq
        for j in range(L):
            for p in range(len(v)):
                if not (p >> j) & 1:
                    s0, s1, _ = tw[(1, j, p & ((1 << j) - 1))]
                    q = p + (1 << j)
                    v[p], v[q] = (v[p] + s0 * v[q]) % P, (v[p] + s1 * v[q]) % P
        return _tensor(v)

    def mg_combine(self, m, i0, u0, v0, u1, v1):
        _, gam, xnn = self._level(m)
        out = []
        for t, (a, b, c, d) in enumerate(zip(_ints(u0), _ints(v0), _ints(u1), _ints(v1))):
            i = i0 + t
            out.append((a + b * xnn[2 * i]) % P)
            out.append(gam[1][i] * (c + d * xnn[2 * i + 1]) % P)
        return _tensor(out)


def _worker(rank, world, port, n, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from ecfft_b200.dist import enter_sharded, enter_sharded_allgather
        tree = OracleBackedTree(n)
        x = O.random_elements(n, seed=11)
        c = n // world
        chunk = torch.from_numpy(x[rank * c:(rank + 1) * c].view(np.int64).copy())
        want = tree.t.enter(x)
        ok = (enter_sharded_allgather(tree, chunk, n).numpy().view(np.uint64) == want).all()
        ok = ok and (enter_sharded(tree, chunk, n).numpy().view(np.uint64) == want).all()
        part = enter_sharded(tree, chunk, n, gather=False).numpy().view(np.uint64)
        ok = ok and (part == want[rank * c:(rank + 1) * c]).all()
        np.save(os.path.join(result_dir, f"ok_{rank}.npy"), np.array([ok]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_enter_sharded_matches_single_enter(world, tmp_path):
    n = 128
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"ok_{r}.npy")[0]


def test_enter_sharded_rejects_bad_chunking():
    from ecfft_b200.dist import enter_sharded

    class FakeGroup:
        pass

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        tree = OracleBackedTree(16)
        with pytest.raises(ValueError):
            enter_sharded(tree, torch.zeros((3, 4), dtype=torch.int64), 16)
        x = torch.from_numpy(__import__("oracle.oracle", fromlist=["x"]).random_elements(16).view(np.int64))
        out = enter_sharded(tree, x, 16)
        assert out.shape == (16, 4)
    finally:
        dist.destroy_process_group()


# ---- sharded EXIT (csrc/sharded.cu exit_peer): the per-rank bookkeeping, played on CPU over gloo -------------------
def _exit_worker(rank, world, port, n, result_dir):
    """Each rank holds evaluations [rank c, (rank+1) c).  Top log2(world) depths: MOD = REDC, x c, REDC
    (reference src/fftree.rs:232-259, 277-281) on vectors spread over r ranks — every pointwise step works on the
    rank's chunk with table offsets, the two EXTENDs of a REDC are EXTENDs of the distributed half-length vector
    (played here by gathering the group's chunks and calling the oracle) — then the split u0 | v0 (src/fftree.rs:
    206-220) moves half-chunks to their new owners exactly as exit_peer does; below, a local EXIT(c)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        tree = O.OracleTree.build(n)
        x = O.random_elements(n, seed=17)
        want = tree.exit(x)
        c = n // world
        cc = c // 2
        cur = _ints(torch.from_numpy(x[rank * c:(rank + 1) * c].view(np.int64).copy()))   # Montgomery values as ints
        R = 2**256 % P
        rinv = pow(R, -1, P)

        def mmul(a, b):   # product of two Montgomery-form values, Montgomery form (ark-ff `*`)
            return a * b * rinv % P

        def gather_group(vals, group0, r):
            """the r ranks group0 .. group0 + r - 1 each hold `vals`; returns their concatenation"""
            bufs = [torch.zeros((len(vals), 4), dtype=torch.int64) for _ in range(world)]
            dist.all_gather(bufs, _tensor(vals))
            out = []
            for g in range(group0, group0 + r):
                out += _ints(bufs[g])
            return out

        m, r = n, world
        while r > 1:
            st = tree.subtree_with_size(m)
            tab = {name: _ints(torch.from_numpy(st.table(name).view(np.int64))) for name in
                   ("xnn_s", "xnn_s_inv", "z0_inv_s1", "z0z0_rem_xnn_s")}
            k, group0 = rank % r, rank - rank % r
            pos0 = k * cc

            def dist_extend(vals, moiety):
                full = gather_group(vals, group0, r)
                arr = _tensor(full).numpy().view(np.uint64)
                return _ints(torch.from_numpy(st.extend(arr, moiety).view(np.int64)))[pos0:pos0 + cc]

            def redc(ev):   # fftree.rs:232-259 with a = xnn_s on this rank's chunk
                e0, e1 = ev[0::2], ev[1::2]
                a = tab["xnn_s"]
                t0 = [mmul(e0[i], tab["xnn_s_inv"][2 * (pos0 + i)]) for i in range(cc)]
                g1 = dist_extend(t0, 1)
                h1 = [mmul((e1[i] - mmul(g1[i], a[2 * (pos0 + i) + 1])) % P, tab["z0_inv_s1"][pos0 + i]) for i in range(cc)]
                h0 = dist_extend(h1, 0)
                out = [0] * c
                out[0::2], out[1::2] = h0, h1
                return out

            hb = redc(cur)
            hb = [mmul(hb[j], tab["z0z0_rem_xnn_s"][k * c + j]) for j in range(c)]
            M = redc(hb)
            u0 = M[0::2]
            v0 = [mmul((cur[2 * i] - u0[i]) % P, tab["xnn_s_inv"][2 * (pos0 + i)]) for i in range(cc)]
            # exchange: rank k' of a half takes its two pieces from ranks 2k', 2k'+1 of the parent group
            r2 = r // 2
            newk, second = k % r2, k >= r2
            src_a, src_b = group0 + 2 * newk, group0 + 2 * newk + 1
            bufs_u = [torch.zeros((cc, 4), dtype=torch.int64) for _ in range(world)]
            bufs_v = [torch.zeros((cc, 4), dtype=torch.int64) for _ in range(world)]
            dist.all_gather(bufs_u, _tensor(u0))
            dist.all_gather(bufs_v, _tensor(v0))
            part = bufs_v if second else bufs_u
            cur = _ints(part[src_a]) + _ints(part[src_b])
            m //= 2
            r = r2
        local = tree.subtree_with_size(c).exit(_tensor(cur).numpy().view(np.uint64))
        ok = (local == want[rank * c:(rank + 1) * c]).all()
        np.save(os.path.join(result_dir, f"exit_ok_{rank}.npy"), np.array([ok]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_exit_sharded_bookkeeping_matches_single_exit(world, tmp_path):
    n = 64
    mp.spawn(_exit_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"exit_ok_{r}.npy")[0]
