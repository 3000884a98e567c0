"""Multi-GPU ENTER: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

ENTER's recursion (reference src/fftree.rs:143-161) splits the coefficient vector into contiguous
halves entered independently on the half-size subtree, so with G ranks the bottom log2(n/G) depths of
chunk g are an independent ENTER(n/G) on rank g.

Two schedules for the top log2(G) depths:

* `enter_sharded_allgather` — the single-exchange form: ONE all-gather of the evaluation chunks, then
  every rank runs the top depths on the whole vector (replicated work: 1.85x / 2.7x / 3.0x ideal
  speed-up at 2 / 4 / 8 ranks).
* `enter_sharded` (default) — the top depths stay sharded: a rank keeps a contiguous chunk of one
  vector; butterfly levels whose pairs straddle ranks exchange chunks pairwise (one send/recv per
  level), rank-local levels run the usual tile kernel on the chunk, and the combine exchanges
  half-chunks with the sibling vector's ranks.  Per-rank work is ~1/G of the total; one final
  all-gather returns the full evaluation vector on every rank (the reference's `Vec<F>` result).

The schedule talks to its peers through a small comm object so that tests can run it with virtual
ranks (threads) on one device; `TorchComm` is the real one.
"""
import torch
import torch.distributed as dist


class TorchComm:
    """pairwise exchanges and the final all-gather over a torch.distributed process group"""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def sendrecv(self, sends, recvs):
        """sends: [(tensor, dst)], recvs: [(rows, like_tensor, src)] -> received tensors in order"""
        ops, out = [], []
        for tensor, dst in sends:
            ops.append(dist.P2POp(dist.isend, tensor.contiguous(), dst, self.group))
        for rows, like, src in recvs:
            buf = torch.empty((rows, 4), dtype=like.dtype, device=like.device)
            out.append(buf)
            ops.append(dist.P2POp(dist.irecv, buf, src, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out

    def all_gather(self, tensor):
        out = torch.empty((tensor.shape[0] * self.world, 4), dtype=tensor.dtype, device=tensor.device)
        dist.all_gather_into_tensor(out, tensor.contiguous(), group=self.group)
        return out


def _check(n, world, chunk):
    if n % world or (n // world) & (n // world - 1) or world & (world - 1):
        raise ValueError("world_size and n / world_size must be powers of two")
    if chunk.shape[0] != n // world:
        raise ValueError("chunk must hold n / world_size coefficients")


def enter_sharded_allgather(tree, chunk, n, group=None, comm=None):
    """chunk: this rank's n/G coefficients ((n/G, 4) limb tensor, rank order = coefficient order).
    Returns the full evaluation vector (n, 4) on every rank."""
    comm = comm or TorchComm(group)
    _check(n, comm.world, chunk)
    local = tree.enter_range(chunk, 1, n // comm.world)
    if comm.world == 1:
        return local
    return tree.enter_range(comm.all_gather(local), n // comm.world, n)


def enter_sharded(tree, chunk, n, group=None, gather=True, comm=None):
    """Fully sharded ENTER.  Returns the full (n, 4) evaluation vector on every rank (gather=True) or
    this rank's chunk of it, positions [rank*n/G, (rank+1)*n/G)."""
    comm = comm or TorchComm(group)
    world, rank = comm.world, comm.rank
    _check(n, world, chunk)
    c = n // world
    log_c = c.bit_length() - 1
    A = tree.enter_range(chunk, 1, c)           # evaluations of coefficient chunk `rank` on the c-leaf subtree
    r = 1                                       # ranks per vector at the current depth
    m = 2 * c
    while m <= n:
        h = m // 2                              # vector length = r * c
        log_h = h.bit_length() - 1
        k = rank % r                            # this rank's chunk index inside its vector
        pos0 = k * c
        vidx = rank // r                        # vector index; even = u (low coefficients), odd = v
        block0 = (vidx // 2) * 2 * r            # first rank of the 2r ranks that own this block
        # ---- EXTEND -> S1 of the vector this rank holds a chunk of (normalised butterflies) ----
        W = tree.mg_prescale(m, pos0, A)
        for j in range(log_h - 1, log_c - 1, -1):           # decompose levels whose pairs straddle ranks
            bit = (k >> (j - log_c)) & 1
            peer = rank ^ (1 << (j - log_c))
            (P,) = comm.sendrecv([(W, peer)], [(c, W, peer)])
            W = tree.mg_cross(m, 0, j, bit, pos0 - (bit << j), W, P)
        W = tree.mg_local(m, W)                             # all levels with half-stride < c
        for j in range(log_c, log_h):                       # recombine levels that straddle ranks
            bit = (k >> (j - log_c)) & 1
            peer = rank ^ (1 << (j - log_c))
            (P,) = comm.sendrecv([(W, peer)], [(c, W, peer)])
            W = tree.mg_cross(m, 1, j, bit, pos0 - (bit << j), W, P)
        # ---- combine (src/fftree.rs:155-159): output rank k' of the block needs i in [k'c/2, (k'+1)c/2)
        # of u0,u1 (from u-rank k'//2) and v0,v1 (from v-rank k'//2)
        is_u = vidx % 2 == 0
        half = c // 2
        kp = rank - block0                                  # my output-rank index in the block
        usrc, vsrc = block0 + kp // 2, block0 + r + kp // 2
        parts = {}
        sends, recvs, names = [], [], []
        for hsel in (0, 1):                                 # my two halves feed output ranks 2k and 2k+1
            dst = block0 + 2 * k + hsel
            part = torch.cat([A[hsel * half:(hsel + 1) * half], W[hsel * half:(hsel + 1) * half]])
            if dst == rank:
                parts["u" if is_u else "v"] = part
            else:
                sends.append((part, dst))
        for name, src in (("u", usrc), ("v", vsrc)):
            if src != rank:
                recvs.append((2 * half, A, src))
                names.append(name)
        for name, buf in zip(names, comm.sendrecv(sends, recvs)):
            parts[name] = buf
        u, v = parts["u"], parts["v"]
        A = tree.mg_combine(m, kp * half, u[:half], v[:half], u[half:], v[half:])
        r *= 2
        m *= 2
    if not gather or world == 1:
        return A
    return comm.all_gather(A)
