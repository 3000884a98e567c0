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
"""
import torch
import torch.distributed as dist


def _check(n, world, chunk):
    if n % world or (n // world) & (n // world - 1) or world & (world - 1):
        raise ValueError("world_size and n / world_size must be powers of two")
    if chunk.shape[0] != n // world:
        raise ValueError("chunk must hold n / world_size coefficients")


def enter_sharded_allgather(tree, chunk, n, group=None):
    """chunk: this rank's n/G coefficients ((n/G, 4) limb tensor, rank order = coefficient order).
    Returns the full evaluation vector (n, 4) on every rank."""
    world = dist.get_world_size(group)
    _check(n, world, chunk)
    local = tree.enter_range(chunk, 1, n // world)
    if world == 1:
        return local
    gathered = torch.empty((n, 4), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    return tree.enter_range(gathered, n // world, n)


def _exchange(send, peer, group):
    """pairwise exchange of equal-sized tensors with `peer`"""
    recv = torch.empty_like(send)
    ops = [dist.P2POp(dist.isend, send, peer, group), dist.P2POp(dist.irecv, recv, peer, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return recv


def enter_sharded(tree, chunk, n, group=None, gather=True):
    """Fully sharded ENTER.  Returns the full (n, 4) evaluation vector on every rank (gather=True) or
    this rank's chunk of it, positions [rank*n/G, (rank+1)*n/G)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
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
            P = _exchange(W, peer, group)
            W = tree.mg_cross(m, 0, j, bit, pos0 - (bit << j), W, P)
        W = tree.mg_local(m, W)                             # all levels with half-stride < c
        for j in range(log_c, log_h):                       # recombine levels that straddle ranks
            bit = (k >> (j - log_c)) & 1
            peer = rank ^ (1 << (j - log_c))
            P = _exchange(W, peer, group)
            W = tree.mg_cross(m, 1, j, bit, pos0 - (bit << j), W, P)
        # ---- combine (src/fftree.rs:155-159): output rank k' of the block needs i in [k'c/2, (k'+1)c/2)
        # of u0,u1 (from u-rank k'//2) and v0,v1 (from v-rank k'//2)
        is_u = vidx % 2 == 0
        half = c // 2
        mine = torch.cat([A, W])                            # [x0 | x1] of my vector chunk
        dests = [block0 + 2 * k, block0 + 2 * k + 1]        # output ranks fed by my two halves
        kp = rank - block0                                  # my output-rank index in the block
        usrc, vsrc = block0 + kp // 2, block0 + r + kp // 2
        want = kp % 2                                       # which half of the sources' chunks I need
        ops, bufs = [], {}
        for hsel, dst in enumerate(dests):
            part = torch.cat([mine[hsel * half:(hsel + 1) * half], mine[c + hsel * half:c + (hsel + 1) * half]])
            if dst == rank:
                bufs["u" if is_u else "v"] = part
            else:
                ops.append(dist.P2POp(dist.isend, part, dst, group))
        for name, src in (("u", usrc), ("v", vsrc)):
            if src != rank:
                bufs[name] = torch.empty((2 * half, 4), dtype=A.dtype, device=A.device)
                ops.append(dist.P2POp(dist.irecv, bufs[name], src, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        u, v = bufs["u"], bufs["v"]
        A = tree.mg_combine(m, kp * half, u[:half], v[:half], u[half:], v[half:])
        del want
        r *= 2
        m *= 2
    if not gather or world == 1:
        return A
    out = torch.empty((n, 4), dtype=A.dtype, device=A.device)
    dist.all_gather_into_tensor(out, A, group=group)
    return out
