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
    if world > 1 and n // world < 2:
        raise ValueError("the sharded schedules need at least 2 coefficients per rank")
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


# ----------------------------------------------------------------------------------------------------
# Peer-memory schedule: the same sharded top depths, but no send/recv — every rank keeps the buffers of
# the current ENTER in an arena the other ranks of the node map through CUDA IPC, and the butterfly /
# combine kernels load the partner's operands over NVLink themselves (`ecfft_mg_cross_dev`'s partner
# pointer and `ecfft_mg_combine_dev`'s u/v pointers are peer pointers).  Ordering is by stream-ordered
# u64 flags in the arenas (`ecfft_mg_signal_dev` / `ecfft_mg_wait_dev`); every produced buffer of one
# ENTER has its own slot, so within a call there is nothing to protect against overwriting; a call starts
# by waiting for every peer's "finished the previous call" flag and ends by publishing its own.
# ----------------------------------------------------------------------------------------------------
import ctypes

from . import _lib

_FLAG_BYTES = 4096      # 512 flags: one per synchronisation step of a call
_WAIT_MS = 20000


def _peer_slots(world):
    """buffers one call produces per rank: A0, then per top depth W_pre, one per cross level, W_local, A_next"""
    slots, r = 1, 1
    while r < world:
        slots += 2 + 2 * (r.bit_length() - 1) + 1
        r *= 2
    return slots


class PeerArena:
    """Per-rank arena [flags | slots] plus the peers' mappings.  `PeerArena.create` is collective over the
    process group; `PeerArena.local_group` builds the arenas of several virtual ranks inside one process
    (tests: threads on one GPU share plain device pointers)."""

    def __init__(self, rank, world, n, device, own_ptr, bases, owner=True, opened=()):
        self.rank, self.world, self.n, self.device = rank, world, n, device
        self.c = n // world
        self.own = own_ptr
        self.bases = bases            # bases[r]: rank r's arena as seen from this rank
        self.epoch = 0
        self._owner, self._opened = owner, list(opened)

    @staticmethod
    def nbytes(n, world):
        L = _lib.load()
        size, size_exit = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(L.ecfft_mg_arena_bytes(n, world, ctypes.byref(size)))            # what ecfft_enter_peer_dev needs
        _lib.check(L.ecfft_mg_exit_arena_bytes(n, world, ctypes.byref(size_exit)))  # ... and ecfft_exit_peer_dev
        return max(size.value, size_exit.value, _FLAG_BYTES + _peer_slots(world) * (n // world) * 32)

    @classmethod
    def create(cls, n, device, group=None):
        L = _lib.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        _lib.check(L.ecfft_mg_arena_alloc(device, cls.nbytes(n, world), ctypes.byref(own), handle))
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw, group=group)
        bases, opened = [], []
        for r in range(world):
            if r == rank:
                bases.append(own.value)
                continue
            p = ctypes.c_void_p()
            _lib.check(L.ecfft_mg_arena_open(device, handles[r], ctypes.byref(p)))
            bases.append(p.value)
            opened.append(p.value)
        dist.barrier(group=group)
        return cls(rank, world, n, device, own.value, bases, owner=True, opened=opened)

    @classmethod
    def local_group(cls, n, world, device):
        L = _lib.load()
        ptrs = []
        for _ in range(world):
            own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            _lib.check(L.ecfft_mg_arena_alloc(device, cls.nbytes(n, world), ctypes.byref(own), handle))
            ptrs.append(own.value)
        return [cls(r, world, n, device, ptrs[r], list(ptrs), owner=True) for r in range(world)]

    def slot(self, r, idx, elem_off=0):
        return self.bases[r] + _FLAG_BYTES + (idx * self.c + elem_off) * 32

    def flag(self, r, sid):
        return self.bases[r] + 8 * sid

    def status(self):
        """0, or the record of the first flag wait on this rank that timed out (include/ecfft_b200.h)"""
        s = ctypes.c_ulonglong()
        _lib.check(_lib.load().ecfft_mg_arena_status(ctypes.c_void_p(self.own), ctypes.byref(s)))
        return s.value

    def close(self):
        L = _lib.load()
        for p in self._opened:
            L.ecfft_mg_arena_close(ctypes.c_void_p(p))
        self._opened = []
        if self._owner and self.own:
            L.ecfft_mg_arena_free(ctypes.c_void_p(self.own))
            self.own = None


def _finish_peer(out_t, n, world, group, gather, all_gather):
    """calls are ordered against each other on the device (done flags), so gather=False needs no barrier"""
    if gather and world > 1:
        if all_gather is not None:
            return all_gather(out_t)
        full = torch.empty((n, 4), dtype=out_t.dtype, device=out_t.device)
        dist.all_gather_into_tensor(full, out_t, group=group)
        return full
    return out_t


def enter_sharded_peer(tree, chunk, n, arena, group=None, gather=True, all_gather=None, native=True):
    """Fully sharded ENTER with peer-memory exchange.  chunk: this rank's n/G coefficients (CUDA tensor).
    Returns the full (n, 4) evaluation vector on every rank (gather=True) or this rank's chunk of it.
    `all_gather` defaults to the process group's (tests with virtual ranks pass their own).
    native=True runs the whole per-rank schedule inside the library (`ecfft_enter_peer_dev`,
    csrc/sharded.cu); native=False drives the same building blocks step by step from here."""
    L = _lib.load()
    world, rank = arena.world, arena.rank
    _check(n, world, chunk)
    if arena.n != n:
        raise ValueError("arena was created for a different n")
    if native:
        chunk = chunk.contiguous()
        arena.epoch += 1
        out_t = torch.empty((n // world, 4), dtype=chunk.dtype, device=chunk.device)
        bases = (ctypes.c_void_p * world)(*arena.bases)
        _lib.check(L.ecfft_enter_peer_dev(tree._h, ctypes.c_void_p(chunk.data_ptr()), n, rank, world, bases, arena.epoch,
                                          ctypes.c_void_p(out_t.data_ptr()),
                                          ctypes.c_void_p(torch.cuda.current_stream(chunk.device).cuda_stream)))
        return _finish_peer(out_t, n, world, group, gather, all_gather)
    c = n // world
    log_c = c.bit_length() - 1
    st = ctypes.c_void_p(torch.cuda.current_stream(chunk.device).cuda_stream)
    h = tree._h
    vp = ctypes.c_void_p
    arena.epoch += 1
    epoch = arena.epoch
    state = {"slot": 0, "sid": 0}

    def new_slot():
        state["slot"] += 1
        return state["slot"] - 1

    def sync(peers):
        """everything this rank has enqueued is published; then wait for the same step of `peers`"""
        sid = state["sid"]
        state["sid"] += 1
        if sid >= _FLAG_BYTES // 8 - 1:
            raise RuntimeError("peer arena: too many synchronisation steps")
        _lib.check(L.ecfft_mg_signal_dev(vp(arena.flag(rank, sid)), epoch, st))
        for p in peers:
            if p != rank:
                _lib.check(L.ecfft_mg_wait_dev(vp(arena.flag(p, sid)), epoch, _WAIT_MS, st))

    chunk = chunk.contiguous()
    done = _FLAG_BYTES // 8 - 1                                 # last flag: "finished call <epoch>"
    if epoch > 1:                                               # peers may still read this arena from the previous call
        for p in range(world):
            if p != rank:
                _lib.check(L.ecfft_mg_wait_dev(vp(arena.flag(p, done)), epoch - 1, _WAIT_MS, st))
    sA = new_slot()
    _lib.check(L.ecfft_enter_range_dev(h, vp(chunk.data_ptr()), c, 1, c, vp(arena.slot(rank, sA)), st))
    out_t = None
    r, m = 1, 2 * c
    while m <= n:
        hlen = m // 2
        log_h = hlen.bit_length() - 1
        k = rank % r
        pos0 = k * c
        vidx = rank // r
        block0 = (vidx // 2) * 2 * r
        # ---- EXTEND -> S1 of the vector this rank holds a chunk of
        sW = new_slot()
        _lib.check(L.ecfft_mg_prescale_dev(h, m, pos0, vp(arena.slot(rank, sA)), c, vp(arena.slot(rank, sW)), st))
        for phase, levels in ((0, range(log_h - 1, log_c - 1, -1)), (None, None), (1, range(log_c, log_h))):
            if phase is None:                                   # all levels with half-stride < c: rank-local
                sN = new_slot()
                _lib.check(L.ecfft_mg_local_dev(h, m, vp(arena.slot(rank, sW)), c, vp(arena.slot(rank, sN)), st))
                sW = sN
                continue
            for j in levels:                                    # levels whose pairs straddle two ranks
                bit = (k >> (j - log_c)) & 1
                peer = rank ^ (1 << (j - log_c))
                sync([peer])
                sN = new_slot()
                _lib.check(L.ecfft_mg_cross_dev(h, m, phase, j, bit, pos0 - (bit << j), vp(arena.slot(rank, sW)),
                                                vp(arena.slot(peer, sW)), c, vp(arena.slot(rank, sN)), st))
                sW = sN
        # ---- combine (src/fftree.rs:155-159): output rank kp of the block takes i in [kp c/2, (kp+1) c/2)
        # straight out of the u-rank's and the v-rank's A and W
        half = c // 2
        kp = rank - block0
        usrc, vsrc = block0 + kp // 2, block0 + r + kp // 2
        off = (kp % 2) * half
        sync([usrc, vsrc])
        last = 2 * m > n
        if last and not gather:
            out_t = torch.empty((c, 4), dtype=chunk.dtype, device=chunk.device)
            dst = out_t.data_ptr()
            sNext = None
        elif last:
            out_t = torch.empty((c, 4), dtype=chunk.dtype, device=chunk.device)
            dst = out_t.data_ptr()
            sNext = None
        else:
            sNext = new_slot()
            dst = arena.slot(rank, sNext)
        _lib.check(L.ecfft_mg_combine_dev(h, m, kp * half, vp(arena.slot(usrc, sA, off)), vp(arena.slot(vsrc, sA, off)),
                                          vp(arena.slot(usrc, sW, off)), vp(arena.slot(vsrc, sW, off)), half, vp(dst), st))
        sA = sNext
        r *= 2
        m *= 2
    _lib.check(L.ecfft_mg_signal_dev(vp(arena.flag(rank, done)), epoch, st))
    if out_t is None:                                           # world == 1
        out_t = torch.empty((c, 4), dtype=chunk.dtype, device=chunk.device)
        _lib.check(L.ecfft_enter_range_dev(h, vp(arena.slot(rank, 0)), c, c, c, vp(out_t.data_ptr()), st))
    return _finish_peer(out_t, n, world, group, gather, all_gather)


def exit_sharded_peer(tree, chunk, n, arena, group=None, gather=True, all_gather=None):
    """Fully sharded EXIT with peer-memory exchange (reference src/fftree.rs:200-224; csrc/sharded.cu exit_peer).
    chunk: this rank's n/G evaluations (CUDA tensor, rank order = leaf order).  Returns the full (n, 4) coefficient
    vector on every rank (gather=True) or this rank's chunk of it.  Shares the arena (and its epoch counter) with
    `enter_sharded_peer`; the tree must carry all tables (PARTS_FULL)."""
    L = _lib.load()
    world, rank = arena.world, arena.rank
    _check(n, world, chunk)
    if arena.n != n:
        raise ValueError("arena was created for a different n")
    if world > 1 and n // world < 4:
        raise ValueError("the sharded EXIT needs at least 4 evaluations per rank")
    chunk = chunk.contiguous()
    arena.epoch += 1
    out_t = torch.empty((n // world, 4), dtype=chunk.dtype, device=chunk.device)
    bases = (ctypes.c_void_p * world)(*arena.bases)
    _lib.check(L.ecfft_exit_peer_dev(tree._h, ctypes.c_void_p(chunk.data_ptr()), n, rank, world, bases, arena.epoch,
                                     ctypes.c_void_p(out_t.data_ptr()),
                                     ctypes.c_void_p(torch.cuda.current_stream(chunk.device).cuda_stream)))
    return _finish_peer(out_t, n, world, group, gather, all_gather)
