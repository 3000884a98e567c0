"""pinned-host <-> device copy bandwidth and the host-buffer ENTER's breakdown (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
import ecfft_b200
from ecfft_b200 import _lib
from oracle import oracle as O
import bench
print(bench.bind_to_gpu_numa_node(0))
n = 1 << 22
h = torch.from_numpy(O.random_elements(n, seed=1).view(np.int64)).pin_memory()
d = torch.empty_like(h, device="cuda")
o = torch.empty((n, 4), dtype=torch.int64).pin_memory()
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: o.copy_(d, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{name} 128 MiB: {dt*1e3:.3f} ms = {n*32/dt/1e9:.1f} GB/s")
tree = ecfft_b200.build_fftree(n, parts=ecfft_b200.PARTS_ENTER_ONLY)
L = _lib.load()
hp = h.numpy().view(np.uint64); op = o.numpy().view(np.uint64)
for _ in range(3): _lib.check(L.ecfft_enter(tree._h, hp.ctypes.data_as(ctypes.c_void_p), n, op.ctypes.data_as(ctypes.c_void_p)))
t0 = time.perf_counter()
for _ in range(10): _lib.check(L.ecfft_enter(tree._h, hp.ctypes.data_as(ctypes.c_void_p), n, op.ctypes.data_as(ctypes.c_void_p)))
print(f"ecfft_enter host path: {(time.perf_counter()-t0)/10*1e3:.3f} ms")
for _ in range(3): tree.enter(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): tree.enter(d)
torch.cuda.synchronize(); print(f"device path: {(time.perf_counter()-t0)/10*1e3:.3f} ms")
# split ENTER on device-resident data: n/8, n/8, n/4, n/2 chunks + merges (what the host path computes)
def split_enter(x):
    c0 = n // 8
    a0 = tree.enter_range(x[:c0], 1, c0); a1 = tree.enter_range(x[c0:2*c0], 1, c0)
    b = tree.enter_range(torch.cat([a0, a1]), c0, 2*c0)
    a2 = tree.enter_range(x[2*c0:4*c0], 1, 2*c0)
    cc = tree.enter_range(torch.cat([b, a2]), 2*c0, 4*c0)
    a3 = tree.enter_range(x[4*c0:], 1, 4*c0)
    return tree.enter_range(torch.cat([cc, a3]), 4*c0, n)
def halves(x):
    c0 = n // 2
    a0 = tree.enter_range(x[:c0], 1, c0); a1 = tree.enter_range(x[c0:], 1, c0)
    return tree.enter_range(torch.cat([a0, a1]), c0, n)
for name, fn in (("split 1/8,1/8,1/4,1/2", split_enter), ("two halves", halves)):
    for _ in range(3): fn(d)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): r = fn(d)
    torch.cuda.synchronize(); print(f"{name} (device-resident, incl. torch.cat copies): {(time.perf_counter()-t0)/10*1e3:.3f} ms", bool((r == tree.enter(d)).all()))
