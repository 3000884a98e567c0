#!/bin/bash
# round 2, series n: m31 field on the GPU (tests, sanitizer on a small run, timings)
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests/test_m31.py -x -q 2>&1 | tail -15 | tee gpurun_out/r02_n_pytest_m31.log
timeout 600 compute-sanitizer --tool memcheck python - <<'PY' 2>&1 | tail -8 | tee gpurun_out/r02_n_m31_memcheck.log
import numpy as np, ecfft_b200
t = ecfft_b200.m31.build_fftree(1 << 14)
c = np.arange(1 << 14, dtype=np.uint32)
ev = t.enter(c)
assert (t.exit(ev) == c).all()
print("m31 small run ok", t.degree(ev))
PY
python - <<'PY' 2>&1 | tee gpurun_out/r02_n_m31_timings.txt
import time, numpy as np, torch, ecfft_b200
from ecfft_b200 import _lib
L = _lib.load()
for lg in (12, 16, 20, 22):
    n = 1 << lg
    t0 = time.time(); t = ecfft_b200.m31.build_fftree(n); tb = time.time() - t0
    x = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
    for op in ("enter", "exit"):
        fn = getattr(t, op)
        for _ in range(3): y = fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.ecfft_launch_count()
        e0.record()
        reps = 10
        for _ in range(reps): y = fn(x)
        e1.record(); torch.cuda.synchronize()
        print(f"m31 {op} n=2^{lg}: {e0.elapsed_time(e1)/reps:.3f} ms, {(L.ecfft_launch_count()-l0)/reps:.0f} launches, tree build {tb:.2f} s")
    assert torch.equal(t.exit(t.enter(x)), x)
PY
