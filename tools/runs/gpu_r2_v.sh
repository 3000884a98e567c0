#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_v_pytest.log
for round in 1 2 3; do
python tools/ab_variants.py exit 22 5 '' 'ECFFT_B200_NO_EXIT_CHAIN=1' 2>&1 | tee -a gpurun_out/r02_v_ab_exit_chain.txt
done
python tools/ab_variants.py exit 18 20 '' 'ECFFT_B200_NO_EXIT_CHAIN=1' 2>&1 | tee -a gpurun_out/r02_v_ab_exit_chain.txt
python - <<'PY' 2>&1 | tee gpurun_out/r02_v_m31_timings.txt
import time, numpy as np, torch, ecfft_b200
from ecfft_b200 import _lib
L = _lib.load()
for lg in (16, 20, 22, 24):
    n = 1 << lg
    t0 = time.time(); t = ecfft_b200.m31.build_fftree(n); torch.cuda.synchronize(); tb = time.time() - t0
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
    del t
PY
