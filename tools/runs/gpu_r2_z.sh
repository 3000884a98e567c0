#!/bin/bash
# PDL on the generic map kernels and the m31 kernels: tests + A/B (ECFFT_B200_PDL=0 switches every PDL launch off)
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_z_pytest.log
python tools/ab_variants.py exit 22 5 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_z_ab_pdl.txt
python tools/ab_variants.py exit 12 50 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_z_ab_pdl.txt
python tools/ab_variants.py enter 12 100 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_z_ab_pdl.txt
python tools/bench_configs.py 20 2>&1 | cut -c1-120 | tee gpurun_out/r02_z_configs20.txt
python - <<'PY' 2>&1 | tee gpurun_out/r02_z_m31_timings.txt
import time, numpy as np, torch, ecfft_b200
from ecfft_b200 import _lib
L = _lib.load()
for lg in (12, 16, 20, 22):
    n = 1 << lg
    t = ecfft_b200.m31.build_fftree(n)
    x = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
    for op in ("enter", "exit"):
        fn = getattr(t, op)
        for _ in range(3): y = fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): y = fn(x)
        e1.record(); torch.cuda.synchronize()
        print(f"m31 {op} n=2^{lg}: {e0.elapsed_time(e1)/10:.3f} ms")
    assert torch.equal(t.exit(t.enter(x)), x)
PY
