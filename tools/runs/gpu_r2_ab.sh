#!/bin/bash
mkdir -p gpurun_out
for v in 'ECFFT_B200_M31_PDL=0' 'ECFFT_B200_M31_PDL=1' 'ECFFT_B200_M31_PDL=2' 'ECFFT_B200_M31_PDL=0,ECFFT_B200_M31_MATRIX=1'; do
env $(echo $v | tr ',' ' ') python - <<'PY' 2>&1 | tee -a gpurun_out/r02_ab_m31_pdl.txt
import os, numpy as np, torch, ecfft_b200
for lg in (16, 20, 22):
    n = 1 << lg
    t = ecfft_b200.m31.build_fftree(n)
    x = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
    out = []
    for op in ("enter", "exit"):
        fn = getattr(t, op)
        for _ in range(3): y = fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): y = fn(x)
        e1.record(); torch.cuda.synchronize()
        out.append(f"{op} {e0.elapsed_time(e1)/10:.3f} ms")
    print(f"m31 n=2^{lg}: {', '.join(out)}  [{os.environ.get('ECFFT_B200_M31_PDL')} matrix={os.environ.get('ECFFT_B200_M31_MATRIX','0')}]")
    del t
PY
done
