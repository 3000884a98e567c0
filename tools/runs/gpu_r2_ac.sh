#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_ac_pytest.log
for v in '' 'ECFFT_B200_PDL=0'; do
env $v python - <<'PY' 2>&1 | tee -a gpurun_out/r02_ac_glue_pdl.txt
import os, numpy as np, torch, ecfft_b200
from oracle import oracle as O
t = ecfft_b200.build_fftree(1 << 21)
def tm(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for lg in (12, 16, 20):
    x = torch.from_numpy(O.random_elements(1 << lg, seed=5).view(np.int64)).cuda()
    a = torch.from_numpy(t.table("xnn_s", 1 << lg).view(np.int64)).cuda()
    c = torch.from_numpy(t.table("z0z0_rem_xnn_s", 1 << lg).view(np.int64)).cuda()
    ev = t.enter(x)
    print(f"n=2^{lg}: VANISH {tm(lambda: t.vanish(x)):.3f} DEGREE {tm(lambda: t.degree(ev)):.3f} REDC {tm(lambda: t.redc_z0(x, a)):.3f} MOD {tm(lambda: t.modular_reduce(x, a, c)):.3f} MEXTEND {tm(lambda: t.mextend(x, 1)):.3f} roundtrip {tm(lambda: t.exit(t.enter(x))):.3f} ms  [PDL={os.environ.get('ECFFT_B200_PDL','1')}]")
PY
done
python - <<'PY' 2>&1 | tee gpurun_out/r02_ac_m31_timings.txt
import numpy as np, torch, ecfft_b200
for lg in (12, 16, 20, 22, 24):
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
    assert torch.equal(t.exit(t.enter(x)), x)
    print(f"m31 n=2^{lg}: {', '.join(out)}")
    del t
PY
