#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests/test_m31.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_aa_pytest_m31_sym.log
ECFFT_B200_M31_MATRIX=1 timeout 900 python -m pytest tests/test_m31.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_aa_pytest_m31_matrix.log
for v in '' 'ECFFT_B200_M31_MATRIX=1'; do
env $v python - <<'PY' 2>&1 | tee -a gpurun_out/r02_aa_m31_timings.txt
import os, hashlib, numpy as np, torch, ecfft_b200
for lg in (12, 16, 20, 22, 24):
    n = 1 << lg
    t = ecfft_b200.m31.build_fftree(n)
    torch.manual_seed(1)
    x = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
    for op in ("enter", "exit"):
        fn = getattr(t, op)
        for _ in range(3): y = fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): y = fn(x)
        e1.record(); torch.cuda.synchronize()
        print(f"m31 {op} n=2^{lg}: {e0.elapsed_time(e1)/10:.3f} ms sha1 {hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:10]} matrix={os.environ.get('ECFFT_B200_M31_MATRIX','0')}")
    assert torch.equal(t.exit(t.enter(x)), x)
    del t
PY
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_aa_pytest.log
