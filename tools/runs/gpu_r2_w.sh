#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_w_pytest.log
for v in '' 'ECFFT_B200_NO_VANISH_FUSION=1'; do
env $v python - <<'PY' 2>&1 | tee -a gpurun_out/r02_w_vanish.txt
import os, hashlib, numpy as np, torch, ecfft_b200
from oracle import oracle as O
t = ecfft_b200.build_fftree(1 << 21)
for lg in (12, 16, 20):
    x = torch.from_numpy(O.random_elements(1 << lg, seed=5).view(np.int64)).cuda()
    for _ in range(3): y = t.vanish(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = t.vanish(x)
    e1.record(); torch.cuda.synchronize()
    print(f"VANISH n=2^{lg}: {e0.elapsed_time(e1)/10:.3f} ms sha1 {hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:12]} fusion_off={os.environ.get('ECFFT_B200_NO_VANISH_FUSION','0')}")
PY
done
