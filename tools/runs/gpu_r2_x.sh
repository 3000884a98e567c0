#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_x_pytest.log
for v in '' 'ECFFT_B200_DEGREE_HOST_BRANCH=1'; do
env $v python - <<'PY' 2>&1 | tee -a gpurun_out/r02_x_degree.txt
import os, numpy as np, torch, ecfft_b200
from oracle import oracle as O
t = ecfft_b200.build_fftree(1 << 20)
for lg, deg in ((12, 1000), (16, 40000), (20, 700001), (20, 5)):
    c = O.random_elements(1 << lg, seed=5)
    c[deg + 1:] = 0
    ev = t.enter(torch.from_numpy(c.view(np.int64)).cuda())
    for _ in range(3): d = t.degree(ev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d = t.degree(ev)
    e1.record(); torch.cuda.synchronize()
    print(f"DEGREE n=2^{lg} (degree {deg}): {e0.elapsed_time(e1)/10:.3f} ms -> {d} {'ok' if d == deg else 'WRONG'} host_branch={os.environ.get('ECFFT_B200_DEGREE_HOST_BRANCH','0')}")
PY
done
