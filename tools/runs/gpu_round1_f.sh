#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python tools/parity_quick.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_f.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_f.json; python - <<PY
import json
d=json.load(open("gpurun_out/bench_f.json"))
print("fused-small ms/step", round(d["ms_per_step"],3), "extend ms", round(d["roofline"]["kernel_ms_per_step"],3), "combine", round(d["roofline"]["other_kernels"]["k_enter_combine"]["ms_per_step"],3), "launches", d["gpu_launches"])
PY
ECFFT_B200_NO_SMALL_FUSION=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_f_nofuse.json; python - <<PY
import json
d=json.load(open("gpurun_out/bench_f_nofuse.json"))
print("no-fusion ms/step", round(d["ms_per_step"],3), "extend ms", round(d["roofline"]["kernel_ms_per_step"],3), "combine", round(d["roofline"]["other_kernels"]["k_enter_combine"]["ms_per_step"],3), "launches", d["gpu_launches"])
PY
python tools/bench_configs.py 20 2>&1 | head -3 | cut -c1-200
