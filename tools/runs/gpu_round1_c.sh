#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for v in 0 1 2 3 4; do
  echo "=== variant $v"
  ECFFT_B200_TILE_VARIANT=$v python tools/parity_quick.py 2>&1 | tail -1
  ECFFT_B200_TILE_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_c_v$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c_v$v.json"))
print("variant $v ms/step", round(d["ms_per_step"],3), "extend ms", round(d["roofline"]["kernel_ms_per_step"],3), "combine ms", round(d["roofline"]["other_kernels"]["k_enter_combine"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3))
PY
done
ECFFT_B200_TILE_VARIANT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_tile -s 45 -c 4 -o gpurun_out/prof_extend_c_v1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out | tail -5
