#!/bin/bash
# launch-shape variants of the default library (sum-form + chained reduce), same box, two rounds
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in 1 2 5 6 7 8; do ECFFT_B200_TILE_VARIANT=$v python tools/parity_quick.py 2>&1 | tail -1; done
for round in 1 2; do
for v in 1 2 5 6 7 8; do
  ECFFT_B200_TILE_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab2.json
  python -c "
import json; d=json.load(open('gpurun_out/ab2.json')); print('round $round variant $v', round(d['ms_per_step'],3),'ms; extend', round(d['roofline']['kernel_ms_per_step'],3), '; e2e ms', round(d['e2e']['ms_per_step'],2))"
done; done
