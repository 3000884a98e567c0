#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for s in 1 2 4; do echo -n "streams $s "; ECFFT_B200_ENTER_STREAMS=$s python tools/parity_quick.py 2>&1 | tail -1; done
for round in 1 2; do for s in 1 2 4; do
ECFFT_B200_ENTER_STREAMS=$s python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab.json; python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('streams $s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])"; done; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
