#!/bin/bash
# first run of the symmetric (one-product) butterflies: parity, launch-shape A/B, bench
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python tools/parity_quick.py 2>&1 | tail -3
for v in 1 2 11 12 17; do echo "variant $v"; ECFFT_B200_TILE_VARIANT=$v python tools/parity_quick.py 2>&1 | tail -1; done
for round in 1 2; do
for v in 7 17 1 11 2 12; do
  ECFFT_B200_TILE_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab.json
  python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('round $round sym variant $v', round(d['ms_per_step'],3),'ms; extend', round(d['roofline']['kernel_ms_per_step'],3), '; combine', round(d['roofline']['other_kernels']['k_enter_combine']['ms_per_step'],3), '; e2e ms', round(d['e2e']['ms_per_step'],2))"
done; done
ECFFT_B200_BUTTERFLY=normalised python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab.json
python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('normalised variant 7', round(d['ms_per_step'],3),'ms; extend', round(d['roofline']['kernel_ms_per_step'],3))"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_sym_a.json 2> gpurun_out/bench_sym_a.err; cat gpurun_out/bench_sym_a.json
