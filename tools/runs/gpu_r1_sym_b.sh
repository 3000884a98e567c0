#!/bin/bash
# radix-4 register-stage variants of the symmetric tile kernel: parity, A/B, ncu captures
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for v in 40 41 42 44 50 51 52 54; do echo -n "variant $v "; ECFFT_B200_TILE_VARIANT=$v python tools/parity_quick.py 2>&1 | tail -1; done
for round in 1 2; do
for v in 7 40 41 42 44 50 51 52 54; do
  ECFFT_B200_TILE_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab.json
  python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('round $round variant $v', round(d['ms_per_step'],3),'ms; extend', round(d['roofline']['kernel_ms_per_step'],3), '; combine', round(d['roofline']['other_kernels']['k_enter_combine']['ms_per_step'],3), '; e2e ms', round(d['e2e']['ms_per_step'],2))"
done; done
ECFFT_B200_TILE_VARIANT=7 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_extend_tile -s 100 -c 12 -o gpurun_out/prof_sym_v7 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v7.log 2>&1
ECFFT_B200_TILE_VARIANT=50 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_extend_tile -s 100 -c 12 -o gpurun_out/prof_sym_v50 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v50.log 2>&1
ls -la gpurun_out | tail -5
