#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
ECFFT_B200_TILE_VARIANT=51 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_extend_tile -s 104 -c 6 -o gpurun_out/prof_sym_v51 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v51.log 2>&1
ECFFT_B200_TILE_VARIANT=7 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_extend_tile -s 106 -c 3 -o gpurun_out/prof_sym_v7 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v7.log 2>&1
ls -la gpurun_out | tail -5
