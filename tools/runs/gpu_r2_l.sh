#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for lg in 11 13 14 15; do python tools/variant_diff.py extend $lg '' 'ECFFT_B200_TMA=1' 2>&1 | tee -a gpurun_out/r02_l_tma_diff.txt; done
for lg in 10 11 12 13 14; do python tools/variant_diff.py enter $lg '' 'ECFFT_B200_TMA=1' 2>&1 | tee -a gpurun_out/r02_l_tma_diff.txt; done
