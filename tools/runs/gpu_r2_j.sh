#!/bin/bash
# round 2, series j: TMA tile load and twiddle prefetch A/B (same box), DFMA product microbenchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
make -C oracle >/dev/null 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu && ./tools/microbench | grep -E "device|bps=(1|4|8)" > gpurun_out/r02_j_microbench_dfma.txt
tail -12 gpurun_out/r02_j_microbench_dfma.txt
for round in 1 2; do
python tools/ab_variants.py enter 22 20 '' 'ECFFT_B200_TMA=1' 'ECFFT_B200_TW_PREFETCH=1' 'ECFFT_B200_TW_PREFETCH=2' 'ECFFT_B200_TMA=1,ECFFT_B200_TW_PREFETCH=1' 'ECFFT_B200_SYM_VARIANT=0' 'ECFFT_B200_SYM_VARIANT=0,ECFFT_B200_TW_PREFETCH=1' 2>&1 | tee -a gpurun_out/r02_j_ab_enter22.txt
done
python tools/ab_variants.py enter 19 50 '' 'ECFFT_B200_TMA=1' 'ECFFT_B200_TW_PREFETCH=1' 'ECFFT_B200_TMA=1,ECFFT_B200_TW_PREFETCH=1' 2>&1 | tee gpurun_out/r02_j_ab_enter19.txt
python tools/ab_variants.py exit 20 20 '' 'ECFFT_B200_TMA=1' 'ECFFT_B200_TW_PREFETCH=1' 2>&1 | tee gpurun_out/r02_j_ab_exit20.txt
ECFFT_B200_TMA=1 ECFFT_B200_TW_PREFETCH=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_j_pytest_tma_pf.log
