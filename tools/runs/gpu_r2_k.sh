#!/bin/bash
# round 2, series k: where the TMA tile load differs; folded combines: parity tests and A/B
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for lg in 12 16 21; do
python tools/variant_diff.py extend $lg '' 'ECFFT_B200_TMA=1' 'ECFFT_B200_TMA=1' 'ECFFT_B200_TMA=1,ECFFT_B200_PDL=0' 'ECFFT_B200_TMA=2' 2>&1 | tee -a gpurun_out/r02_k_tma_diff.txt
done
python tools/variant_diff.py enter 16 '' 'ECFFT_B200_TMA=1' 'ECFFT_B200_TMA=1,ECFFT_B200_PDL=0' 'ECFFT_B200_TMA=2' 'ECFFT_B200_FOLD=0' 2>&1 | tee -a gpurun_out/r02_k_tma_diff.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_k_pytest_fold.log
for round in 1 2; do
python tools/ab_variants.py enter 22 20 '' 'ECFFT_B200_FOLD=0' 2>&1 | tee -a gpurun_out/r02_k_ab_fold.txt
done
python tools/ab_variants.py enter 19 50 '' 'ECFFT_B200_FOLD=0' 2>&1 | tee -a gpurun_out/r02_k_ab_fold.txt
python tools/ab_variants.py enter 16 100 '' 'ECFFT_B200_FOLD=0' 2>&1 | tee -a gpurun_out/r02_k_ab_fold.txt
