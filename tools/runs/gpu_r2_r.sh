#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for round in 1 2; do
python tools/ab_variants.py enter 22 20 'ECFFT_B200_L2PF=0' '' 'ECFFT_B200_L2PF=0,ECFFT_B200_SYM_VARIANT=0' 'ECFFT_B200_SYM_VARIANT=0' 2>&1 | tee -a gpurun_out/r02_r_ab_l2pf.txt
done
for lg in 16 19 20; do
python tools/ab_variants.py enter $lg 50 '' 'ECFFT_B200_SYM_VARIANT=0' 'ECFFT_B200_L2PF=0' 2>&1 | tee -a gpurun_out/r02_r_ab_shape_small.txt
done
python tools/ab_variants.py exit 22 5 '' 'ECFFT_B200_SYM_VARIANT=0' 'ECFFT_B200_L2PF=0' 2>&1 | tee -a gpurun_out/r02_r_ab_exit.txt
python tools/ab_variants.py extend 20 50 '' 'ECFFT_B200_SYM_VARIANT=0' 'ECFFT_B200_L2PF=0' 2>&1 | tee -a gpurun_out/r02_r_ab_exit.txt
ECFFT_B200_SYM_VARIANT=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
