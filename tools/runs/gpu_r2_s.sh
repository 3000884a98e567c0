#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for round in 1 2; do
python tools/ab_variants.py enter 22 20 '' 'ECFFT_B200_L2PF=1' 'ECFFT_B200_L2PF=2' 2>&1 | tee -a gpurun_out/r02_s_ab_l2pf_v2.txt
done
python tools/ab_variants.py enter 20 50 '' 'ECFFT_B200_L2PF=1' 'ECFFT_B200_L2PF=2' 2>&1 | tee -a gpurun_out/r02_s_ab_l2pf_v2.txt
