#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for lg in 18 20 21; do
python tools/ab_variants.py enter $lg 30 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_ad_ab_pdl_sizes.txt
python tools/ab_variants.py exit $lg 10 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_ad_ab_pdl_sizes.txt
done
