#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for round in 1 2; do
python tools/ab_variants.py enter 22 20 '' 'ECFFT_B200_SYM_AUTO=1' 'ECFFT_B200_SYM_VARIANT=0' 2>&1 | tee -a gpurun_out/r02_ak_ab_auto_shape.txt
done
python tools/ab_variants.py exit 22 5 '' 'ECFFT_B200_SYM_AUTO=1' 2>&1 | tee -a gpurun_out/r02_ak_ab_auto_shape.txt
python tools/ab_variants.py enter 19 50 '' 'ECFFT_B200_SYM_VARIANT=0' 2>&1 | tee -a gpurun_out/r02_ak_ab_auto_shape.txt
