#!/bin/bash
# compile-time arithmetic variants (tools/_ab/*.so built with -DFP_REDUCE_ALIGNED / -DFP_BRANCHFREE_ADDSUB / both), same box
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
R=$PWD/tools/_ab
for round in 1 2; do
python tools/ab_variants.py enter 22 20 '' "ECFFT_B200_LIB=$R/libecfft_b200_A.so" "ECFFT_B200_LIB=$R/libecfft_b200_B.so" "ECFFT_B200_LIB=$R/libecfft_b200_C.so" 2>&1 | tee -a gpurun_out/r02_aj_ab_arith_variants.txt
done
python tools/ab_variants.py enter 19 50 '' "ECFFT_B200_LIB=$R/libecfft_b200_A.so" "ECFFT_B200_LIB=$R/libecfft_b200_B.so" "ECFFT_B200_LIB=$R/libecfft_b200_C.so" 2>&1 | tee -a gpurun_out/r02_aj_ab_arith_variants.txt
python tools/ab_variants.py exit 22 5 '' "ECFFT_B200_LIB=$R/libecfft_b200_A.so" "ECFFT_B200_LIB=$R/libecfft_b200_B.so" "ECFFT_B200_LIB=$R/libecfft_b200_C.so" 2>&1 | tee -a gpurun_out/r02_aj_ab_arith_variants.txt
