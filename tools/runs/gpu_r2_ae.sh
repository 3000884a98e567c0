#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for lg in 16 19; do
python tools/ab_variants.py enter $lg 30 '' 'ECFFT_B200_PDL=0' 'ECFFT_B200_TMA=0' 'ECFFT_B200_PDL=0,ECFFT_B200_TMA=0' 'ECFFT_B200_ENTER_STREAMS=1' 'ECFFT_B200_ENTER_STREAMS=1,ECFFT_B200_PDL=0' 'ECFFT_B200_ENTER_STREAMS=2,ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_ae_ab_pdl_tma_streams.txt
python tools/ab_variants.py exit $lg 10 '' 'ECFFT_B200_PDL=0' 'ECFFT_B200_TMA=0' 'ECFFT_B200_PDL=0,ECFFT_B200_TMA=0' 2>&1 | tee -a gpurun_out/r02_ae_ab_pdl_tma_streams.txt
done
python tools/ab_variants.py enter 22 20 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_ae_ab_pdl_tma_streams.txt
python tools/ab_variants.py exit 22 5 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_ae_ab_pdl_tma_streams.txt
python tools/ab_variants.py enter 12 100 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_ae_ab_pdl_tma_streams.txt
python tools/ab_variants.py exit 12 50 '' 'ECFFT_B200_PDL=0' 2>&1 | tee -a gpurun_out/r02_ae_ab_pdl_tma_streams.txt
