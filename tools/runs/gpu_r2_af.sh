#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for lg in 18 19 20; do
python tools/ab_variants.py enter $lg 30 '' 'ECFFT_B200_ENTER_STREAMS=1' 'ECFFT_B200_ENTER_STREAMS=2' 'ECFFT_B200_ENTER_STREAMS=4' 'ECFFT_B200_PDL=0' 'ECFFT_B200_PDL=2' 2>&1 | tee -a gpurun_out/r02_af_ab_streams_after_pdl_rule.txt
done
python tools/ab_variants.py enter 21 20 '' 'ECFFT_B200_ENTER_STREAMS=1' 'ECFFT_B200_ENTER_STREAMS=4' 2>&1 | tee -a gpurun_out/r02_af_ab_streams_after_pdl_rule.txt
for lg in 18 19 20 21; do
python tools/ab_variants.py exit $lg 10 '' 'ECFFT_B200_PDL=0' 'ECFFT_B200_PDL=2' 2>&1 | tee -a gpurun_out/r02_af_ab_streams_after_pdl_rule.txt
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_af_pytest.log
