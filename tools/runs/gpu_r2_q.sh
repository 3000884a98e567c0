#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python tools/e2e_probe.py 22 10 '' 'ECFFT_B200_HOST_TRACE=1' 'ECFFT_B200_HOST_PIPE=0' 'ECFFT_B200_ENTER_STREAMS=2' 'ECFFT_B200_ENTER_STREAMS=1' 'ECFFT_B200_ENTER_STREAMS=2,ECFFT_B200_HOST_TRACE=1' 2>&1 | tee gpurun_out/r02_q_e2e_probe.txt
python tools/ab_variants.py enter 22 20 '' 'ECFFT_B200_SYM_VARIANT=0' '' 'ECFFT_B200_SYM_VARIANT=0' 2>&1 | tee gpurun_out/r02_q_ab_shape.txt
