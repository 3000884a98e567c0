#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python tools/e2e_probe.py 22 10 'ECFFT_B200_HOST_TRACE=1' 'ECFFT_B200_PDL=2,ECFFT_B200_HOST_TRACE=1' '' 'ECFFT_B200_PDL=2' 2>&1 | tee gpurun_out/r02_ah_e2e_probe.txt
