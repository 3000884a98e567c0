#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_u_pytest.log
for round in 1 2; do
python tools/ab_variants.py exit 22 5 '' 'ECFFT_B200_TMA=0' 2>&1 | tee -a gpurun_out/r02_u_ab_exit_tma_views.txt
done
python tools/ab_variants.py exit 16 30 '' 'ECFFT_B200_TMA=0' 2>&1 | tee -a gpurun_out/r02_u_ab_exit_tma_views.txt
