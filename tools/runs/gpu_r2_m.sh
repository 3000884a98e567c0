#!/bin/bash
# round 2, series m: TMA tile load after the box-layout fix (diff + tests + A/B), enter_many, fold cuts
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for lg in 11 12 16 21; do python tools/variant_diff.py extend $lg '' 'ECFFT_B200_TMA=1' 2>&1 | tee -a gpurun_out/r02_m_tma_diff.txt; done
for lg in 12 13 16 20; do python tools/variant_diff.py enter $lg '' 'ECFFT_B200_TMA=1' 2>&1 | tee -a gpurun_out/r02_m_tma_diff.txt; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_m_pytest.log
ECFFT_B200_TMA=1 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_m_pytest_tma.log
for round in 1 2; do
python tools/ab_variants.py enter 22 20 '' 'ECFFT_B200_TMA=1' 2>&1 | tee -a gpurun_out/r02_m_ab_tma.txt
done
python tools/ab_variants.py enter 19 50 '' 'ECFFT_B200_TMA=1' 2>&1 | tee -a gpurun_out/r02_m_ab_tma.txt
python tools/ab_variants.py exit 22 5 '' 'ECFFT_B200_TMA=1' 2>&1 | tee -a gpurun_out/r02_m_ab_tma.txt
python tools/ab_variants.py extend 20 50 '' 'ECFFT_B200_TMA=1' 2>&1 | tee -a gpurun_out/r02_m_ab_tma.txt
