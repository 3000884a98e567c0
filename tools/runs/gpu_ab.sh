#!/bin/bash
# same-box A/B of library variants: parity (full gpu suite on each) + ENTER 2^22 timing, two rounds
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
LIBS="default tools/_ab/libecfft_b200_chained.so tools/_ab/libecfft_b200_sumform.so tools/_ab/libecfft_b200_sumform_chained.so"
for lib in $LIBS; do
  if [ "$lib" = default ]; then unset ECFFT_B200_LIB; else export ECFFT_B200_LIB=$PWD/$lib; fi
  echo "=== $lib"
  timeout 600 python -m pytest tests -m gpu -x -q -k "not full_size and not matrix_butterfly" 2>&1 | tail -2
done
for round in 1 2; do
for lib in $LIBS; do
  if [ "$lib" = default ]; then unset ECFFT_B200_LIB; else export ECFFT_B200_LIB=$PWD/$lib; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab.json
  python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('round $round $lib', round(d['ms_per_step'],3),'ms; extend', round(d['roofline']['kernel_ms_per_step'],3))"
done; done
