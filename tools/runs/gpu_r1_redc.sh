#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/bench_configs.py 22 2>&1 | tee gpurun_out/bench_configs_redc.jsonl | cut -c1-120
echo "--- without REDC fusion"
ECFFT_B200_NO_REDC_FUSION=1 python tools/bench_configs.py 22 2>&1 | grep -E "REDC|MOD|EXIT|ENTER n" | cut -c1-120
