#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -k "full_size" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_e.log
python tools/bench_configs.py 22 2>&1 | tee gpurun_out/bench_configs_e.jsonl | cut -c1-220
