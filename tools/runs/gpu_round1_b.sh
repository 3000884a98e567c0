#!/bin/bash
set -x
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_b.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; cat gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
ECFFT_B200_BUTTERFLY=matrix python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b_matrix.json 2>/dev/null; cat gpurun_out/bench_b_matrix.json
