#!/bin/bash
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python tools/sanitize_small.py 2>&1 | tail -1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitize_memcheck.log python tools/sanitize_small.py 2>&1 | tail -1
tail -3 gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitize_racecheck.log python tools/sanitize_small.py 2>&1 | tail -1
tail -3 gpurun_out/sanitize_racecheck.log
python bench.py --log-n 24 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2p24_1gpu.json 2>gpurun_out/bench_2p24.err; python -c "
import json; d=json.load(open('gpurun_out/bench_2p24_1gpu.json')); print('n=2^24 1 GPU', round(d['ms_per_step'],3), 'ms', round(d['value']/1e6,1), 'M evals/s; e2e ms', round(d['e2e']['ms_per_step'],2), d['e2e']['matches_device_path'], 'build s', d['config']['tree_build_s'])" || tail -5 gpurun_out/bench_2p24.err
