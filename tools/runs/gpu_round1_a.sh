#!/bin/bash
# smoke + bench + ncu launch list + ncu full capture of the top-level k_extend_tile launches
set -x
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_bench.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
kill $SMI
cat gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_tile -s 45 -c 5 -o gpurun_out/prof_extend_a -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
