#!/bin/bash
# bench + ncu launch list + dram traffic of one step + full capture of the top-level launches
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_final.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_bench.csv &
SMI=$!
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
kill $SMI
cat gpurun_out/bench_final.json | cut -c1-1500
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 45 -c 6 -o gpurun_out/prof_extend_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
python tools/bench_configs.py 22 2>&1 | tee gpurun_out/bench_configs_final.jsonl | cut -c1-150
ls -la gpurun_out | tail -8
