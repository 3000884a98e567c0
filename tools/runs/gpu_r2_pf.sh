#!/bin/bash
# round 2, final state (per-launch shape selection on): GPU tests, bench line, launch list of one ENTER with DRAM bytes,
# full capture of six k_extend_sym launches with source
S=p
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_${S}_pytest_gpu.log
python bench.py > gpurun_out/r02_${S}_bench_1gpu.json 2> gpurun_out/r02_${S}_bench_1gpu.err
cut -c1-400 gpurun_out/r02_${S}_bench_1gpu.json; tail -3 gpurun_out/r02_${S}_bench_1gpu.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_${S}_launches_one_enter.csv python tools/one_op.py enter 22 > /dev/null 2>&1
python tools/traffic_from_launches.py gpurun_out/r02_${S}_launches_one_enter.csv gpurun_out/r02_traffic.json "ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on tools/one_op.py enter 22 (one device-resident ENTER n=2^22 after warm-up; launches of the two streams serialised under ncu), tools/runs/gpu_r2_pf.sh"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_extend_sym -s 48 -c 6 -o gpurun_out/r02_${S}_ncu_full_extend_sym -f python tools/one_op.py enter 22 > gpurun_out/r02_${S}_ncu_full.log 2>&1
python tools/ncu_summary.py full gpurun_out/r02_${S}_ncu_full_extend_sym.ncu-rep > gpurun_out/r02_${S}_ncu_full_extend_sym.txt 2>&1; head -12 gpurun_out/r02_${S}_ncu_full_extend_sym.txt | cut -c1-160
