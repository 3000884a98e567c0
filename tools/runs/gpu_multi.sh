#!/bin/bash
# usage: gpu_multi.sh N  — parity + bench at N GPUs under torchrun, all three top-depth schedules
N=$1
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29513 tools/mg_check.py 18 2>gpurun_out/mg_check_$N.err | grep "rank" || tail -20 gpurun_out/mg_check_$N.err
timeout 300 $RUN --master-port 29514 tools/mg_check.py 22 2>>gpurun_out/mg_check_$N.err | grep "rank" || tail -20 gpurun_out/mg_check_$N.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_multi_1.err > gpurun_out/bench_multi_1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_multi_1.json')); print('N=1', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,1),'M evals/s; e2e', round(d['e2e']['value']/1e6,1))"
for mode in peer sharded allgather; do
timeout 300 $RUN --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --multi-gpu $mode 2>gpurun_out/bench_multi_${N}_$mode.err > gpurun_out/bench_multi_${N}_$mode.json
python -c "
import json; d=json.load(open('gpurun_out/bench_multi_${N}_$mode.json')); print('N=$N $mode', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,1),'M evals/s; e2e', round(d['e2e']['value']/1e6,1), 'kernel ms', round(d['roofline']['kernel_ms_per_step'],2))" || tail -5 gpurun_out/bench_multi_${N}_$mode.err
done
