#!/bin/bash
# usage: gpu_multi_quick.sh N — parity at 2^22 + bench of the peer and all-gather schedules at N GPUs
N=$1
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29514 tools/mg_check.py 22 2>gpurun_out/mg_check_$N.err | grep "rank" || tail -20 gpurun_out/mg_check_$N.err
for mode in peer allgather; do
timeout 300 $RUN --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --multi-gpu $mode 2>gpurun_out/bench_multi_${N}_$mode.err > gpurun_out/bench_multi_${N}_$mode.json
python -c "
import json; d=json.load(open('gpurun_out/bench_multi_${N}_$mode.json')); print('N=$N $mode', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,1),'M evals/s; e2e', round(d['e2e']['value']/1e6,1), 'M evals/s', round(d['e2e']['ms_per_step'],3), 'ms; kernel ms', round(d['roofline']['kernel_ms_per_step'],2), 'launches', d['gpu_launches'])" || tail -5 gpurun_out/bench_multi_${N}_$mode.err
done
