#!/bin/bash
# usage: gpu_multi_final.sh N [LOGNS...] — parity at 2^22 + peer-schedule bench at N GPUs for each log n
N=$1; shift
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29514 tools/mg_check.py 22 2>gpurun_out/mg_check_$N.err | grep "rank" | sort | head -3 || tail -20 gpurun_out/mg_check_$N.err
for LOGN in "$@"; do
timeout 400 $RUN --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --log-n $LOGN 2>gpurun_out/bench_final_${N}gpu_2p$LOGN.err > gpurun_out/bench_final_${N}gpu_2p$LOGN.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final_${N}gpu_2p$LOGN.json')); print('N=$N n=2^$LOGN', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,1),'M evals/s; e2e', round(d['e2e']['value']/1e6,1), 'M evals/s', round(d['e2e']['ms_per_step'],3), 'ms; kernel ms', round(d['roofline']['kernel_ms_per_step'],2), 'launches', d['gpu_launches'])" || tail -5 gpurun_out/bench_final_${N}gpu_2p$LOGN.err
done
