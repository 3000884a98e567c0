#!/bin/bash
# round 2, series t: bench line on N GPUs (peer schedule; checks the sharded ENTER / EXIT / 2^24 results against one GPU)
N=$1
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r02_t_bench_${N}gpu.err > gpurun_out/r02_t_bench_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_t_bench_${N}gpu.json')); print('N=$N', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,1),'M evals/s; e2e', round(d['e2e']['ms_per_step'],3), 'ms;', {k:v for k,v in d.items() if k.startswith('cfg_') or 'matches' in k or 'allgather' in k}, d['self_check_ok'])" || tail -20 gpurun_out/r02_t_bench_${N}gpu.err
