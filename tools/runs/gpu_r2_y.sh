#!/bin/bash
# N-GPU ENTER 2^22 under stream-count variants (headline step only)
N=$1
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for v in 'ECFFT_B200_ENTER_STREAMS=1' 'ECFFT_B200_ENTER_STREAMS=2' 'ECFFT_B200_ENTER_STREAMS=4'; do
env $(echo $v | tr ',' ' ') timeout 600 $RUN --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-configs 2>gpurun_out/r02_y_${N}gpu.err > gpurun_out/r02_y_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r02_y_${N}gpu.json')); print('N=$N $v:', round(d['ms_per_step'],3),'ms; launches/step', d['gpu_launches']/10, '; without all-gather', round(d['ms_per_step_without_final_allgather'],3), d['multi_gpu_matches_single'])" | tee -a gpurun_out/r02_ag_${N}gpu_streams_after_pdl_rule.txt || tail -5 gpurun_out/r02_y_${N}gpu.err
done
