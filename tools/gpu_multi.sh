#!/bin/bash
# usage: gpu_multi.sh N  — bench at N GPUs under torchrun (plus N=1 for the same box)
N=$1
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
python tools/parity_quick.py 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_multi_1.err > gpurun_out/bench_multi_1.json; cat gpurun_out/bench_multi_1.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_multi_$N.err > gpurun_out/bench_multi_$N.json
cat gpurun_out/bench_multi_$N.json; tail -5 gpurun_out/bench_multi_$N.err
