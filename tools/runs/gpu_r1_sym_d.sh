#!/bin/bash
# k_extend_sym (cp.async tile load, fused pre/post scale, fused ENTER combine): parity, A/B
mkdir -p gpurun_out
make -C oracle >/dev/null 2>&1
for v in 0 1 2 3; do echo -n "sym variant $v "; ECFFT_B200_SYM_VARIANT=$v python tools/parity_quick.py 2>&1 | tail -1; done
echo -n "no combine fusion "; ECFFT_B200_NO_COMBINE_FUSION=1 python tools/parity_quick.py 2>&1 | tail -1
b() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab.json; python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('$1', round(d['ms_per_step'],3),'ms; extend', round(d['roofline']['kernel_ms_per_step'],3), '; combine', round(d['roofline']['other_kernels']['k_enter_combine']['ms_per_step'],3), '; e2e ms', round(d['e2e']['ms_per_step'],2), '; launches', d['gpu_launches'])"; }
for round in 1 2; do
for v in 0 1 2 3; do ECFFT_B200_SYM_VARIANT=$v b "round $round sym variant $v"; done
ECFFT_B200_NO_COMBINE_FUSION=1 b "round $round variant 0 without combine fusion"
ECFFT_B200_SYM_RADIX2=1 b "round $round radix-2 kernel v7"
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
