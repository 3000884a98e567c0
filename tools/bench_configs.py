#!/usr/bin/env python
"""Times every single-GPU configuration of BASELINE.json with device-resident buffers (CUDA events,
3 warm-ups + 10 samples, median/best) and prints one JSON object per config.  Not the headline
bench (bench.py); evidence for the other rows of SURVEY.md section 8."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ecfft_b200
from oracle import oracle as O


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def dev(a):
    return torch.from_numpy(a.view(np.int64)).cuda()


def main():
    log_top = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    tree = ecfft_b200.build_fftree(1 << log_top)
    res = []

    def rec(name, n, fn, modmul_per_elem):
        med, best = timeit(fn)
        res.append({"config": name, "n": n, "ms_median": med, "ms_best": best, "elems_per_s": n / (med * 1e-3),
                    "modmul_per_s_reference_count": modmul_per_elem * n / (med * 1e-3)})
        print(json.dumps(res[-1]), flush=True)

    n = 1 << 12
    x = dev(O.random_elements(n, seed=1))
    rec("ENTER->EXIT roundtrip n=2^12", n, lambda: tree.exit(tree.enter(x)), 276 + 618)
    n = 1 << 20
    x = dev(O.random_elements(n, seed=2))
    rec("EXTEND n=2^20 (tree 2^21)", n, lambda: tree.extend(x, 1), 80)
    xnn = dev(tree.table("xnn_s", n))
    zz = dev(tree.table("z0z0_rem_xnn_s", n))
    rec("REDC n=2^20", n, lambda: tree.redc_z0(x, xnn), 79)
    rec("MOD n=2^20", n, lambda: tree.modular_reduce(x, xnn, zz), 159)
    rec("MEXTEND n=2^20", n, lambda: tree.mextend(x, 1), 80)
    rec("VANISH n=2^20 (out 2^21)", n, lambda: tree.vanish(x), 0)
    rec("DEGREE n=2^20", n, lambda: tree.degree(x), 0)
    n = 1 << log_top
    x = dev(O.random_elements(n, seed=3))
    L = log_top
    rec(f"ENTER n=2^{log_top}", n, lambda: tree.enter(x), 2 * L * (L - 1) + L)
    rec(f"EXIT n=2^{log_top}", n, lambda: tree.exit(x), 2013)


if __name__ == "__main__":
    main()
