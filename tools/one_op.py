#!/usr/bin/env python
"""Runs ONE device-resident FFTree operation after warm-up (for ncu launch lists):
  tools/one_op.py <enter|exit|extend|redc|mod|roundtrip> <log_n> [log_tree]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ecfft_b200
from oracle import oracle as O


def main():
    op, log_n = sys.argv[1], int(sys.argv[2])
    log_tree = int(sys.argv[3]) if len(sys.argv) > 3 else (log_n + 1 if op == "extend" else log_n)
    n = 1 << log_n
    tree = ecfft_b200.build_fftree(1 << log_tree, parts=ecfft_b200.PARTS_ENTER_ONLY if op == "enter" else ecfft_b200.PARTS_FULL)
    x = torch.from_numpy(O.random_elements(n, seed=1).view(np.int64)).cuda()
    if op in ("redc", "mod"):
        a = torch.from_numpy(tree.table("xnn_s", n).view(np.int64)).cuda()
        c = torch.from_numpy(tree.table("z0z0_rem_xnn_s", n).view(np.int64)).cuda()
    fn = {"enter": lambda: tree.enter(x), "exit": lambda: tree.exit(x), "extend": lambda: tree.extend(x, 1),
          "redc": lambda: tree.redc_z0(x, a), "mod": lambda: tree.modular_reduce(x, a, c),
          "roundtrip": lambda: tree.exit(tree.enter(x))}[op]
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
