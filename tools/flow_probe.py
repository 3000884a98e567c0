#!/usr/bin/env python
"""Times a device-resident ENTER and prints the flow kernel's diagnostics (ecfft_flow_stats): share of the CTAs'
time spent waiting for input blocks, in tile bodies and publishing results.  Variants come from the environment
(ECFFT_B200_FLOW, ECFFT_B200_FLOW_ORDER, ECFFT_B200_SYM_VARIANT); one JSON line per run."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ecfft_b200
from ecfft_b200 import _lib
from oracle import oracle as O


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    n = 1 << log_n
    L = _lib.load()
    tree = ecfft_b200.build_fftree(n, parts=ecfft_b200.PARTS_ENTER_ONLY)
    xs = [torch.from_numpy(O.random_elements(n, seed=s).view(np.int64)).cuda() for s in (1, 2)]
    for i in range(3):
        tree.enter(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.ecfft_launch_count()
    e0.record()
    for i in range(reps):
        tree.enter(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    launches = (L.ecfft_launch_count() - l0) / reps
    out = {"log_n": log_n, "ms": ms, "launches": launches,
           "env": {k: v for k, v in os.environ.items() if k.startswith("ECFFT_B200")}}
    if os.environ.get("ECFFT_B200_FLOW", "0") != "0":
        _lib.check(L.ecfft_flow_stats(1, None))
        for i in range(4):
            tree.enter(xs[i % 2])
        st = (ctypes.c_ulonglong * 4)()
        _lib.check(L.ecfft_flow_stats(0, st))
        wait, body, sig, tiles = [int(v) for v in st]
        tot = max(wait + body + sig, 1)
        tiles = max(tiles, 1)
        out.update({"tiles_per_enter": tiles / 4, "wait_frac": wait / tot, "body_frac": body / tot, "sig_frac": sig / tot,
                    "wait_cycles_per_tile": wait / tiles, "body_cycles_per_tile": body / tiles, "sig_cycles_per_tile": sig / tiles})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
