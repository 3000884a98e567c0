#!/usr/bin/env python
"""Times the host-buffer ecfft_enter (pinned memory in and out) under environment variants, one process each:
  tools/e2e_probe.py <log_n> <reps> 'K=V,...' ...   (ECFFT_B200_HOST_TRACE=1 prints the device timeline of every call)"""
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(log_n, reps):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import ecfft_b200
    from ecfft_b200 import _lib
    from oracle import oracle as O
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(0))
    except Exception:
        pass
    n = 1 << log_n
    L = _lib.load()
    tree = ecfft_b200.build_fftree(n, parts=ecfft_b200.PARTS_ENTER_ONLY)
    hin = [torch.from_numpy(O.random_elements(n, seed=1 + i).view(np.int64)).pin_memory() for i in range(2)]
    hout = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    ip = [ctypes.c_void_p(h.data_ptr()) for h in hin]
    op = ctypes.c_void_p(hout.data_ptr())
    for i in range(3):
        _lib.check(L.ecfft_enter(tree._h, ip[i % 2], n, op))
    t0 = time.perf_counter()
    for i in range(reps):
        _lib.check(L.ecfft_enter(tree._h, ip[i % 2], n, op))
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"log_n": log_n, "e2e_ms": round(ms, 3), "env": {k: v for k, v in os.environ.items() if k.startswith("ECFFT_B200")}}))


def main():
    if sys.argv[1] == "--child":
        return child(int(sys.argv[2]), int(sys.argv[3]))
    for spec in sys.argv[3:] or [""]:
        env = dict(os.environ)
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            env[k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", sys.argv[1], sys.argv[2]], env=env, capture_output=True, text=True)
        tr = [l for l in r.stderr.splitlines() if "trace" in l]
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else f"FAILED {r.stderr[-300:]}", flush=True)
        for l in tr[-3:]:
            print("   ", l, flush=True)


if __name__ == "__main__":
    main()
