#!/usr/bin/env python
"""Same-box A/B of library variants selected by environment variables (read once per process, so every variant
runs in its own process).  For each variant: device-resident time of one operation and a digest of its output
(the variants must agree bit for bit).

  tools/ab_variants.py <enter|exit|extend> <log_n> <reps> 'K=V,K=V' 'K=V' ...      ('' = defaults)
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(op, log_n, reps):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import ecfft_b200
    from ecfft_b200 import _lib
    from oracle import oracle as O
    n = 1 << log_n
    L = _lib.load()
    log_tree = log_n + 1 if op == "extend" else log_n
    tree = ecfft_b200.build_fftree(1 << log_tree, parts=ecfft_b200.PARTS_ENTER_ONLY if op == "enter" else ecfft_b200.PARTS_FULL)
    xs = [torch.from_numpy(O.random_elements(n, seed=s).view(np.int64)).cuda() for s in (1, 2)]
    fn = {"enter": tree.enter, "exit": tree.exit, "extend": lambda x: tree.extend(x, 1)}[op]
    for i in range(3):
        y = fn(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.ecfft_launch_count()
    e0.record()
    for i in range(reps):
        y = fn(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    y = fn(xs[0])
    torch.cuda.synchronize()
    digest = hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps({"op": op, "log_n": log_n, "ms": round(e0.elapsed_time(e1) / reps, 4),
                      "launches": (L.ecfft_launch_count() - l0) / reps, "sha1": digest,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("ECFFT_B200")}}))


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    op, log_n, reps = sys.argv[1], sys.argv[2], sys.argv[3]
    for spec in sys.argv[4:] or [""]:
        env = dict(os.environ)
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            env[k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", op, log_n, reps], env=env,
                           capture_output=True, text=True, timeout=900)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else f"FAILED rc={r.returncode}: {r.stderr[-400:]}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
