#!/usr/bin/env python
"""Debugging aid: runs one operation under several environment variants (one process each), keeps the outputs
and reports where each variant differs from the first (element indices, as ranges).

  tools/variant_diff.py <enter|extend> <log_n> 'K=V,...' 'K=V,...' ...
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(op, log_n, path):
    sys.path.insert(0, ROOT)
    import torch
    import ecfft_b200
    from oracle import oracle as O
    n = 1 << log_n
    log_tree = log_n + 1 if op == "extend" else log_n
    tree = ecfft_b200.build_fftree(1 << log_tree, parts=ecfft_b200.PARTS_ENTER_ONLY)
    x = torch.from_numpy(O.random_elements(n, seed=1).view(np.int64)).cuda()
    y = tree.enter(x) if op == "enter" else tree.extend(x, 1)
    torch.cuda.synchronize()
    np.save(path, y.cpu().numpy())


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2], int(sys.argv[3]), sys.argv[4])
    op, log_n = sys.argv[1], sys.argv[2]
    base = None
    for k, spec in enumerate(sys.argv[3:]):
        env = dict(os.environ)
        for kv in filter(None, spec.split(",")):
            a, b = kv.split("=")
            env[a] = b
        path = f"/tmp/variant_{k}.npy"
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", op, log_n, path], env=env, capture_output=True, text=True)
        if r.returncode:
            print(f"[{spec}] FAILED: {r.stderr[-300:]}")
            continue
        y = np.load(path).reshape(-1, 4)
        if base is None:
            base = y
            print(f"[{spec}] baseline, {len(y)} elements")
            continue
        bad = np.nonzero((y != base).any(axis=1))[0]
        if len(bad) == 0:
            print(f"[{spec}] identical")
            continue
        runs, s0, prev = [], bad[0], bad[0]
        for b in bad[1:]:
            if b != prev + 1:
                runs.append((s0, prev))
                s0 = b
            prev = b
        runs.append((s0, prev))
        print(f"[{spec}] {len(bad)} elements differ in {len(runs)} runs; first runs: {[(int(a), int(b - a + 1)) for a, b in runs[:12]]}")


if __name__ == "__main__":
    main()
