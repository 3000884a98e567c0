"""quick bit-exactness check of the current env's kernel variant against the oracle (GPU box)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ecfft_b200
from oracle import oracle as O
n = 1 << 14
g = ecfft_b200.build_fftree(n)
c = O.OracleTree.build(n)
x = O.random_elements(n, seed=3)
ok = (g.enter(x) == c.enter(x)).all() and (g.exit(x) == c.exit(x)).all()
for h in (1 << 13, 1 << 12, 64):
    for m in (0, 1):
        ok = ok and (g.extend(x[:h], m) == c.extend(x[:h], m)).all()
print("parity", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
