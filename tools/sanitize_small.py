"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family at sizes that
cross the tile (packed, strided, pair tiles, fused combine, strided REDC views, radix-2 fallbacks)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ecfft_b200
from oracle import oracle as O

n = 1 << 13
g = ecfft_b200.build_fftree(n)
c = O.OracleTree.build(n)
x = O.random_elements(n, seed=3)
ok = (g.enter(x) == c.enter(x)).all()
ev = c.enter(x)
ok = ok and (g.exit(ev) == x).all()
for h in (n // 2, 1024, 64, 2):
    for m in (0, 1):
        ok = ok and (g.extend(x[:h], m) == c.extend(x[:h], m)).all()
xnn = c.table("xnn_s")
ok = ok and (g.redc_z0(ev, xnn) == c.redc_z0(ev, xnn)).all()
ok = ok and (g.modular_reduce(ev, xnn, c.table("z0z0_rem_xnn_s")) == c.modular_reduce(ev, xnn, c.table("z0z0_rem_xnn_s"))).all()
ok = ok and (g.vanish(x[: n // 2]) == c.vanish(x[: n // 2])).all()
ok = ok and g.degree(ev) == c.degree(ev)
print("sanitize run parity", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
