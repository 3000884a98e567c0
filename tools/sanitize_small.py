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
# round 2: the batched host call, a cut of ENTER's depth range (folded combines), the second field
xs = np.stack([O.random_elements(1 << 11, seed=20 + i) for i in range(3)])
ys = g.enter_many(xs)
ok = ok and all((ys[i] == c.enter(xs[i])).all() for i in range(3))
from oracle import m31_ref
g31, c31 = ecfft_b200.m31.build_fftree(1 << 9), m31_ref.FFTree.build(1 << 9)
v = [(i * 2654435761) % m31_ref.P for i in range(1 << 9)]
e31 = g31.enter(np.asarray(v, dtype=np.uint32))
ok = ok and e31.tolist() == c31.enter(v) and g31.exit(e31).tolist() == v
ok = ok and g31.degree(e31) == c31.degree(c31.enter(v)) and g31.vanish(np.asarray(v[:128], dtype=np.uint32)).tolist() == c31.vanish(v[:128])
print("sanitize run parity", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
