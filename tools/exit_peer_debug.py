"""debug: sharded EXIT with virtual ranks on one GPU, prints arena status words"""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ecfft_b200
from ecfft_b200.dist import PeerArena, exit_sharded_peer
from oracle import oracle as O
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = 1 << 14
gpu = ecfft_b200.build_fftree(n)
cpu = O.OracleTree.build(n)
x = O.random_elements(n, seed=5)
want = cpu.exit(x)
xd = torch.from_numpy(x.view(np.int64)).cuda()
c = n // world
arenas = PeerArena.local_group(n, world, device=0)
res, errs = [None] * world, []
torch.cuda.synchronize()
def run(rank):
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            from ecfft_b200.dist import enter_sharded_peer
            seq = os.environ.get("SEQ", "xex")
            for ch in seq:
                if ch == "x":
                    out = exit_sharded_peer(gpu, xd[rank*c:(rank+1)*c], n, arenas[rank], gather=False)
                else:
                    enter_sharded_peer(gpu, xd[rank*c:(rank+1)*c], n, arenas[rank], gather=False)
            s.synchronize()
            res[rank] = out.cpu().numpy().view(np.uint64)
    except Exception as e:
        errs.append((rank, repr(e)[:200]))
ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
[t.start() for t in ts]; [t.join(timeout=120) for t in ts]
print("errors", errs)
try:
    for r in range(world):
        st = arenas[r].status()
        print("rank", r, "status", hex(st), "epoch", (st >> 24) & 0xffffffff, "sid", (st >> 8) & 0xffff, "which", st & 0xff)
except Exception as e:
    print("status read failed", e)
for r in range(world):
    if res[r] is not None:
        print("rank", r, "match", bool((res[r] == want[r*c:(r+1)*c]).all()))
