"""Multi-process parity check of the sharded ENTER schedules (run under torch.distributed.run on N GPUs):
every rank compares enter_sharded_peer / enter_sharded / enter_sharded_allgather against a single-GPU
ENTER of the whole vector on its own device.  Prints one line per rank; exit code 1 on any mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import ecfft_b200
from ecfft_b200.dist import PeerArena, enter_sharded, enter_sharded_allgather, enter_sharded_peer, exit_sharded_peer
from oracle import oracle as O


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    n = 1 << log_n
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tree = ecfft_b200.build_fftree(n, device=local)
    x = torch.from_numpy(O.random_elements(n, seed=9).view(np.int64)).to(dev)
    want = tree.enter(x)
    c = n // world
    chunk = x[rank * c:(rank + 1) * c]
    arena = PeerArena.create(n, local)
    ok = True
    for rep in range(3):
        got = enter_sharded_peer(tree, chunk, n, arena)
        ok = ok and bool((got == want).all())
    ok = ok and bool((enter_sharded_peer(tree, chunk, n, arena, native=False) == want).all())   # step-by-step driver
    part = enter_sharded_peer(tree, chunk, n, arena, gather=False)
    ok = ok and bool((part == want[rank * c:(rank + 1) * c]).all())
    # sharded EXIT on the same arenas: inverse of the sharded ENTER, and arbitrary values against single-GPU EXIT
    back = exit_sharded_peer(tree, want[rank * c:(rank + 1) * c].contiguous(), n, arena, gather=False)
    ok = ok and bool((back == chunk).all())
    ok = ok and bool((exit_sharded_peer(tree, chunk, n, arena) == tree.exit(x)).all())
    ok = ok and arena.status() == 0
    ok_nccl = bool((enter_sharded(tree, chunk, n) == want).all()) and bool((enter_sharded_allgather(tree, chunk, n) == want).all())
    torch.cuda.synchronize()
    print(f"rank {rank}/{world} n=2^{log_n}: peer {'OK' if ok else 'MISMATCH'}, nccl schedules {'OK' if ok_nccl else 'MISMATCH'}", flush=True)
    arena.close()
    dist.destroy_process_group()
    sys.exit(0 if ok and ok_nccl else 1)


if __name__ == "__main__":
    main()
