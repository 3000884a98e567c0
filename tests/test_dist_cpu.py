"""Host-side logic of the multi-GPU ENTER (ecfft_b200/dist.py) on CPU: world_size 2 and 4 over gloo.
Each rank enters its coefficient chunk, one all-gather exchanges the evaluation chunks, the top
recursion depths finish on the gathered vector.  The per-rank compute is played by the CPU oracle's
level-range restatement, so what is tested is the sharding/exchange schedule itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleBackedTree:
    """stands in for ecfft_b200.FFTree.enter_range on CPU tensors"""

    def __init__(self, n):
        from oracle import oracle as O
        self.t = O.OracleTree.build(n, parts=1)

    def enter_range(self, data, m_lo, m_hi):
        arr = data.numpy().view(np.uint64)
        return torch.from_numpy(self.t.enter_range(arr, m_lo, m_hi).view(np.int64))


def _worker(rank, world, port, n, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from ecfft_b200.dist import enter_sharded
        tree = OracleBackedTree(n)
        x = O.random_elements(n, seed=11)
        chunk = torch.from_numpy(x[rank * (n // world):(rank + 1) * (n // world)].view(np.int64).copy())
        got = enter_sharded(tree, chunk, n).numpy().view(np.uint64)
        want = tree.t.enter(x)
        np.save(os.path.join(result_dir, f"ok_{rank}.npy"), np.array([(got == want).all()]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4])
def test_enter_sharded_matches_single_enter(world, tmp_path):
    n = 256
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"ok_{r}.npy")[0]


def test_enter_sharded_rejects_bad_chunking():
    from ecfft_b200.dist import enter_sharded

    class FakeGroup:
        pass

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        tree = OracleBackedTree(16)
        with pytest.raises(ValueError):
            enter_sharded(tree, torch.zeros((3, 4), dtype=torch.int64), 16)
        x = torch.from_numpy(__import__("oracle.oracle", fromlist=["x"]).random_elements(16).view(np.int64))
        out = enter_sharded(tree, x, 16)
        assert out.shape == (16, 4)
    finally:
        dist.destroy_process_group()
