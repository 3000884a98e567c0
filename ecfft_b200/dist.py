"""Multi-GPU ENTER: one process per GPU, torch.distributed for the plumbing.

ENTER's recursion (reference src/fftree.rs:143-161) splits the coefficient vector into
contiguous halves that are entered independently on the half-size subtree, so with G ranks the
bottom log2(n/G) recursion depths of chunk g are an independent ENTER(n/G) on rank g.  The only
exchange step of the path is the recombine: one all-gather of the G evaluation chunks
(n/G x 32 B per rank) over NVLink, after which the top log2(G) depths run on the gathered
vector.  (DESIGN.md "multi-GPU" discusses the Amdahl share of those top depths.)
"""
import torch
import torch.distributed as dist


def enter_sharded(tree, chunk, n, group=None):
    """chunk: this rank's n/G coefficients ((n/G, 4) limb tensor on the tree's device, rank order =
    coefficient order).  Returns the full evaluation vector (n, 4) on every rank."""
    world = dist.get_world_size(group)
    if n % world or (n // world) & (n // world - 1):
        raise ValueError("n / world_size must be a power of two")
    if chunk.shape[0] != n // world:
        raise ValueError("chunk must hold n / world_size coefficients")
    local = tree.enter_range(chunk, 1, n // world)
    if world == 1:
        return local
    gathered = torch.empty((n, 4), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    return tree.enter_range(gathered, n // world, n)
