"""2+ GPU probe: what does an all-gather / pairwise exchange of ENTER-sized chunks cost here?"""
import os, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for mib in (16, 64):
    x = torch.empty((mib << 20) // 8, dtype=torch.int64, device="cuda").random_()
    g = torch.empty(x.numel() * world, dtype=torch.int64, device="cuda")
    t = timed(lambda: dist.all_gather_into_tensor(g, x))
    y = torch.empty_like(x)
    peer = rank ^ 1
    def xchg():
        ops = [dist.P2POp(dist.isend, x, peer), dist.P2POp(dist.irecv, y, peer)]
        for w in dist.batch_isend_irecv(ops): w.wait()
    t2 = timed(xchg)
    t3 = timed(lambda: torch.empty(x.numel() * world, dtype=torch.int64, device="cuda"))
    if rank == 0:
        print(f"{mib} MiB/rank: all_gather {t:.3f} ms ({mib*(world-1)/t/1.024:.0f} GB/s in), sendrecv {t2:.3f} ms ({mib/t2/1.024:.0f} GB/s), alloc {t3:.4f} ms", flush=True)
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
