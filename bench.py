#!/usr/bin/env python
"""Headline benchmark: FFTree<secp256k1::Fp>::enter throughput (evals/s) at n = 2^22.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n 22]

One JSON line on stdout (rank 0).  A step = one ENTER of n synthetic random coefficients.
  value : evals/s with input and output resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the host-buffer C ABI call (pinned host in -> H2D -> ENTER -> D2H)
  roofline : the dominant kernel against the measured HBM peak, using the ALGORITHMIC bytes of the
             level-streaming model (DESIGN.md / SURVEY.md 8d); the kernel fuses ~10 levels per HBM round trip,
             so `achieved` exceeds the physical peak and `traffic` (ncu) is ~10x smaller; the flat
             `integer_pipe_frac` reports the pipe that actually binds it
  cpu_baseline : the CPU oracle (single thread, like the reference library) on a bounded sample
  cfg_* : the other BASELINE.json configurations, timed in the same run (flat keys):
          N = 1: ENTER->EXIT round trip 2^12, EXTEND 2^20, REDC / MOD 2^20, EXIT 2^22
          N > 1: ENTER 2^24 sharded over the N GPUs; every rank checks the sharded results (2^22 and 2^24)
                 bit for bit against a single-GPU ENTER of the whole vector on its own device
--impl reference times the CPU restatement of the reference (oracle/, all host threads) at the same n.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "secp256k1 Fp ENTER evals/sec at n=2^22"
print_json = None
UNIT = "evals/s"


def modmuls_per_elem(log_n):
    # ENTER(n) = 2 n L (L-1) + n L field multiplications
    return 2 * log_n * (log_n - 1) + log_n


def products_per_elem(log_n):
    """256-bit field products THIS engine executes per coefficient of ENTER(n) (symmetric butterflies):
    depth with vectors of length 2^L: 2 L levels of 1/2 product per element, minus the merged centre level
    (L >= 2), and a combine of 2 (both outputs two-product dots: the next depth's pre-scale rides in the tables).
    With ECFFT_B200_FOLD=0: pre-scale 1 (L >= 1) and a combine of 3/2."""
    fold = os.environ.get("ECFFT_B200_FOLD", "1") != "0"
    tot = 0.0
    for L in range(log_n):
        if fold:
            tot += L - (0.5 if L >= 2 else 0) + 2.0
        else:
            tot += L + (1 if L >= 1 else 0) - (0.5 if L >= 2 else 0) + 1.5
    return tot


# measured on this pool's B200 with tools/microbench.cu (profiles/r01_microbench_pipes.txt)
IMAD_WIDE_PEAK = 9.1e12      # IMAD.WIDE.U32.X chains /s, the binding pipe of a 256-bit product
PRODUCT_PEAK = 108e9         # register-resident fp_mul_lazy /s (64 + 13 IMAD.WIDE each)


def alg_bytes_per_elem(log_n):
    # level passes 64 B/elem each, matrices ~256 B/elem in total, combines 128 B/elem each
    return 64 * log_n * (log_n - 1) + 256 + 128 * log_n


def ncu_traffic(log_n):
    """DRAM bytes the dominant kernel moved per step in the committed ncu capture (profiles/), n = 2^22 only"""
    if log_n != 22:
        return None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            k = t.get("k_enter_flow") or t.get("k_extend_sym") or t["k_extend_tile"]
            return k["dram_total_gb_per_step"], f"GB per step over {k['launches_per_step']} launches (profiles/{name}; algorithmic bytes are per step too)"
        except Exception:
            continue
    return None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s >= 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(log_n_sample, log_n_target, threads):
    """time the oracle's ENTER on a bounded sample and scale by modmul count to the target size"""
    from oracle import oracle as O
    ns = 1 << log_n_sample
    tree = O.OracleTree.build(ns, parts=1, threads=os.cpu_count() or 1)
    x = O.random_elements(ns, seed=1)
    t0 = time.perf_counter()
    out = tree.enter(x, threads=threads)
    dt = time.perf_counter() - t0
    scale = (modmuls_per_elem(log_n_target) * (1 << log_n_target)) / (modmuls_per_elem(log_n_sample) * ns)
    est_full = dt * scale
    return {"seconds": dt, "n_sample": ns, "est_seconds_full": est_full, "evals_per_s": (1 << log_n_target) / est_full,
            "out": out, "x": x}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation.  The Rust crate cannot be compiled in this
    image (no cargo/rustc, arkworks not vendored), so this is the C restatement under oracle/ ("port"),
    run at the SAME n as the GPU arm: every step is one whole ENTER(n) on all host threads."""
    if rank != 0:
        return
    log_n = args.log_n
    threads = os.cpu_count() or 1
    log_s = log_n if args.ref_sample_log_n is None else min(log_n, args.ref_sample_log_n)
    from oracle import oracle as O
    ns = 1 << log_s
    t_build = time.perf_counter()
    tree = O.OracleTree.build(ns, parts=1, threads=threads)
    t_build = time.perf_counter() - t_build
    x = O.random_elements(ns, seed=1)
    for _ in range(args.warmup):
        tree.enter(x, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tree.enter(x, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    scale = (modmuls_per_elem(log_n) * (1 << log_n)) / (modmuls_per_elem(log_s) * ns)
    value = (1 << log_n) / (dt * scale)
    if log_s == log_n:
        sample = f"each step = one whole ENTER n=2^{log_n} on {threads} threads ({dt:.3f} s per step, measured, not extrapolated)"
    else:
        sample = (f"each step = ENTER n=2^{log_s} on {threads} threads ({dt:.3f} s), scaled by field-multiplication "
                  f"count x{scale:.1f} to n=2^{log_n}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64x4 Montgomery (ark-ff layout)", "data": "synthetic",
        "config": {"workload": f"secp256k1::Fp ENTER n=2^{log_n} (full log^2 recursion) on a 2^{log_n}-leaf FFTree",
                   "implementation": "CPU restatement of the reference (oracle/, C, all host threads)",
                   "tree_build_s": round(t_build, 1)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(line)


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so that the pinned host buffers of
    the e2e path are allocated (first touch) on the GPU's NUMA node and PCIe copies do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return f"cpu affinity set to GPU {index}'s NUMA node ({len(os.sched_getaffinity(0))} cpus)"
    except Exception as e:  # not fatal: the copies are just slower
        return f"cpu affinity not set ({type(e).__name__})"


def _protect_stdout():
    """Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the whole run and return a writer on the original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    real_stdout = _protect_stdout()
    global print_json
    print_json = lambda obj: (real_stdout.write(json.dumps(obj) + "\n"), real_stdout.flush())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--cpu-sample-log-n", type=int, default=18)
    ap.add_argument("--ref-sample-log-n", type=int, default=None,
                    help="reference arm: time ENTER at this smaller size and scale (default: the real n, measured)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg_* timings of the other BASELINE configurations")
    ap.add_argument("--multi-gpu", default="peer", choices=["peer", "sharded", "allgather"],
                    help="top-depth schedule for N>1: sharded with the kernels reading the partner's buffers over NVLink "
                         "(peer, default), sharded with NCCL send/recv per straddling level, or one all-gather + replicated top depths")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import ecfft_b200
    from ecfft_b200 import _lib
    from ecfft_b200.dist import PeerArena, enter_sharded, enter_sharded_allgather, enter_sharded_peer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ecfft_b200 has no CPU fallback")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()

    log_n = args.log_n
    n = 1 << log_n
    t_build = time.perf_counter()
    tree = ecfft_b200.build_fftree(n, parts=ecfft_b200.PARTS_ENTER_ONLY, device=local_rank)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    # synthetic inputs: uniform field elements as raw Montgomery limbs (splitmix64, seed 1).  Four
    # different vectors are rotated; with the 1.3 GB of tables every step streams far more than L2 holds.
    from oracle import oracle as O  # only the seeded generator + (rank 0) the CPU baseline / check
    NBUF = 4
    host_in = [torch.from_numpy(O.random_elements(n, seed=1 + i).view(np.int64)).pin_memory() for i in range(NBUF)]
    chunk = n // world
    dev_in = [h[rank * chunk:(rank + 1) * chunk].to(dev) for h in host_in]

    arena = None
    if world > 1 and args.multi_gpu == "peer":
        arena = PeerArena.create(n, local_rank)
        shard_fn = lambda tree_, x_, n_, **kw: enter_sharded_peer(tree_, x_, n_, arena, **kw)
    elif args.multi_gpu == "allgather":
        shard_fn = lambda tree_, x_, n_, **kw: enter_sharded_allgather(tree_, x_, n_)
    else:
        shard_fn = enter_sharded

    def step(i):
        x = dev_in[i % NBUF]
        if world == 1:
            return tree.enter(x)
        return shard_fn(tree, x, n)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks_true(flag):
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(fn, steps):
        """K calls bracketed by barrier + synchronize on both sides, CUDA events, max over ranks -> ms per call"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    # ---- timed region: K steps, device-resident input and output --------------------------------
    launches0 = L.ecfft_launch_count()
    ms_per_step = timed(step, args.steps)
    launches = L.ecfft_launch_count() - launches0
    value = n / (ms_per_step * 1e-3)
    extra = {}

    # ---- N > 1: the sharded result against a single-GPU ENTER of the whole vector, on every rank ----
    if world > 1:
        full_x = host_in[0].to(dev)
        want = tree.enter(full_x)
        got = shard_fn(tree, dev_in[0], n)
        ok = bool((got == want).all())
        del full_x, want, got
        extra["multi_gpu_matches_single"] = all_ranks_true(ok)
        if args.multi_gpu != "allgather":   # the final all-gather's share of the step
            ms_nogather = timed(lambda i: shard_fn(tree, dev_in[i % NBUF], n, gather=False), args.steps)
            extra["allgather_ms_per_step"] = ms_per_step - ms_nogather
            extra["ms_per_step_without_final_allgather"] = ms_nogather

    # ---- roofline of the dominant kernel: same steps with per-launch CUDA events ---------------------
    L.ecfft_profile_enable(1)
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    L.ecfft_profile_enable(0)
    prof = {}
    for kid, name in ((0, "k_extend_tile"), (1, "k_enter_combine")):
        kms, kb, kn = ctypes.c_double(), ctypes.c_double(), ctypes.c_ulonglong()
        _lib.check(L.ecfft_profile_read(kid, ctypes.byref(kms), ctypes.byref(kb), ctypes.byref(kn)))
        prof[name] = {"ms_per_step": kms.value / args.steps, "alg_gb_per_step": kb.value / args.steps / 1e9,
                      "launches_per_step": kn.value / args.steps}
    peak, peak_src = measured_peak()
    dom = prof["k_extend_tile"]
    achieved = dom["alg_gb_per_step"] / (dom["ms_per_step"] * 1e-3) if dom["ms_per_step"] > 0 else 0.0

    # ---- e2e: the host-buffer C ABI call (pinned host in -> H2D -> ENTER -> D2H -> pinned host out) ----
    h2d = d2h = 0
    if world == 1:
        host_np = [h.numpy().view(np.uint64) for h in host_in]
        host_out = torch.empty((n, 4), dtype=torch.int64).pin_memory()
        out_np = host_out.numpy().view(np.uint64)
        out_ptr = out_np.ctypes.data_as(ctypes.c_void_p)
        in_ptrs = [a.ctypes.data_as(ctypes.c_void_p) for a in host_np]
        for i in range(2):
            _lib.check(L.ecfft_enter(tree._h, in_ptrs[i % NBUF], n, out_ptr))
        t0 = time.perf_counter()
        for i in range(args.steps):
            _lib.check(L.ecfft_enter(tree._h, in_ptrs[i % NBUF], n, out_ptr))
        e2e_s = (time.perf_counter() - t0) / args.steps
        h2d = d2h = n * 32
        last_e2e_in = (args.steps - 1) % NBUF
        # the same through ecfft_enter_many: a caller's loop over NBUF polynomials as one call, so that the copies of
        # neighbouring vectors overlap the kernels (reported beside the single-call figure, never instead of it)
        many_in = torch.empty((NBUF, n, 4), dtype=torch.int64).pin_memory()
        for i in range(NBUF):
            many_in[i].copy_(host_in[i])
        many_out = torch.empty((NBUF, n, 4), dtype=torch.int64).pin_memory()
        mi, mo = ctypes.c_void_p(many_in.data_ptr()), ctypes.c_void_p(many_out.data_ptr())
        _lib.check(L.ecfft_enter_many(tree._h, mi, n, NBUF, mo))
        reps_many = max(1, args.steps // NBUF)
        t0 = time.perf_counter()
        for _ in range(reps_many):
            _lib.check(L.ecfft_enter_many(tree._h, mi, n, NBUF, mo))
        extra["e2e_many_ms_per_vector"] = (time.perf_counter() - t0) / (reps_many * NBUF) * 1e3
        extra["e2e_many_vectors_per_call"] = NBUF
        extra["e2e_many_matches_single_call"] = bool((many_out[last_e2e_in] == host_out).all())
        del many_in, many_out
    else:
        # every rank uploads its n/N coefficients from pinned memory, the ranks run the sharded ENTER and gather,
        # and RANK 0 downloads the WHOLE evaluation vector into its pinned memory — one consumer process ends up
        # with the whole Vec<F>, as the caller of the reference does.  Variants: the host-side result stays
        # sharded (e2e_sharded_ms_per_step), every rank downloads the whole vector (e2e_all_ranks_full_ms_per_step).
        host_chunks = [h[rank * chunk:(rank + 1) * chunk] for h in host_in]

        def e2e_loop(gather, everyone=False):
            download = gather and (everyone or rank == 0)
            host_out = torch.empty((n if gather else chunk, 4), dtype=torch.int64).pin_memory()
            if args.multi_gpu == "allgather":
                fn = (lambda x_: enter_sharded_allgather(tree, x_, n)) if gather else (lambda x_: enter_sharded_allgather(tree, x_, n)[rank * chunk:(rank + 1) * chunk])
            else:
                fn = lambda x_: shard_fn(tree, x_, n, gather=gather)
            for i in range(2):
                res = fn(host_chunks[i % NBUF].to(dev, non_blocking=True))
                if download or not gather:
                    host_out.copy_(res, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                xd = host_chunks[i % NBUF].to(dev, non_blocking=True)
                res = fn(xd)
                if download or not gather:
                    host_out.copy_(res, non_blocking=True)
                torch.cuda.synchronize()
                if gather and world > 1:
                    dist.barrier()      # a step ends when rank 0 holds the whole result
            barrier()
            return max_over_ranks((time.perf_counter() - t0) / args.steps)

        e2e_s = e2e_loop(True)
        extra["e2e_sharded_ms_per_step"] = e2e_loop(False) * 1e3
        extra["e2e_all_ranks_full_ms_per_step"] = e2e_loop(True, everyone=True) * 1e3
        h2d, d2h = chunk * 32, n * 32
    clocks = sampler.stop()

    products = products_per_elem(log_n) * n
    whole_alg_gb = alg_bytes_per_elem(log_n) * n / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 integer limbs (mod p, exact); API layout u64x4 Montgomery", "data": "synthetic",
        "config": {
            "workload": f"secp256k1::Fp ENTER n=2^{log_n} (full log^2 recursion) on a 2^{log_n}-leaf FFTree",
            "parallelism": "single GPU" if world == 1 else (
                f"{world} ranks: local ENTER(n/{world}) + 1 NCCL all-gather + top {world.bit_length() - 1} depths replicated" if args.multi_gpu == "allgather"
                else f"{world} ranks: local ENTER(n/{world}), top {world.bit_length() - 1} depths sharded (pairwise NCCL send/recv per straddling level), final all-gather of the result" if args.multi_gpu == "sharded"
                else f"{world} ranks: local ENTER(n/{world}), top {world.bit_length() - 1} depths sharded, straddling butterfly levels and combines read the partner's buffers over NVLink (CUDA-IPC peer memory, flag-ordered), final NCCL all-gather of the result"),
            "inputs": f"{NBUF} rotating coefficient vectors of {n * 32 >> 20} MiB; tables ~450 B/leaf resident in HBM; no explicit L2 flush (per-step stream >> 126 MB L2)",
            "tree_build_s": round(t_build, 3),
            "host": numa,
        },
        "e2e": {"value": n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3,
                "path": ("ecfft_enter (host-buffer C ABI): pinned host coefficients -> H2D -> ENTER -> D2H -> pinned host evaluations" if world == 1
                         else "every rank: pinned host chunk (n/N coefficients) -> H2D -> sharded ENTER -> all-gather; rank 0: D2H of the WHOLE evaluation vector into its pinned memory (h2d bytes per rank, d2h bytes on rank 0)")},
        "gpu_launches": int(launches),
        "roofline": {
            "kernel": "k_extend_sym", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": ncu_traffic(log_n)[0], "traffic_unit": ncu_traffic(log_n)[1],
            "peak_source": peak_src,
            "note": ("algorithmic bytes of the level-streaming model (SURVEY 8d); the kernel fuses ~10 levels and the combine "
                     "per HBM round trip, so achieved exceeds the physical peak while `traffic` (ncu DRAM bytes) is ~9x "
                     "smaller: the kernel is bound by the integer multiplier, see integer_pipe_frac"),
            "kernel_ms_per_step": dom["ms_per_step"], "alg_gb_per_step": dom["alg_gb_per_step"],
            "launches_per_step": dom["launches_per_step"],
            # flat scalars (nested objects do not survive the driver's parser)
            "whole_step_alg_gb": whole_alg_gb,
            "whole_step_achieved_gbs": whole_alg_gb / (ms_per_step * 1e-3),
            "whole_step_frac": whole_alg_gb / (ms_per_step * 1e-3) / peak,
            "modmul_per_s_reference_count": modmuls_per_elem(log_n) * n / (ms_per_step * 1e-3),
            # what actually binds the kernel: the 32x32->64 integer multiplier
            "integer_pipe_products_per_step": products,
            "integer_pipe_products_per_s": products / (ms_per_step * 1e-3),
            "integer_pipe_frac": products / (ms_per_step * 1e-3) / PRODUCT_PEAK / world,
            "integer_pipe_peak_products_per_s_per_gpu": PRODUCT_PEAK,
            "integer_pipe_peak_source": "tools/microbench.cu on B200: 108 G register-resident 256-bit products/s (IMAD.WIDE.X 9.1 T/s)",
            "combine_kernel_ms_per_step": prof["k_enter_combine"]["ms_per_step"],
            "combine_kernel_launches_per_step": prof["k_enter_combine"]["launches_per_step"],
        },
        "clocks": clocks,
    }
    line.update(extra)

    ok_all = True
    if not args.no_configs:
        cfg = run_configs(args, tree, world, rank, local_rank, dev, timed, all_ranks_true)
        line.update(cfg)
        ok_all = ok_all and all(v for k, v in cfg.items() if k.endswith("_matches_single") or k.endswith("_roundtrip_ok") or k.endswith("_status_ok"))
    if "e2e_many_matches_single_call" in line:
        ok_all = ok_all and line["e2e_many_matches_single_call"]
    if "multi_gpu_matches_single" in line:
        ok_all = ok_all and line["multi_gpu_matches_single"]

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = cpu_sample(min(args.cpu_sample_log_n, log_n), log_n, threads=1)
        # the same sample on the GPU must agree bit for bit
        got = tree.enter(s["x"])
        same = bool((got == s["out"]).all())
        ok_all = ok_all and same
        line["cpu_baseline"] = {
            "value": s["evals_per_s"], "unit": UNIT, "cores": 1, "kind": "port",
            "sample": (f"oracle ENTER n=2^{min(args.cpu_sample_log_n, log_n)} single thread took {s['seconds']:.2f} s; scaled by "
                       f"field-multiplication count to n=2^{log_n} ({s['est_seconds_full']:.1f} s); GPU output on the sample "
                       f"{'bit-identical' if same else 'DIFFERS'}; host has {os.cpu_count()} logical cores"),
        }
        # and the last e2e result against nothing but itself run on device (same input): consistency of both paths
        chk = tree.enter(host_np[last_e2e_in])
        line["e2e"]["matches_device_path"] = bool((chk == out_np).all())
        ok_all = ok_all and line["e2e"]["matches_device_path"]
    line["self_check_ok"] = bool(ok_all)
    if rank == 0:
        print_json(line)
    if world > 1:
        if arena is not None:
            dist.barrier()          # nobody reads a peer arena any more
            arena.close()
        dist.destroy_process_group()
    if not ok_all:
        raise SystemExit("bench.py: a self-check failed (see *_matches_single / matches_device_path / cpu_baseline.sample)")


def run_configs(args, tree22, world, rank, local_rank, dev, timed, all_ranks_true):
    """The other BASELINE.json configurations as flat cfg_* keys (reference benches/fftree.rs:19-62 times the same
    algorithms).  Device-resident buffers, CUDA events, median of 5 samples after 2 warm-ups."""
    import numpy as np
    import torch
    import ecfft_b200
    from oracle import oracle as O
    from ecfft_b200.dist import PeerArena, enter_sharded_peer
    out = {}

    def median_ms(fn, warm=2, reps=5):
        for _ in range(warm):
            fn()
        return statistics.median(timed(lambda i: fn(), 1) for _ in range(reps))

    def to_dev(a):
        return torch.from_numpy(a.view(np.int64)).to(dev)

    if world == 1:
        del tree22  # the headline tree carries only ENTER's tables; these configurations need the whole FFTree
        t0 = time.perf_counter()
        tree = ecfft_b200.build_fftree(1 << args.log_n, device=local_rank)
        torch.cuda.synchronize()
        out["cfg_full_tree_build_s"] = round(time.perf_counter() - t0, 3)
        # configs[0]: ENTER -> EXIT round trip at n = 2^12
        x12 = to_dev(O.random_elements(1 << 12, seed=11))
        out["cfg_roundtrip_2p12_ms"] = median_ms(lambda: tree.exit(tree.enter(x12)))
        out["cfg_roundtrip_2p12_roundtrip_ok"] = bool((tree.exit(tree.enter(x12)) == x12).all())
        # configs[1]: EXTEND n = 2^20 (S0 -> S1 on the 2^21-leaf subtree)
        x20 = to_dev(O.random_elements(1 << 20, seed=12))
        ms = median_ms(lambda: tree.extend(x20, 1))
        out["cfg_extend_2p20_ms"] = ms
        out["cfg_extend_2p20_elems_per_s"] = (1 << 20) / (ms * 1e-3)
        out["cfg_extend_2p20_roundtrip_ok"] = bool((tree.extend(tree.extend(x20, 1), 0) == x20).all())
        # configs[4]: REDC + MOD n = 2^20 with matching-size tables
        xnn = to_dev(tree.table("xnn_s", 1 << 20))
        zz = to_dev(tree.table("z0z0_rem_xnn_s", 1 << 20))
        out["cfg_redc_2p20_ms"] = median_ms(lambda: tree.redc_z0(x20, xnn))
        out["cfg_mod_2p20_ms"] = median_ms(lambda: tree.modular_reduce(x20, xnn, zz))
        # EXIT at the headline size (reference benches/fftree.rs:32-34) and EXIT(ENTER(x)) = x there
        xn = to_dev(O.random_elements(1 << args.log_n, seed=13))
        ev = tree.enter(xn)
        ms = median_ms(lambda: tree.exit(ev), warm=1, reps=3)
        out[f"cfg_exit_2p{args.log_n}_ms"] = ms
        out[f"cfg_exit_2p{args.log_n}_elems_per_s"] = (1 << args.log_n) / (ms * 1e-3)
        out[f"cfg_exit_2p{args.log_n}_roundtrip_ok"] = bool((tree.exit(ev) == xn).all())
        del tree, xn, ev
        torch.cuda.empty_cache()
        # the reference's second field (src/lib.rs:190-215): FFTree<m31::Fp>, ENTER and EXIT at n = 2^20
        t0 = time.perf_counter()
        t31 = ecfft_b200.m31.build_fftree(1 << 20, device=local_rank)
        torch.cuda.synchronize()
        out["cfg_m31_tree_build_2p20_s"] = round(time.perf_counter() - t0, 3)
        x31 = torch.randint(0, (1 << 31) - 1, (1 << 20,), dtype=torch.int32, device=dev)
        out["cfg_m31_enter_2p20_ms"] = median_ms(lambda: t31.enter(x31))
        e31 = t31.enter(x31)
        out["cfg_m31_exit_2p20_ms"] = median_ms(lambda: t31.exit(e31), warm=1, reps=3)
        out["cfg_m31_2p20_roundtrip_ok"] = bool((t31.exit(e31) == x31).all())
        return out

    # N > 1 — configs[3]: ENTER n = 2^24 sharded over the N GPUs (peer schedule + final all-gather)
    log_big = 24
    nb = 1 << log_big
    tree = ecfft_b200.build_fftree(nb, parts=ecfft_b200.PARTS_ENTER_ONLY, device=local_rank)
    full = to_dev(O.random_elements(nb, seed=21))
    c = nb // world
    mine = full[rank * c:(rank + 1) * c].contiguous()
    arena = PeerArena.create(nb, local_rank)
    ms = median_ms(lambda: enter_sharded_peer(tree, mine, nb, arena), warm=2, reps=5)
    out["cfg_2p24_ms"] = ms
    out["cfg_2p24_evals_s"] = nb / (ms * 1e-3)
    out["cfg_2p24_ms_without_final_allgather"] = median_ms(lambda: enter_sharded_peer(tree, mine, nb, arena, gather=False), warm=1, reps=5)
    got = enter_sharded_peer(tree, mine, nb, arena)
    want = tree.enter(full)
    out["cfg_2p24_matches_single"] = all_ranks_true(bool((got == want).all()))
    out["cfg_2p24_single_gpu_ms"] = median_ms(lambda: tree.enter(full), warm=1, reps=3)
    import torch.distributed as dist
    dist.barrier()
    arena.close()
    del tree, full, mine, got, want
    torch.cuda.empty_cache()

    # sharded EXIT at the headline size (reference src/fftree.rs:200-224): MOD across the ranks for the top
    # log2(N) depths, then an independent EXIT(n/N) per rank; checked against a single-GPU EXIT on every rank
    from ecfft_b200.dist import exit_sharded_peer
    ne = 1 << args.log_n
    tree = ecfft_b200.build_fftree(ne, device=local_rank)
    ev = to_dev(O.random_elements(ne, seed=22))
    ce = ne // world
    mine = ev[rank * ce:(rank + 1) * ce].contiguous()
    arena = PeerArena.create(ne, local_rank)
    out[f"cfg_exit_sharded_2p{args.log_n}_ms"] = median_ms(lambda: exit_sharded_peer(tree, mine, ne, arena, gather=False), warm=2, reps=5)
    got = exit_sharded_peer(tree, mine, ne, arena)
    want = tree.exit(ev)
    out[f"cfg_exit_sharded_2p{args.log_n}_matches_single"] = all_ranks_true(bool((got == want).all()))
    out[f"cfg_exit_single_gpu_2p{args.log_n}_ms"] = median_ms(lambda: tree.exit(ev), warm=1, reps=3)
    out[f"cfg_exit_sharded_2p{args.log_n}_arena_status_ok"] = all_ranks_true(arena.status() == 0)
    dist.barrier()
    arena.close()
    return out


if __name__ == "__main__":
    main()
