#!/usr/bin/env python
"""bench.py -- Chamfer+DCD fwd+bwd throughput (Gpair/s) of the B200 path, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step is one pass of the hot path over one synthetic batch: calc_dcd(x, gt) forward and
loss.sum().backward() (pack -> nn_kernel -> dcd_fwd_kernel -> grad kernels).  Work per step is
counted the reference's way: 2*B*N*M ordered pair evaluations, 8 FLOP each,
whatever the kernel does internally.  The default workload is BASELINE.json configs[1]
("chair retrieval": 64 targets x K=10 deformed candidates, 2048 points each = 640 pairs).

Multi-GPU (weak scaling): every rank scores its own 64 queries x 10 candidates -- independent
units, no data-path collective; the JSON line reports the aggregate over all ranks and the max
step time over ranks.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Chamfer+DCD fwd+bwd Gpair/s"
UNIT = "Gpair/s"
FLOP_PER_PAIR = 8.0  # 3 sub, 3 mul, 2 add: the reference arithmetic

RETRIEVAL = {
    # name: (library shapes S, queries Q, points, description)   -- sharded over ranks, strong scaling
    "cfg3": (1000, 1, 2048, "library retrieval: 1 target vs S=1000 shapes x 2048 pts, top-10, library sharded over ranks (BASELINE configs[2])"),
    "cfg5": (10000, 64, 2048, "retrieval sweep point: 64 targets vs S shapes x 2048 pts, top-10, sharded (BASELINE configs[4]; --library-size)"),
}
WORKLOADS = {
    # name: (pairs B, n_x, n_gt, description)
    "cfg2": (640, 2048, 2048, "chair retrieval: 64 targets x K=10 deformed candidates, 2048 pts (BASELINE configs[1])"),
    "cfg1": (32, 2048, 2048, "chair: batch 32 targets vs deformed sources, 2048 pts (BASELINE configs[0])"),
    "cfg4": (16, 16384, 16384, "dense clouds: batch 16, 16384 x 16384 pts (BASELINE configs[3])"),
}
ALPHA, N_LAMBDA = 1000, 1  # compute_dcd_loss defaults (engine/geometry_utils.py:80-82)


def synth(B, n_x, n_gt, seed):
    """Shape-like synthetic clouds: unit-ball targets, candidates = anisotropically scaled noisy resamples."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    gt = torch.randn(B, n_gt, 3, generator=g)
    gt = gt - gt.mean(1, keepdim=True)
    gt = gt / gt.norm(dim=2).amax(1).view(B, 1, 1)
    x = torch.randn(B, n_x, 3, generator=g)
    x = x - x.mean(1, keepdim=True)
    x = x / x.norm(dim=2).amax(1).view(B, 1, 1)
    x = x * (1 + 0.1 * torch.rand(B, 1, 3, generator=g)) + 0.01 * torch.randn(B, n_x, 3, generator=g)
    return x.contiguous(), gt.contiguous()


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML): sm clock + throttle reasons during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's torch CPU path (oracle port), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_step(x, gt):
    """chamfer_python.distChamfer + calc_dcd torch body, fwd + bwd, on the host cores."""
    from oracle import torch_path
    xs, gts = x.clone().requires_grad_(), gt.clone().requires_grad_()
    loss, _, _ = torch_path.calc_dcd_oracle(xs, gts, alpha=ALPHA, n_lambda=N_LAMBDA, chamfer=torch_path.dist_chamfer_cpu)
    loss.sum().backward()
    return float(loss.detach().sum())


def cpu_baseline(n_x, n_gt, sample_pairs, reps, warmup):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, gt = synth(sample_pairs, n_x, n_gt, seed=1234)
    for _ in range(warmup):
        cpu_step(x, gt)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_step(x, gt)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    pairs = 2.0 * sample_pairs * n_x * n_gt
    return {"value": pairs / med / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_pairs} pairs of {n_x}x{n_gt} pts per step (same clouds/alpha/lambda as the GPU workload), "
                      f"median of {reps} after {warmup} warm-up; oracle/torch_path.py restating chamfer_python.py:18-39 + model_utils.py:13-51",
            "sec_per_step": med}


def run_reference_arm(args, wl):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, n_x, n_gt, desc = WORKLOADS[wl]
    sample_pairs = int(os.environ.get("URED_BENCH_CPU_SAMPLE", "0")) or (32 if n_x * n_gt <= 2048 * 2048 else 1)
    steps = max(1, min(args.steps, 5))       # bounded: each step is ~1-3 s of CPU work on 8-16 cores
    warmup = max(1, min(args.warmup, 1))
    base = cpu_baseline(n_x, n_gt, sample_pairs, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": base["sec_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{wl}: {desc}", "pairs_per_step_sampled": sample_pairs, "n_x": n_x, "n_gt": n_gt,
                   "alpha": ALPHA, "n_lambda": N_LAMBDA},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------
# sharded library retrieval (cfg3 / cfg5): forward scoring + local top-k + one all_gather + merge
# ------------------------------------------------------------------------------------------------
def run_retrieval(args):
    import torch
    import torch.distributed as dist
    import ured_b200 as ured

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    lib = ured._native.load()
    S, Q, n, desc = RETRIEVAL[args.workload]
    S = args.library_size or S
    Q = args.queries or Q
    k = 10
    lo, hi = ured.shard_bounds(S, world, rank)
    # the library is defined in global slabs of 512 shapes (seed = slab id), so every world size sees the SAME
    # library; a rank generates the slabs overlapping its shard, packs the shard once and keeps it resident
    SLAB = 512
    slabs = []
    for sid in range(lo // SLAB, (hi + SLAB - 1) // SLAB if hi > lo else 0):
        x, _ = synth(SLAB, n, 8, seed=7000 + sid)
        a, b = max(lo, sid * SLAB) - sid * SLAB, min(hi, (sid + 1) * SLAB) - sid * SLAB
        slabs.append(x[a:b].to(dev))
    shard = ured.PackedClouds(torch.cat(slabs)) if slabs else None
    del slabs
    _, tg_host = synth(Q, 8, n, seed=99)      # same targets on every rank
    tg_pin = tg_host.pin_memory()
    tg_dev = tg_pin.to(dev)
    out_s = torch.empty(Q, k).pin_memory()
    out_i = torch.empty(Q, k, dtype=torch.int32).pin_memory()
    pairs_per_step = 2.0 * Q * S * n * n

    # one rank: CUDA-graph engine by default; several ranks: eager scoring + one all_gather (measured faster at 8 GPUs
    # for single-query steps), graph incl. the collective on request
    use_graph = (not args.eager) and (world == 1 or args.graph)
    args.eager = not use_graph
    engine = ured.RetrievalEngine(shard, lo, Q, k=k, metric="cd_t", use_graph=use_graph)

    def step_device():
        return engine.query(tg_dev)

    def step_e2e():
        t = tg_pin.to(dev, non_blocking=True)
        v, i = engine.query(t)
        out_s.copy_(v, non_blocking=True)
        out_i.copy_(i, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sampler:
            sampler.start()
        l0 = lib.ured_kernel_launches()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        launches = lib.ured_kernel_launches() - l0
        if sampler:
            sampler.stop()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, launches

    sampler = ClockSampler(local_rank)
    ms_step, launches = timed(step_device, args.steps, args.warmup, sampler)
    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    v, i = step_device()
    torch.cuda.synchronize()
    line = {
        "metric": "Chamfer+DCD fwd Gpair/s (sharded retrieval, top-k merged)", "value": pairs_per_step / (ms_step * 1e-3) / 1e9, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "library_shapes": S, "queries": Q, "points": n, "top_k": k,
                   "shard": [lo, hi], "engine": "eager" if args.eager else ("cuda-graph" + ("" if world == 1 else " incl. all_gather")),
                   "collective": "one all_gather of [Q,k] (score,id) pairs per step" if world > 1 else "none (1 rank)",
                   "l2": "library shard (%.0f MB packed + raw) exceeds L2 except at the smallest sizes" % ((hi - lo) * n * 28 / 1e6)},
        "e2e": {"value": pairs_per_step / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(tg_pin.numel() * 4), "d2h_bytes_per_step": int(Q * k * 8)},
        "gpu_launches": int(launches) if args.eager else int(engine.kernels_per_replay * args.steps), "clocks": sampler.summary(),
        "top1": {"score": float(v[0, 0]), "id": int(i[0, 0])},
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()

# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + sorted(RETRIEVAL))
    ap.add_argument("--library-size", type=int, default=0, help="override S for the retrieval workloads")
    ap.add_argument("--queries", type=int, default=0, help="override Q for the retrieval workloads")
    ap.add_argument("--eager", action="store_true", help="retrieval workloads: eager calls instead of the CUDA-graph engine")
    ap.add_argument("--graph", action="store_true", help="retrieval workloads on >1 rank: capture scoring + all_gather + merge in one CUDA graph")
    ap.add_argument("--exact-only", action="store_true", help="disable the screening pass (difference form on every pair)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args, args.workload if args.workload in WORKLOADS else "cfg2")
        return
    if args.workload in RETRIEVAL:
        run_retrieval(args)
        return

    import torch
    import torch.distributed as dist
    import ured_b200 as ured

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = ured._native.load()
    if args.exact_only:
        os.environ["URED_EXACT_ONLY"] = "1"

    B, n_x, n_gt, desc = WORKLOADS[args.workload]
    pairs_per_step = 2.0 * B * n_x * n_gt
    x_host, gt_host = synth(B, n_x, n_gt, seed=100 + rank)
    if args.workload == "cfg2":  # one target per K=10 candidates: both arms see the same broadcast targets
        gt_host = gt_host[::10].repeat_interleave(10, dim=0).contiguous()
    x_pin, gt_pin = x_host.pin_memory(), gt_host.pin_memory()
    x_dev, gt_dev = x_pin.to(dev), gt_pin.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        x = x_dev.detach().requires_grad_()
        gt = gt_dev.detach().requires_grad_()
        loss, _cd_p, _cd_t = ured.calc_dcd(x, gt, alpha=ALPHA, n_lambda=N_LAMBDA)
        loss.sum().backward()
        return loss, x.grad, gt.grad

    # ---- end to end: host buffers in, host result out, every step ---------------------------------
    # The retrieval workload's natural host-side form: Q=64 targets [Q,N,3] and their K=10 candidates
    # [Q*K,M,3] sit in pinned host memory; each step copies BOTH to the device (copy stream, double
    # buffered so that step i+1's upload overlaps step i's kernels, as a pinned-memory data loader does),
    # broadcasts each target over its K candidates on the device, runs calc_dcd fwd+bwd and copies the
    # per-pair loss back to pinned host memory.
    K_CAND = 10 if B % 10 == 0 else 1
    gt_small_pin = gt_host[::K_CAND].contiguous().pin_memory()
    loss_host = torch.empty(B, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [{"x": torch.empty_like(x_dev), "gt": torch.empty(B // K_CAND, n_gt, 3, device=dev), "ready": torch.cuda.Event(),
              "free": torch.cuda.Event()} for _ in range(2)]
    state = {"i": 0, "primed": False}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot["free"])          # the step that last used this slot has finished reading it
            slot["x"].copy_(x_pin, non_blocking=True)
            slot["gt"].copy_(gt_small_pin, non_blocking=True)
            slot["ready"].record(copy_stream)

    def step_e2e():
        cur = torch.cuda.current_stream(dev)
        if not state["primed"]:
            for sl in slots:
                sl["free"].record(cur)
            upload(slots[0])
            state["primed"] = True
        slot = slots[state["i"] % 2]
        upload(slots[(state["i"] + 1) % 2])               # prefetch the next step's inputs
        cur.wait_event(slot["ready"])
        x = slot["x"].detach().requires_grad_()
        gt = slot["gt"].detach().repeat_interleave(K_CAND, dim=0).requires_grad_()
        loss, _cd_p, _cd_t = ured.calc_dcd(x, gt, alpha=ALPHA, n_lambda=N_LAMBDA)
        loss.sum().backward()
        loss_host.copy_(loss.detach(), non_blocking=True)
        slot["free"].record(cur)
        state["i"] += 1
        return x.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None, whole_loop=False):
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(1 if whole_loop else steps)]
        barrier()
        if sampler:
            sampler.start()
        launches0 = lib.ured_kernel_launches()
        if whole_loop:
            # copies for step i+1 overlap step i, so the loop is bracketed once (no untimed gaps to hide work in);
            # every step's inputs arrive fresh from the host, so there is no L2 flush here
            ev[0][0].record()
            for s in range(steps):
                fn()
            ev[0][1].record()
        else:
            for s in range(steps):
                flush.zero_()          # evict L2 between timed iterations (outside the event pair)
                ev[s][0].record()
                fn()
                ev[s][1].record()
        barrier()
        launches = lib.ured_kernel_launches() - launches0
        if sampler:
            sampler.stop()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, launches

    sampler = ClockSampler(local_rank)
    ms_step, launches = timed(step_device, args.steps, args.warmup, sampler)
    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup, whole_loop=True)
    torch.cuda.synchronize()

    # ---- dominant kernel alone: nn_kernel (both directions, one launch) on packed clouds ---------
    pk_gt, pk_x = ured.PackedClouds(gt_dev), ured.PackedClouds(x_dev)
    d1 = torch.empty(B, n_gt, device=dev); d2 = torch.empty(B, n_x, device=dev)
    i1 = torch.empty(B, n_gt, device=dev, dtype=torch.int32); i2 = torch.empty(B, n_x, device=dev, dtype=torch.int32)
    stream = torch.cuda.current_stream(dev).cuda_stream
    flags = 1 if args.exact_only else 0

    scratch_bytes = lib.ured_nn_scratch_bytes(B, n_gt, n_x)
    scratch = torch.empty(max(scratch_bytes, 256), dtype=torch.uint8, device=dev)

    def nn_only():
        rc = lib.ured_nn_packed(gt_dev.data_ptr(), pk_gt.packed.data_ptr(), n_gt,
                                x_dev.data_ptr(), pk_x.packed.data_ptr(), n_x, B, 1, B, None, None,
                                d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                                scratch.data_ptr(), scratch_bytes, flags, stream)
        ured._native.check(rc, "ured_nn_packed")

    ms_nn, _ = timed(nn_only, args.steps, args.warmup)
    flop_per_launch = FLOP_PER_PAIR * pairs_per_step
    achieved_tflops = flop_per_launch / (ms_nn * 1e-3) / 1e12

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_max_mhz = float(peaks.get("sm_max_mhz") or sampler.max_mhz or 1965.0)
    peak_tflops = sms * 128 * 2 * sm_max_mhz * 1e6 / 1e12  # FP32 FMA lanes x 2 FLOP x max SM clock
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "nn_kernel_traffic.json"))).get(args.workload)
    except Exception:
        pass

    line = {
        "metric": METRIC, "value": world * pairs_per_step / (ms_step * 1e-3) / 1e9, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "pairs_per_gpu": B, "n_x": n_x, "n_gt": n_gt, "alpha": ALPHA,
                   "n_lambda": N_LAMBDA, "step": "calc_dcd fwd + loss.sum().backward()", "sharding": "independent pairs per rank, no collective",
                   "kernel_variant": "exact-only" if args.exact_only else "screen+exact-recheck",
                   "l2": "256 MiB buffer rewritten between timed iterations"},
        "e2e": {"value": world * pairs_per_step / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(x_pin.numel() * 4 + gt_small_pin.numel() * 4), "d2h_bytes_per_step": int(B * 4),
                "api": "pinned-host candidates + targets copied every step (copy stream, double-buffered), targets broadcast on device, "
                       "calc_dcd fwd+bwd, per-pair loss copied back to pinned host"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp32_fma", "kernel": "nn_kernel (both directions, one launch)", "achieved": achieved_tflops,
                     "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops, "traffic": traffic,
                     "peak_source": f"{sms} SMs x 128 FP32 lanes x 2 FLOP x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz; it has no FP32 figure); "
                                    "tools/microbench/fp32_peak.cu measured 72.6 (FFMA) / 74.3 (FFMA2) TFLOP/s on this pool",
                     "frac_of_measured_ffma_peak": achieved_tflops / 72.6,  # profiles/r01_microbench_fp32_peak.txt (scalar FFMA stream)
                     "flop_per_launch": flop_per_launch, "kernel_ms": ms_nn,
                     "tpair_per_s": pairs_per_step / (ms_nn * 1e-3) / 1e12,
                     "note": "FLOP-accounted at the reference's 8 FLOP per ordered pair; the screening variant executes 6 FLOP per pair in its main loop"},
        "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_pairs = 32 if n_x * n_gt <= 2048 * 2048 else 1
        line["cpu_baseline"] = {k: v for k, v in cpu_baseline(n_x, n_gt, sample_pairs, 3, 1).items() if k != "sec_per_step"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
