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
TMEM_READ_BYTES_PER_CLK = 327.7  # per SM, measured: tools/microbench/tmem_read.cu (profiles/r02_tmem_read.txt)

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
# shared helpers of the GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of one bench run (device, ranks, library handle)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import ured_b200 as ured
        self.torch, self.dist, self.ured = torch, dist, ured
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # rank 0 prints exactly one JSON line on stdout: NCCL's version banner (NCCL_DEBUG=VERSION and above) must not join it
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ.pop("NCCL_DEBUG", None)
            else:
                os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = ured._native.load()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup, sampler=None, whole_loop=False, flush=True, collective=True, finish=None):
        """W untimed warm-up steps, then EXACTLY `steps` timed ones, bracketed by barrier + synchronize, CUDA events on
        the launching stream, max over ranks.  flush: rewrite a 256 MiB buffer between timed iterations (outside the
        event pairs); whole_loop: one event pair around all steps (pipelined e2e loops).  collective=False: this rank
        alone is measuring (no barrier, no max).  finish: called after the last step, before the end event (whole_loop only) --
        work submitted to other streams is joined into the timed region there."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        if finish:
            finish()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(1 if whole_loop else steps)]
        self.barrier() if collective else torch.cuda.synchronize()
        if sampler:
            sampler.start()
        launches0 = self.lib.ured_kernel_launches()
        if whole_loop:
            ev[0][0].record()
            for _ in range(steps):
                fn()
            if finish:
                finish()
            ev[0][1].record()
        else:
            for s in range(steps):
                if flush:
                    self.flush.zero_()
                ev[s][0].record()
                fn()
                ev[s][1].record()
        self.barrier() if collective else torch.cuda.synchronize()
        launches = self.lib.ured_kernel_launches() - launches0
        if sampler:
            sampler.stop()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        ms = total_ms / steps
        return (self.max_over_ranks(ms) if collective else ms), int(launches)


def measure_ffma_peak(ctx):
    """FP32 FMA peak of THIS device, measured now: a pure FFMA stream (ured_probe_ffma), best of 5, CUDA events."""
    import ctypes
    torch = ctx.torch
    sms = torch.cuda.get_device_properties(ctx.dev).multi_processor_count
    sink = torch.zeros(16, device=ctx.dev)
    flop = ctypes.c_double()
    stream = torch.cuda.current_stream(ctx.dev).cuda_stream
    best = None
    for it in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.ured._native.check(ctx.lib.ured_probe_ffma(sink.data_ptr(), sms * 8, 4096, ctypes.byref(flop), stream), "ured_probe_ffma")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if it >= 2:
            best = ms if best is None else min(best, ms)
    return flop.value / (best * 1e-3) / 1e12


def nn_roofline(ctx, x_dev, gt_dev, B, n_x, n_gt, steps, warmup, exact_only, peaks, ffma_peak, workload):
    """The dominant kernel alone (nn_tc_kernel, or nn_kernel with --exact-only / URED_NN_TC=0): both directions, one launch, on
    packed clouds, L2 flushed between launches."""
    import ctypes
    torch, ured, lib, dev = ctx.torch, ctx.ured, ctx.lib, ctx.dev
    pk_gt, pk_x = ured.PackedClouds(gt_dev), ured.PackedClouds(x_dev)
    d1 = torch.empty(B, n_gt, device=dev); d2 = torch.empty(B, n_x, device=dev)
    i1 = torch.empty(B, n_gt, device=dev, dtype=torch.int32); i2 = torch.empty(B, n_x, device=dev, dtype=torch.int32)
    stream = torch.cuda.current_stream(dev).cuda_stream
    flags = 1 if exact_only else 0
    scratch_bytes = lib.ured_nn_scratch_bytes(B, n_gt, n_x)
    scratch = torch.empty(max(scratch_bytes, 256), dtype=torch.uint8, device=dev)

    def nn_only(fl=flags):
        rc = lib.ured_nn_packed(gt_dev.data_ptr(), pk_gt.packed.data_ptr(), n_gt,
                                x_dev.data_ptr(), pk_x.packed.data_ptr(), n_x, B, 1, B, None, None,
                                d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                                scratch.data_ptr(), scratch_bytes, fl, stream)
        ured._native.check(rc, "ured_nn_packed")

    ms_nn, _ = ctx.timed(nn_only, steps, warmup)
    v, q, t, ns, it, si = (ctypes.c_int() for _ in range(6))
    lib.ured_nn_launch_shape(B, n_gt, n_x, flags, ctypes.byref(v), ctypes.byref(q), ctypes.byref(t), ctypes.byref(ns), ctypes.byref(it), ctypes.byref(si))
    pairs = 2.0 * B * n_x * n_gt
    flop = FLOP_PER_PAIR * pairs
    achieved = flop / (ms_nn * 1e-3) / 1e12
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_max_mhz = float(peaks.get("sm_max_mhz") or 1965.0)
    peak = sms * 128 * 2 * sm_max_mhz * 1e6 / 1e12  # FP32 FMA lanes x 2 FLOP x max SM clock
    traffic, traffic_src = None, None
    tensor = v.value == 100                           # URED_NN_VARIANT_TENSOR: the screening pass runs on the tensor cores
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "nn_kernel_traffic.json")))
        sect = tj.get("nn_tc_kernel" if tensor else "nn_kernel", tj)
        traffic, traffic_src = sect.get(workload), sect.get("source")
    except Exception:
        pass
    shape = {"variant": v.value, "queries_per_cta": q.value, "threads": t.value, "work_items": it.value,
             "split_items": si.value, "candidate_splits": ns.value}
    fp32 = {"fp32_fma_peak": peak,
            "fp32_fma_peak_source": f"{sms} SMs x 128 FP32 lanes x 2 FLOP x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz; it has no FP32 figure)",
            "algorithmic_tflops": achieved, "frac_of_fp32_fma_peak": achieved / peak,
            "measured_ffma_peak": ffma_peak, "frac_of_measured_ffma_peak": (achieved / ffma_peak) if ffma_peak else None,
            "measured_ffma_peak_source": "ured_probe_ffma (pure FFMA stream) timed in this run, best of 5"}
    if tensor:
        # the screen is a [queries x 32] . [candidates x 32]^T bf16 product (27 exact piece products per pair, K padded to 32):
        # 64 executed tensor FLOP per ordered pair; every score is then read out of TMEM once (4 bytes per pair), which is what binds
        tflop = 2.0 * 32 * pairs
        t_ach = tflop / (ms_nn * 1e-3) / 1e12
        t_peak = float(peaks.get("bf16_tflops") or 2250.0)
        tmem_peak = sms * TMEM_READ_BYTES_PER_CLK * sm_max_mhz * 1e6 / 1e9
        tmem_ach = 4.0 * pairs / (ms_nn * 1e-3) / 1e9
        keep = (d1.clone(), d2.clone(), i1.clone(), i2.clone())
        ms_fp32, _ = ctx.timed(lambda: nn_only(flags | ured._native.URED_FLAG_FP32_SCREEN), max(3, steps // 4), 3)   # same data, FP32-pipe screen
        same = all(torch.equal(a, b) for a, b in zip(keep, (d1, d2, i1, i2)))
        if not same:
            raise SystemExit("bench: the tensor-core and FP32-pipe screening kernels disagree")
        return {"bound": "tensor", "kernel": "nn_tc_kernel (tcgen05 screening + exact re-check, both directions, one launch%s)" % (", %d candidate ranges per cloud" % ns.value if ns.value > 1 else ""),
                "achieved": t_ach, "peak": t_peak, "unit": "TFLOP/s", "frac": t_ach / t_peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS bf16 burst figure; the kernel is timed alone)" if peaks.get("bf16_tflops") else "B200_PROFILING.md fallback: 2250 TFLOP/s dense bf16 (MEASURED_PEAKS.json absent)",
                "flop_per_launch": tflop, "kernel_ms": ms_nn, "tpair_per_s": pairs / (ms_nn * 1e-3) / 1e12, "launch_shape": shape,
                "fp32_screen_kernel_ms": ms_fp32, "outputs_equal_fp32_screen": same,
                "tmem_read": {"bytes_per_launch": 4.0 * pairs, "achieved_gbs": tmem_ach, "peak_gbs": tmem_peak, "frac": tmem_ach / tmem_peak,
                              "peak_source": f"{TMEM_READ_BYTES_PER_CLK} B/clk/SM: tools/microbench/tmem_read.cu, eight warps draining a 128 x 256 fp32 accumulator "
                                             "with tcgen05.ld.32x32b.x32 and a running minimum (profiles/r02_tmem_read.txt)"},
                **fp32,
                "note": "achieved = EXECUTED tensor FLOP (2 x 32 per ordered pair: 27 exact bf16 piece products, K padded to 32) / nn_tc_kernel time; "
                        "the reference's accounting (8 FLOP per pair, SURVEY 8d) is algorithmic_tflops, which now exceeds the FP32-FMA peak because "
                        "the pair arithmetic left the FP32 pipes; the binding resource is the TMEM read port (tmem_read)"}
    return {"bound": "fp32_fma", "kernel": "nn_kernel (both directions, one launch%s)" % (", + merge of %d candidate splits" % ns.value if ns.value > 1 else ""),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": fp32["fp32_fma_peak_source"],
            "measured_ffma_peak": ffma_peak, "frac_of_measured_ffma_peak": (achieved / ffma_peak) if ffma_peak else None,
            "measured_ffma_peak_source": "ured_probe_ffma (pure FFMA stream) timed in this run, best of 5",
            "flop_per_launch": flop, "kernel_ms": ms_nn, "tpair_per_s": pairs / (ms_nn * 1e-3) / 1e12, "launch_shape": shape,
            "note": "FLOP-accounted at the reference's 8 FLOP per ordered pair; the screening variant executes 6 FLOP per pair in its main loop "
                    "(executed-FLOP fraction = 0.75 x frac)"}


# ------------------------------------------------------------------------------------------------
# sharded library retrieval (cfg3 / cfg5): forward scoring + fused top-k / peer exchange / merge
# ------------------------------------------------------------------------------------------------
SLAB = 512   # the library is defined in global slabs of 512 shapes (seed = slab id): every world size sees the SAME library


def library_rows(lo, hi, n, dev):
    import torch
    slabs = []
    for sid in range(lo // SLAB, (hi + SLAB - 1) // SLAB if hi > lo else 0):
        x, _ = synth(SLAB, n, 8, seed=7000 + sid)
        a, b = max(lo, sid * SLAB) - sid * SLAB, min(hi, (sid + 1) * SLAB) - sid * SLAB
        slabs.append(x[a:b].to(dev))
    return torch.cat(slabs) if slabs else None


def retrieval_record(ctx, name, steps, warmup, exchange="peer", S=None, Q=None, use_graph=True, pipeline_depth=3):
    """One sharded-retrieval workload: every rank scores its shard of the library, ONE fused kernel exchanges and merges
    the top-k.  Strong scaling.  Rank 0 also holds the whole library and produces, in the same run, the 1-rank time
    and the 1-rank answer; the sharded ids must equal it (asserted) on every rank."""
    import hashlib
    torch, dist, ured, dev = ctx.torch, ctx.dist, ctx.ured, ctx.dev
    S0, Q0, n, desc = RETRIEVAL[name]
    S, Q, k = S or S0, Q or Q0, 10
    heavy = 2.0 * Q * S * n * n > 5e11            # more than ~70 ms per step on one GPU: fewer steps
    st_n, wu_n = (max(2, min(steps, 5)), 2) if heavy else (max(steps, 100), max(warmup, 10))
    lo, hi = ured.shard_bounds(S, ctx.world, ctx.rank)
    shard_x = library_rows(lo, hi, n, dev)
    shard = ured.PackedClouds(shard_x) if shard_x is not None else None
    _, tg_host = synth(Q, 8, n, seed=99)          # same targets on every rank
    tg_pin = tg_host.pin_memory()
    tg_dev = tg_pin.to(dev)
    pairs_per_step = 2.0 * Q * S * n * n
    note = None
    # Query batches stream in: two lanes (streams + graphs + exchange buffers) keep two batches in flight, so the next batch's
    # nn_kernel fills the SMs while the previous batch's tail, epilogue and exchange drain.  Heavy steps (compute-bound for
    # tens of milliseconds) gain nothing from that and run on one lane.
    depth = 1 if (heavy or pipeline_depth < 2) else pipeline_depth
    try:
        engine = ured.RetrievalEngine(shard, lo, Q, k=k, metric="cd_t", use_graph=use_graph, exchange=exchange, pipeline_depth=depth)
        for _ in range(depth):
            v, i = engine.query(tg_dev)
        engine.check()
    except ured.NativeLibraryError as exc:        # peer mapping refused on this box: time the NCCL exchange and say so
        if exchange != "peer" or ctx.world == 1:
            raise
        note = f"peer mapping unavailable ({exc}); NCCL all_gather exchange timed instead"
        exchange = "nccl"
        engine = ured.RetrievalEngine(shard, lo, Q, k=k, metric="cd_t", use_graph=use_graph, exchange="nccl", pipeline_depth=depth)
        for _ in range(depth):
            v, i = engine.query(tg_dev)
    host = [(torch.empty(Q, k).pin_memory(), torch.empty(Q, k, dtype=torch.int32).pin_memory()) for _ in range(depth)]
    state = {"i": 0}

    def step_device():
        engine.submit(tg_dev)

    def step_e2e():
        t = tg_pin.to(dev, non_blocking=True)
        engine.submit(t, host_out=host[state["i"] % depth])
        state["i"] += 1

    def median_of(fn_step, reps):
        # a light query batch is ~0.1 ms: one host hiccup inside a 50-step loop moves the figure by 10 %, so such workloads are
        # timed `reps` times (each loop: st_n steps, one event pair, max over ranks) and the median loop is reported
        runs = [ctx.timed(fn_step, st_n, wu_n, flush=False, whole_loop=True, finish=engine.drain) for _ in range(reps)]
        runs.sort(key=lambda r: r[0])
        return runs[len(runs) // 2]

    reps = 1 if heavy else 3
    ms_step, launches = median_of(step_device, reps)
    ms_e2e, _ = median_of(step_e2e, reps)
    ms_query, _ = ctx.timed(lambda: engine.query(tg_dev), st_n, wu_n, flush=False)      # one batch at a time: the latency of a query
    v, i = engine.query(tg_dev)
    engine.check()
    torch.cuda.synchronize()
    ids = i.cpu().contiguous()
    sha = hashlib.sha1(ids.numpy().tobytes()).hexdigest()
    rec = {"workload": f"{name}: {desc}", "library_shapes": S, "queries": Q, "points": n, "top_k": k, "n_gpus": ctx.world,
           "scaling": "strong", "steps": st_n, "warmup": wu_n, "timed_loops": reps, "ms_per_step": ms_step, "pipeline_depth": depth,
           "ms_per_query_one_at_a_time": ms_query,
           "value": pairs_per_step / (ms_step * 1e-3) / 1e9, "unit": UNIT,
           "e2e": {"ms_per_step": ms_e2e, "value": pairs_per_step / (ms_e2e * 1e-3) / 1e9, "h2d_bytes_per_step": int(tg_pin.numel() * 4),
                   "d2h_bytes_per_step": int(Q * k * 8)},
           "engine": "cuda-graph" if use_graph else "eager",
           "exchange": ("none (1 rank)" if ctx.world == 1 else
                        (f"fused top-k + peer stores + merge in one kernel ({engine.exchange_mapping})" if exchange == "peer"
                         else "top-k kernel + NCCL all_gather + merge kernel")),
           "kernels_per_step": engine.kernels_per_replay if use_graph else launches // max(st_n, 1),
           "ids_sha1": sha, "top1": {"score": float(v[0, 0]), "id": int(i[0, 0])},
           "l2": "library shard (%.0f MB packed + raw per rank) streamed every step; no flush needed above 126 MB, stated for smaller shards" % ((hi - lo) * n * 28 / 1e6)}
    if note:
        rec["note"] = note
    # ---- the 1-rank answer and the 1-rank time, measured in this same run on rank 0 over the WHOLE library --------
    if ctx.world > 1:
        ctx.barrier()
        one = {}
        if ctx.rank == 0:
            full = ured.PackedClouds(library_rows(0, S, n, dev))
            e1 = ured.RetrievalEngine(full, 0, Q, k=k, metric="cd_t", use_graph=use_graph, exchange="peer", pipeline_depth=depth)
            for e in [e1] + e1.lanes:
                e.world, e.exchange = 1, "none"        # a single-rank engine inside a multi-rank job
            st_1, wu_1 = (2, 1) if heavy else (st_n, wu_n)
            ms_1 = sorted(ctx.timed(lambda: e1.submit(tg_dev), st_1, max(wu_1, depth), flush=False, collective=False, whole_loop=True, finish=e1.drain)[0]
                          for _ in range(reps))[reps // 2]
            ms_1q, _ = ctx.timed(lambda: e1.query(tg_dev), st_1, wu_1, flush=False, collective=False)
            v1, i1 = e1.query(tg_dev)
            torch.cuda.synchronize()
            one = {"ms": ms_1, "ms_query": ms_1q, "sha": hashlib.sha1(i1.cpu().contiguous().numpy().tobytes()).hexdigest(), "steps": st_1,
                   "scores_equal": bool(torch.equal(v1.cpu(), v.cpu()))}
            del e1, full
        box = [one]
        dist.broadcast_object_list(box, src=0)
        one = box[0]
        shas = [None] * ctx.world
        dist.all_gather_object(shas, sha)
        rec["ms_per_step_1rank"] = one["ms"]
        rec["ms_per_query_one_at_a_time_1rank"] = one["ms_query"]
        rec["steps_1rank"] = one["steps"]
        rec["ids_sha1_1rank"] = one["sha"]
        rec["ids_identical_on_all_ranks"] = all(s_ == sha for s_ in shas)
        rec["ids_match_1rank"] = one["sha"] == sha
        rec["scores_match_1rank"] = one["scores_equal"]
        if not (rec["ids_identical_on_all_ranks"] and rec["ids_match_1rank"]):
            raise SystemExit(f"bench.py: sharded retrieval {name} at {ctx.world} ranks does not reproduce the 1-rank ranking: {rec}")
    else:
        rec["ms_per_step_1rank"] = ms_step
        rec["ms_per_query_one_at_a_time_1rank"] = ms_query
        rec["ids_sha1_1rank"] = sha
        rec["ids_match_1rank"] = True
    engine.close()
    del engine, shard, shard_x
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------------
# non-product leg: the reference's own CUDA op (recompiled for sm_100a, oracle/_ref) on this GPU
# ------------------------------------------------------------------------------------------------
def reference_cuda_leg(ctx, x_dev, gt_dev, pairs_per_step, ours_ms):
    """time the UNMODIFIED reference op + the reference's torch-op calc_dcd body, fwd+bwd, same inputs."""
    torch = ctx.torch
    try:
        from oracle import ref_cuda
        ref = ref_cuda.load()
        if ref is None:
            return {"unavailable": "oracle/_ref not built on this box"}
        fn = lambda: ref_cuda.calc_dcd_fwd_bwd(ref, x_dev, gt_dev, alpha=ALPHA, n_lambda=N_LAMBDA)  # noqa: E731
        ms, _ = ctx.timed(fn, 10, 3, collective=False)
        ms_fwd, _ = ctx.timed(lambda: ref_cuda.forward(ref, gt_dev, x_dev), 10, 3, collective=False)
        return {"what": "unmodified chamfer3D.cu/chamfer_cuda.cpp built for sm_100a (oracle/_ref) + the reference's torch-op calc_dcd body, fwd+bwd; "
                        "checker code timed as a yardstick, not part of the product path",
                "ms": ms, "gpair_s": pairs_per_step / (ms * 1e-3) / 1e9, "forward_only_ms": ms_fwd,
                "speedup": ms / ours_ms, "steps": 10, "warmup": 3, "l2": "256 MiB buffer rewritten between timed iterations"}
    except Exception as exc:  # the checker is optional
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def emd_rerank_leg(ctx):
    """The re-rank step after the Chamfer top-k (engine/generate_pair.py:95-104): EMD of one target against its 20 best
    candidates, 2048 points, eps 0.005, 50 auction iterations -- our cluster kernel (one launch) next to the reference's own
    op (oracle/_ref/emd: seven launches per iteration), same inputs.  A sub-record, not part of the headline metric."""
    torch, ured, dev = ctx.torch, ctx.ured, ctx.dev
    k, n, eps, iters = 20, 2048, 0.005, 50
    g = torch.Generator().manual_seed(321)
    tgt = torch.rand(1, n, 3, generator=g).to(dev).repeat(k, 1, 1).contiguous()
    cands = torch.rand(k, n, 3, generator=g).to(dev)
    mod = ured.emdModule()
    ms, _ = ctx.timed(lambda: mod(tgt, cands, eps, iters), 10, 3, collective=False)
    rec = {"what": "EMD (auction) of 1 target vs its top-20 candidates, 2048 pts, eps 0.005, 50 iterations", "ms": ms,
           "kernel": "emd_auction_kernel: one launch, a cluster of 8 CTAs per pair, phases separated by cluster barriers"}
    try:
        from oracle import ref_cuda
        ref = ref_cuda.load_emd()
        if ref is not None:
            ms_ref, _ = ctx.timed(lambda: ref_cuda.emd_forward(ref, tgt, cands, eps, iters), 5, 2, collective=False)
            d_ref, a_ref = ref_cuda.emd_forward(ref, tgt, cands, eps, iters)
            d_our, a_our = mod(tgt, cands, eps, iters)
            rec["reference_op"] = {"ms": ms_ref, "speedup": ms_ref / ms, "pairs_with_identical_assignment": int((a_ref == a_our).all(1).sum()),
                                   "pairs": k, "what": "unmodified emd.cpp/emd_cuda.cu built for sm_100a (oracle/_ref/emd), incl. its buffer set-up as in emd_module.py"}
    except Exception as exc:
        rec["reference_op"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    return rec


# ------------------------------------------------------------------------------------------------
# the Chamfer+DCD fwd+bwd workloads (cfg2 headline; cfg1 / cfg4 as sub-records)
# ------------------------------------------------------------------------------------------------
def dcd_workload(ctx, wl, steps, warmup, exact_only, peaks, ffma_peak, sampler=None, with_e2e=True, with_graph=False):
    torch, ured, dev = ctx.torch, ctx.ured, ctx.dev
    B, n_x, n_gt, desc = WORKLOADS[wl]
    pairs_per_step = 2.0 * B * n_x * n_gt
    x_host, gt_host = synth(B, n_x, n_gt, seed=100 + ctx.rank)
    if wl == "cfg2":  # one target per K=10 candidates: both arms see the same broadcast targets
        gt_host = gt_host[::10].repeat_interleave(10, dim=0).contiguous()
    x_pin, gt_pin = x_host.pin_memory(), gt_host.pin_memory()
    x_dev, gt_dev = x_pin.to(dev), gt_pin.to(dev)

    def step_device():
        x = x_dev.detach().requires_grad_()
        gt = gt_dev.detach().requires_grad_()
        loss, _cd_p, _cd_t = ured.calc_dcd(x, gt, alpha=ALPHA, n_lambda=N_LAMBDA)
        loss.sum().backward()
        return loss, x.grad, gt.grad

    ms_step, launches = ctx.timed(step_device, steps, warmup, sampler)
    rec = {"workload": f"{wl}: {desc}", "pairs_per_gpu": B, "n_x": n_x, "n_gt": n_gt, "ms_per_step": ms_step,
           "value": ctx.world * pairs_per_step / (ms_step * 1e-3) / 1e9, "gpu_launches": launches, "pairs_per_step": pairs_per_step}

    if with_graph:
        # the training-step form for repeated shapes: forward + unit-gradient backward captured once (graphed.GraphedDCD)
        g = ured.GraphedDCD(B, n_x, n_gt, alpha=ALPHA, n_lambda=N_LAMBDA, device=dev)

        def step_graph():
            x = x_dev.detach().requires_grad_()
            gt = gt_dev.detach().requires_grad_()
            loss, _cd_p, _cd_t = g(x, gt)
            loss.sum().backward()
            return x.grad

        ms_g, _ = ctx.timed(step_graph, steps, warmup)
        rec["graph"] = {"ms_per_step": ms_g, "value": ctx.world * pairs_per_step / (ms_g * 1e-3) / 1e9,
                        "what": "GraphedDCD as an autograd op: pack -> nn_kernel -> dcd_fwd_kernel -> grad kernel replayed as one CUDA graph, "
                                "gradients scaled by the upstream g_loss (torch's own small launches around it remain)"}
        ms_r, _ = ctx.timed(lambda: g.forward_backward(x_dev, gt_dev), steps, warmup)
        rec["graph_step"] = {"ms_per_step": ms_r, "value": ctx.world * pairs_per_step / (ms_r * 1e-3) / 1e9,
                             "what": "GraphedDCD.forward_backward: the whole step (two input copies + one replay of pack, nn_kernel, dcd_fwd_kernel, "
                                     "grad kernel) outside autograd; returns loss and d sum(loss)/d clouds"}

    if with_e2e:
        # ---- end to end: host buffers in, host result out, every step ---------------------------------
        # Q targets [Q,N,3] and their K candidates [Q*K,M,3] sit in pinned host memory; each step copies BOTH to the device
        # (copy stream, double buffered so that step i+1's upload overlaps step i's kernels, as a pinned-memory data loader
        # does), broadcasts each target over its K candidates on the device, runs calc_dcd fwd+bwd and copies the per-pair
        # loss back to pinned host memory.
        K_CAND = 10 if (wl == "cfg2" and B % 10 == 0) else 1
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

        ms_e2e, _ = ctx.timed(step_e2e, steps, warmup, whole_loop=True)
        torch.cuda.synchronize()
        rec["e2e"] = {"value": ctx.world * pairs_per_step / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e,
                      "h2d_bytes_per_step": int(x_pin.numel() * 4 + gt_small_pin.numel() * 4), "d2h_bytes_per_step": int(B * 4),
                      "api": "pinned-host candidates + targets copied every step (copy stream, double-buffered), targets broadcast on device, "
                             "calc_dcd fwd+bwd, per-pair loss copied back to pinned host"}
    rec["roofline"] = nn_roofline(ctx, x_dev, gt_dev, B, n_x, n_gt, steps, warmup, exact_only, peaks, ffma_peak, wl)
    rec["_inputs"] = (x_dev, gt_dev)
    return rec


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
    ap.add_argument("--library-size", type=int, default=0, help="override S for a stand-alone retrieval workload")
    ap.add_argument("--queries", type=int, default=0, help="override Q for a stand-alone retrieval workload")
    ap.add_argument("--eager", action="store_true", help="retrieval: eager calls instead of the CUDA-graph engine")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="retrieval on >1 rank: fused peer-memory exchange or NCCL all_gather")
    ap.add_argument("--pipeline-depth", type=int, default=3, help="retrieval: query batches kept in flight (lanes of the engine); 1 = one at a time")
    ap.add_argument("--exact-only", action="store_true", help="disable the screening pass (difference form on every pair)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload: no cfg1/cfg4 sub-records, reference-op leg or retrieval record")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args, args.workload if args.workload in WORKLOADS else "cfg2")
        return

    ctx = Ctx()
    torch, dist = ctx.torch, ctx.dist
    if args.exact_only:
        os.environ["URED_EXACT_ONLY"] = "1"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    if args.workload in RETRIEVAL:   # stand-alone retrieval run (development): its record is the line
        rec = retrieval_record(ctx, args.workload, args.steps, args.warmup, exchange=args.exchange, S=args.library_size or None,
                               Q=args.queries or None, use_graph=not args.eager, pipeline_depth=args.pipeline_depth)
        if ctx.rank == 0:
            print(json.dumps(rec), flush=True)
        if ctx.world > 1:
            dist.destroy_process_group()
        return

    ffma_peak = measure_ffma_peak(ctx)
    sampler = ClockSampler(ctx.local_rank)
    wl = args.workload
    main_rec = dcd_workload(ctx, wl, args.steps, args.warmup, args.exact_only, peaks, ffma_peak, sampler=sampler)
    x_dev, gt_dev = main_rec.pop("_inputs")
    B, n_x, n_gt, desc = WORKLOADS[wl]
    line = {
        "metric": METRIC, "value": main_rec["value"], "unit": UNIT,
        "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_rec["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": main_rec["workload"], "pairs_per_gpu": B, "n_x": n_x, "n_gt": n_gt, "alpha": ALPHA,
                   "n_lambda": N_LAMBDA, "step": "calc_dcd fwd + loss.sum().backward()", "sharding": "independent pairs per rank, no collective "
                   "(the sharded workload with its exchange is the `retrieval` record)",
                   "kernel_variant": "exact-only" if args.exact_only else "screen+exact-recheck",
                   "l2": "256 MiB buffer rewritten between timed iterations"},
        "e2e": main_rec["e2e"], "gpu_launches": main_rec["gpu_launches"], "roofline": main_rec["roofline"], "clocks": sampler.summary(),
    }
    if not args.no_extras:
        if ctx.world == 1:
            # the other single-GPU configurations of BASELINE.json, each with its own roofline
            subs = {}
            for name in ("cfg1", "cfg4"):
                if name == wl:
                    continue
                r = dcd_workload(ctx, name, max(10, min(args.steps, 30)), args.warmup, args.exact_only, peaks, ffma_peak,
                                 with_e2e=False, with_graph=(name == "cfg1"))
                r.pop("_inputs")
                subs[name] = r
                torch.cuda.empty_cache()
            line["configs"] = subs
            line["reference_cuda_op"] = reference_cuda_leg(ctx, x_dev, gt_dev, main_rec["pairs_per_step"], main_rec["ms_per_step"])
            line["emd_rerank"] = emd_rerank_leg(ctx)
        del x_dev, gt_dev
        torch.cuda.empty_cache()
        retr = {}
        for name in ("cfg3", "cfg5"):
            retr[name] = retrieval_record(ctx, name, args.steps, args.warmup, exchange=args.exchange, pipeline_depth=args.pipeline_depth)
        line["retrieval"] = retr
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        sample_pairs = 32 if n_x * n_gt <= 2048 * 2048 else 1
        line["cpu_baseline"] = {k: v for k, v in cpu_baseline(n_x, n_gt, sample_pairs, 3, 1).items() if k != "sec_per_step"}
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
