#!/usr/bin/env python
"""bench.py -- samples/sec folded into ln Z (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--workload cfg4|cfg2|cfg2x|cfg3|cfg5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic chains: evaluate the flow on
every sample, fold ln phi - ln posterior into the per-chain state, (N > 1: one all-gather of the
per-chain partials), finalise ln(1/Z) and its error estimates.  Default workload: cfg4, the
configuration the north_star target is quoted on (RealNVP 2+4 coupling layers, ndim 64, 1e8 samples
as 1000 chains x 1e5).

Scaling.  BASELINE.json's configs shard a FIXED problem over 1/2/4/8 GPUs, so `value` is STRONG scaling:
the 1000 chains (1e8 samples) are split into contiguous chain ranges, one per rank.  `value_weak` (N > 1)
is the same step with every rank owning its own 1000 chains (1e8 samples per GPU).

`value`            device-resident inputs, CUDA-event timed, max over ranks.
`e2e`              the public API (Evidence.add_chains) on pinned HOST float32 chains: H2D inside the timed region.
`e2e_pageable_f64` the same call on numpy float64 Chains built as harmonic/chains.py builds them (pageable).
`roofline`         the dominant kernel timed live with CUDA events inside the library.
`secondary`        short runs of the other BASELINE configs (cfg5, cfg3, cfg2), same sharding rule.
`mgpu_parity`      (N > 1) sharded add_chains against the single-GPU result on the same small chains.
`cpu_baseline`     the float32 CPU restatement (oracle/) on a bounded sample, all host cores.
--impl reference: the reference's CPU algorithm (the oracle port: the real JAX path is not installable
in this image) timed the same way on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "samples/sec folded into ln Z"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


# ------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_pass(name, model, nchains, per_chain, seed, threads):
    """The reference's CPU algorithm (float32 restatement in oracle/) over one bounded sample:
    predict on every sample (all host threads, row blocks) + add_chains + process_run."""
    from concurrent.futures import ThreadPoolExecutor

    import bench_workloads as wl
    from oracle import evidence as oe
    from oracle import flows as of

    n = nchains * per_chain
    start = np.arange(nchains + 1, dtype=np.int64) * per_chain
    if name == "cfg5":
        rng = np.random.default_rng(seed)
        lnpred = (rng.standard_normal(n) - 20.0).astype(np.float32)
        lnp = (lnpred + 10.0 + 0.05 * rng.standard_normal(n)).astype(np.float32)
        t0 = time.perf_counter()
    else:
        x, lnp = wl.host_sample(name, n, seed)
        spec = model.oracle_spec()
        blocks = [(a, min(n, a + 8192)) for a in range(0, n, 8192)]
        lnpred = np.empty(n, dtype=np.float32)

        def work(ab):
            a, b = ab
            lnpred[a:b] = of.predict(spec, x[a:b], np.float32)

        try:
            from threadpoolctl import threadpool_limits
            ctx = threadpool_limits(limits=1)
        except Exception:
            ctx = None
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, blocks))
        if ctx is not None:
            ctx.restore_original_limits() if hasattr(ctx, "restore_original_limits") else None
    ev = oe.EvidenceOracle(nchains, oe.MEAN_SHIFT, f32=True)
    ev.add(lnpred, lnp, start)
    dt = time.perf_counter() - t0
    return n / dt, dt, ev


def cpu_sample_shape(name):
    # about 10-30 s of CPU work on a 16-core host
    return {"cfg4": (160, 25000), "cfg3": (400, 20000), "cfg1": (100, 5000), "cfg2": (20, 10000), "cfg2x": (40, 25000),
            "cfg5": (1000, 20000)}[name]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import bench_workloads as wl

    name = args.workload
    threads = os.cpu_count() or 1
    model = None if name == "cfg5" else wl.make_model(name)
    nchains, per_chain = cpu_sample_shape(name)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, ev = cpu_reference_pass(name, model, nchains, per_chain, seed=i, threads=threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = "%d chains x %d samples per step of the %s workload (float32 CPU restatement of " \
             "harmonic/flows.py + evidence.py; the JAX reference is not installable in this image)" % (
                 nchains, per_chain, name)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": scaling_mode(args), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, args.gpus, scaling_mode(args), args.scale),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


WORKLOAD_DESC = {
    "cfg4": "cfg4: RealNVP 2 scaled + 4 unscaled coupling layers (hidden 128), ndim 64, T=0.8, "
            "Gaussian target, predict fused with harmonic-mean accumulation",
    "cfg3": "cfg3: RealNVP 6+2, ndim 5, T=0.9",
    "cfg1": "cfg1: RealNVP 2+4, ndim 5, T=0.8",
    "cfg2": "cfg2: RQSpline 8 layers x 8 bins, hidden [64,64], ndim 2 (Rosenbrock), T=0.9",
    "cfg2x": "cfg2 with the example's architecture: RQSpline 3 layers x 8 bins, hidden [32,32], ndim 2, T=0.9",
    "cfg5": "cfg5: accumulation only, precomputed ln phi and ln posterior (float32)",
}
BYTES = {"cfg4": 260, "cfg3": 24, "cfg1": 24, "cfg2": 12, "cfg2x": 12, "cfg5": 8}


def scaling_mode(args):
    return args.scaling or "strong"


def shard_shape(name, world, scaling, scale=1.0):
    """(total chains, samples per chain, chain bounds per rank).  strong: BASELINE's fixed problem split into
    contiguous chain ranges; weak: every rank owns a full copy of the problem's shape."""
    import bench_workloads as wl

    nchains, per_chain = wl.default_shape(name)
    per_chain = max(1, int(per_chain * scale))
    if scaling == "weak":
        total = nchains * world
        bounds = [r * nchains for r in range(world + 1)]
    else:
        total = nchains
        bounds = [int(v) for v in np.linspace(0, nchains, world + 1).astype(np.int64)]
    return total, per_chain, bounds


def workload_config(name, n_gpus, scaling, scale=1.0):
    total, per_chain, bounds = shard_shape(name, n_gpus, scaling, scale)
    per_gpu = max(bounds[r + 1] - bounds[r] for r in range(n_gpus)) * per_chain
    return {"workload": WORKLOAD_DESC[name], "nchains_total": total, "samples_per_chain": per_chain,
            "samples_total": total * per_chain, "samples_per_gpu": per_gpu, "scaling": scaling,
            "sharding": "contiguous chain ranges over %d GPU(s), one all-gather of per-chain partials" % n_gpus,
            "l2": "inputs (%.2f GB per GPU) %s the 126 MB L2%s" % (
                per_gpu * BYTES[name] / 1e9, "are far larger than" if per_gpu * BYTES[name] > 1e9 else "vs",
                "; no flush needed" if per_gpu * BYTES[name] > 1e9 else
                "; 192 MB of zeros are written between steps to flush it")}


# ------------------------------------------------------------------------------------- GPU arm
_JSON_OUT = None


def emit(text):
    """The one JSON line goes to the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def claim_stdout():
    """Libraries write banners to file descriptor 1 (NCCL prints its version there on communicator creation):
    keep a private handle on the real stdout for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


class Ctx:
    """Per-process state shared by the timed passes."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world > 1:
            # staging threads of the host-input path: share the host's cores between the ranks
            os.environ.setdefault("HB_HOST_THREADS", str(max(1, min(16, (os.cpu_count() or 1) // self.world))))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        from harmonic_b200 import _lib
        self.lib = _lib.require_gpu()
        self._lib = _lib
        self.numa_cpus = self.lib.hb_bind_host_to_device_node(self.local) if self.world > 1 else 0
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.comm = True if self.world > 1 else None
        self.args = args
        self._flush = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def flush_l2(self):
        if self._flush is None:
            self._flush = self.torch.empty(192 << 20, dtype=self.torch.uint8, device=self.dev)
        self._flush.zero_()


class Precomputed:  # cfg5: a fitted model is only needed for ndim / is_fitted
    ndim = 1
    flow = object()
    device = 0

    def is_fitted(self):
        return True


def make_model(ctx, name):
    import bench_workloads as wl

    if name == "cfg5":
        m = Precomputed()
        m.device = ctx.local
        return m
    m = wl.make_model(name)
    m.device = ctx.local
    m.kernel_path = {"auto": ctx._lib.HB_PATH_AUTO, "simt": ctx._lib.HB_PATH_SIMT,
                     "tcgen05": ctx._lib.HB_PATH_TCGEN05}[ctx.args.path]
    return m


def timed_device_run(ctx, name, scaling, steps, warmup, scale=1.0, sample_clocks=False):
    """Device-resident run of one workload: returns value (whole job, samples/s), ms per step (max over
    ranks), the dominant kernel's mean duration on this rank, launches, the last Evidence."""
    import bench_workloads as wl
    import harmonic_b200 as hm

    torch = ctx.torch
    total_chains, per_chain, bounds = shard_shape(name, ctx.world, scaling, scale)
    lo, hi = bounds[ctx.rank], bounds[ctx.rank + 1]
    nloc = hi - lo
    n_local = nloc * per_chain
    model = make_model(ctx, name)
    a, b = wl.device_data(name, nloc, per_chain, ctx.dev, seed=ctx.rank)
    start = np.arange(nloc + 1, dtype=np.int64) * per_chain
    chains = None if name == "cfg5" else hm.Chains.from_arrays(a, b, start)
    small = n_local * BYTES[name] < 1e9  # inputs comparable to the L2: flush between steps
    torch.cuda.synchronize()
    holder = []

    def one_pass():
        ev = hm.Evidence(total_chains, model)
        ev.device = ctx.local
        if name == "cfg5":
            ev.add_lnpred(a, b, start, chain_offset=lo, comm=ctx.comm)
        else:
            ev.add_chains(chains, chain_offset=lo, comm=ctx.comm)
        holder[:] = [ev]

    for _ in range(warmup):
        one_pass()
    ctx.lib.hb_set_kernel_timing(1)
    sampler = ClockSampler(ctx.local) if (sample_clocks and ctx.rank == 0) else None
    if sampler:
        sampler.start()
    ctx.barrier()
    launches0 = ctx.lib.hb_launch_count()
    kernel_ms = []
    if small:
        # per-step events so that the L2 flush between steps stays outside the timed intervals
        ms = 0.0
        for _ in range(steps):
            ctx.flush_l2()
            ctx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one_pass()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
            kernel_ms.append(ctx.lib.hb_last_kernel_ms(ctx.local))
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one_pass()
            kernel_ms.append(ctx.lib.hb_last_kernel_ms(ctx.local))
        e1.record()
        ctx.barrier()
        ms = e0.elapsed_time(e1)
    launches = ctx.lib.hb_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ctx.lib.hb_set_kernel_timing(0)
    ms = ctx.max_over_ranks(ms)
    total = total_chains * per_chain
    kms = [k for k in kernel_ms if k > 0]
    out = {"value": total * steps / (ms * 1e-3), "ms_per_step": ms / steps,
           "kernel_ms": float(np.mean(kms)) if kms else None, "launches": int(launches), "clocks": clocks,
           "n_local": n_local, "total": total, "ev": holder[0], "model": model, "data": (a, b, start, lo),
           "total_chains": total_chains, "per_chain": per_chain, "l2_flushed": small}
    return out


def roofline_block(name, path, run, peaks, peak_kind):
    import bench_workloads as wl

    kms, n_local = run["kernel_ms"], run["n_local"]
    if name == "cfg5":
        ach = wl.BYTES_PER_SAMPLE[name] * n_local / (kms * 1e-3) / 1e9 if kms else None
        peak = float(peaks["hbm_gbs"])
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None,
                # ncu --set full (profiles/r2_prof_acc_r2e_raw_summary.csv): dram read 8.0002 + write 0.0510 B per sample
                "traffic": 8.0512 * n_local, "traffic_unit": "bytes per launch (ncu dram bytes/sample x samples)",
                "kernel": "accumulate_precomputed_kernel", "kernel_ms": kms,
                "peak_source": "HBM copy bandwidth, of %s" % peak_kind}
    if name == "cfg4" and path != "simt":
        ach = wl.FLOP_PER_SAMPLE[name] * n_local / (kms * 1e-3) / 1e12 if kms else None
        peak = float(peaks["bf16_tflops"]) / 2.0
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": (ach / peak) if ach else None,
                # ncu --set full (profiles/r2_prof_tc_r2e_raw_summary.csv): 262.1 B of DRAM traffic per sample (read
                # 260.2 + write 1.8; algorithmic 260 B): the kernel is not memory bound
                "traffic": 262.1 * n_local, "traffic_unit": "DRAM bytes per launch (ncu bytes/sample x samples)",
                "kernel": "realnvp_tc_kernel", "kernel_ms": kms,
                "peak_source": "TF32 dense = bf16 burst / 2, of %s (the roofline BASELINE.json names; the kernel issues "
                               "kind::f16 MMAs, against the measured 16-bit peak the fraction is half of frac)" % peak_kind,
                "frac_of_16bit_peak": (ach / (2.0 * peak)) if ach else None,
                "algorithmic_flop_per_sample": wl.FLOP_PER_SAMPLE[name]}
    if name == "cfg3" and path != "simt":
        # RealNVP at ndim 5 with >= 2^20 rows per call: realnvp_small_tc_kernel, both contractions of the coupling MLP on
        # the tensor cores (three FP16 MMAs per float32-equivalent product).  Tiles are 90 % padding at this shape (K = 16
        # for d + 1 = 3 inputs, N = 16 for 2u = 6 outputs), so the tensor-pipe fraction says little; the figure beside
        # it is the algorithmic FLOP rate against what the FP32 CUDA cores could deliver at most (> 1 means the kernel
        # is beyond any CUDA-core implementation's reach).
        ach = wl.FLOP_PER_SAMPLE[name] * n_local / (kms * 1e-3) / 1e12 if kms else None
        peak = float(peaks["bf16_tflops"]) / 2.0
        fp32_peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": (ach / peak) if ach else None,
                "frac_of_fp32_cuda_core_peak": (ach / fp32_peak) if ach else None,
                "traffic": None, "kernel": "realnvp_small_tc_kernel", "kernel_ms": kms,
                "peak_source": "TF32 dense = bf16 burst / 2, of %s; FP32 CUDA-core peak nominal" % peak_kind,
                "algorithmic_flop_per_sample": wl.FLOP_PER_SAMPLE[name]}
    if name in ("cfg2", "cfg2x") and path != "simt":
        # RQSpline at ndim 2: rqspline_tc_kernel.  The h1 x 25 contraction (93 % of the executed FLOPs) runs on the
        # tensor cores as three FP16 MMAs per float32-equivalent product; what bounds the kernel is the CUDA cores'
        # work around it -- two tanh layers (MUFU ex2 / rcp) and the spline -- so the fraction that says how close the
        # kernel is to the hardware is the MUFU pipe's (16 results per clock and SM), beside the tensor-pipe figures.
        # SURVEY.md 8(d) counts the reference's dense conditioner (all D rows of the (D, 3K+1) output, two separate
        # last Denses); the kernel pre-multiplies the last two Denses (no activation between them,
        # harmonic/flows.py:357,434-438) and evaluates only the transformed row, so it EXECUTES 4.3x fewer FLOPs.
        L, h = (8, 64) if name == "cfg2" else (3, 32)
        ex = L * 2.0 * (2 * 2 + 2 * h + h * 25)
        mufu = L * (1.5 * (h + 2) + 25.0)  # per tanh pair 2 ex2 + 1 rcp; spline: 16 ex2, 6 rcp, 2 ex2 (softplus), 1 lg2
        ach = wl.FLOP_PER_SAMPLE[name] * n_local / (kms * 1e-3) / 1e12 if kms else None
        peak = float(peaks["bf16_tflops"]) / 2.0
        mufu_peak = 148 * 16 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
        return {"bound": "mufu", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": (mufu * n_local / (kms * 1e-3) / mufu_peak) if kms else None,
                "frac_is": "MUFU results per second / (148 SMs x 16 per clock x max SM clock): the pipe that bounds it",
                "frac_algorithmic_of_tf32": (ach / peak) if ach else None,
                "frac_executed_of_tf32": (ach / peak * ex / wl.FLOP_PER_SAMPLE[name]) if ach else None,
                "executed_flop_per_sample": ex, "mufu_per_sample": mufu,
                "traffic": None, "kernel": "rqspline_tc_kernel", "kernel_ms": kms,
                "peak_source": "TF32 dense = bf16 burst / 2, of %s; MUFU peak nominal" % peak_kind,
                "algorithmic_flop_per_sample": wl.FLOP_PER_SAMPLE[name]}
    # small-ndim flows run on the FP32 CUDA cores (tensor tiles would be >= 75 % padding):
    # denominator = 148 SMs x 128 FMA lanes x 2 FLOP x max SM clock (nominal, not measured)
    ach = wl.FLOP_PER_SAMPLE[name] * n_local / (kms * 1e-3) / 1e12 if kms else None
    peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
    roof = {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": (ach / peak) if ach else None, "traffic": None,
            "kernel": {"cfg2": "rqspline_fast_kernel", "cfg2x": "rqspline_fast_kernel", "cfg3": "realnvp_small_kernel",
                       "cfg1": "realnvp_small_kernel"}.get(name, "simt flow kernel"), "kernel_ms": kms,
            "peak_source": "FP32 CUDA-core FMA peak at the max SM clock (nominal)",
            "algorithmic_flop_per_sample": wl.FLOP_PER_SAMPLE[name]}
    if name in ("cfg2", "cfg2x") and ach:
        # the CUDA-core register kernel (--path simt): executed FLOPs against the FP32 FMA peak (the spline's MUFU
        # work is not counted)
        L, h = (8, 64) if name == "cfg2" else (3, 32)
        ex = L * 2.0 * (2 * 2 + 2 * h + h * 25)
        roof["executed_flop_per_sample"] = ex
        roof["frac_executed"] = roof["frac"] * ex / wl.FLOP_PER_SAMPLE[name]
        roof["frac_algorithmic"] = roof["frac"]
        roof["frac"] = roof["frac_executed"]
    return roof


def e2e_runs(ctx, name, run, steps):
    """The public API on HOST chains, H2D inside the timed region: (a) pinned float32, the whole per-rank shard;
    (b) pageable float64 Chains built the way harmonic/chains.py builds them (bounded sample)."""
    import harmonic_b200 as hm

    torch = ctx.torch
    a, b, start, lo = run["data"]
    per_chain, total_chains, model = run["per_chain"], run["total_chains"], run["model"]
    nloc = len(start) - 1
    nst = max(5, min(steps, 8))
    out = {}

    def timed(make_pass, rows, h2d, key, extra):
        make_pass()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(nst):
            res = make_pass()
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        tot_rows = ctx.sum_over_ranks(rows)
        out[key] = dict({"value": tot_rows * nst / dt, "unit": "samples/s", "steps": nst,
                         "h2d_bytes_per_step": int(ctx.sum_over_ranks(h2d)),
                         "d2h_bytes_per_step": 128 * ctx.world,
                         "samples_per_step": int(tot_rows), "ln_evidence": float(res[0])}, **extra)

    # (a) pinned float32
    ha = torch.empty(a.shape, dtype=a.dtype).pin_memory()
    hb_ = torch.empty(b.shape, dtype=b.dtype).pin_memory()
    ha.copy_(a)
    hb_.copy_(b)
    xa, xb = ha.numpy(), hb_.numpy()
    ch_pinned = None if name == "cfg5" else hm.Chains.from_arrays(xa, xb, start)

    def pass_pinned():
        evh = hm.Evidence(total_chains, model)
        evh.device = ctx.local
        if name == "cfg5":
            evh.add_lnpred(xa, xb, start, chain_offset=lo, comm=ctx.comm)
        else:
            evh.add_chains(ch_pinned, chain_offset=lo, comm=ctx.comm)
        return evh.compute_ln_evidence()  # reads the step's result (128 bytes D2H inside add_chains)

    timed(pass_pinned, a.shape[0], xa.nbytes + xb.nbytes, "e2e",
          {"input": "pinned host float32, %d samples per GPU" % a.shape[0]})
    del ha, hb_, xa, xb, ch_pinned

    # (b) pageable float64 Chains, as harmonic/chains.py:65-69 builds them (np.concatenate of float64 arrays)
    if name != "cfg5":
        ch_n = max(1, min(nloc, int(2.0e7 // max(1, per_chain) // ctx.world) or 1))
        rows = ch_n * per_chain
        D = a.shape[1]
        ch64 = hm.Chains(D)
        ch64.add_chains_3d(a[:rows].cpu().numpy().astype(np.float64).reshape(ch_n, per_chain, D),
                           b[:rows].cpu().numpy().astype(np.float64).reshape(ch_n, per_chain))
        assert ch64.samples.dtype == np.float64 and ch64.ln_posterior.dtype == np.float64
        tot_e = int(ctx.sum_over_ranks(ch_n))
        off_e = ch_n * ctx.rank

        def pass_pageable():
            evh = hm.Evidence(tot_e, model)
            evh.device = ctx.local
            evh.add_chains(ch64, chain_offset=off_e, comm=ctx.comm)
            return evh.compute_ln_evidence()

        timed(pass_pageable, rows, ch64.samples.nbytes // 2 + ch64.ln_posterior.nbytes // 2, "e2e_pageable_f64",
              {"input": "numpy float64 Chains (pageable), %d samples per GPU; host threads convert to float32 "
                        "into pinned staging, so the H2D bytes are half the source bytes" % rows,
               "source_bytes_per_step": int(ctx.sum_over_ranks(ch64.samples.nbytes + ch64.ln_posterior.nbytes))})
        del ch64
    # host-side ceiling of the staged path: aggregate STREAM-style copy rate with the staging threads
    ctx.barrier()
    th = int(os.environ.get("HB_HOST_THREADS", "0")) or min(16, os.cpu_count() or 1)
    gbs = ctx.lib.hb_host_copy_gbs(256 << 20, th)
    out["host_copy_gbs"] = {"value": ctx.sum_over_ranks(gbs), "threads_per_rank": th, "numa_cpus_bound": ctx.numa_cpus,
                            "note": "sum over ranks of a concurrent 256 MB copy loop (read + write bytes)"}
    return out


def mgpu_parity(ctx):
    """N > 1: chain-sharded add_chains (all-gather of partials) against the unsharded call on the SAME small
    chains on every rank; returns the largest absolute difference over the finalised statistics."""
    import bench_workloads as wl
    import harmonic_b200 as hm

    worst = worst_rel = 0.0
    for name, nch, per in (("cfg4", 8 * ctx.world + 3, 1500), ("cfg2x", 4 * ctx.world + 1, 700)):
        model = make_model(ctx, name)
        x, lnp = wl.host_sample(name, nch * per, seed=11)
        start = np.arange(nch + 1, dtype=np.int64) * per
        bounds = np.linspace(0, nch, ctx.world + 1).astype(np.int64)
        lo, hi = int(bounds[ctx.rank]), int(bounds[ctx.rank + 1])
        full = hm.Chains.from_arrays(x, lnp, start)
        mine = hm.Chains.from_arrays(x[start[lo]:start[hi]], lnp[start[lo]:start[hi]], start[lo:hi + 1] - start[lo])
        ev_s = hm.Evidence(nch, model)
        ev_s.device = ctx.local
        ev_1 = hm.Evidence(nch, model)
        ev_1.device = ctx.local
        for _ in range(2):  # two streaming calls
            ev_s.add_chains(mine, chain_offset=lo, comm=ctx.comm)
            ev_1.add_chains(full)
        for k in ("ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff",
                  "lnargmax", "lnargmin"):
            a, b = float(getattr(ev_s, k)), float(getattr(ev_1, k))
            worst = max(worst, abs(a - b))
            worst_rel = max(worst_rel, abs(a - b) / max(abs(b), 1.0))
        d = np.abs(ev_s.ln_evidence_inv_per_chain - ev_1.ln_evidence_inv_per_chain)
        worst = max(worst, float(np.max(d)))
        worst_rel = max(worst_rel, float(np.max(d / np.maximum(np.abs(ev_1.ln_evidence_inv_per_chain), 1.0))))
    # (not exactly zero in general: the per-thread float32 partial sums of the fold see other tile boundaries when
    # the rows of a rank start elsewhere; the per-sample ln phi are bit-identical)
    return {"max_abs_diff": ctx.max_over_ranks(worst), "max_rel_diff": ctx.max_over_ranks(worst_rel), "cases": "cfg4 (tcgen05) and cfg2x (RQSpline) flows, ragged "
            "chain split, two streaming calls; ln_evidence_inv, its variance / variance-of-variance, kurtosis, n_eff, "
            "extrema and per-chain ln_evidence_inv"}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg3", "cfg2", "cfg2x", "cfg5", "cfg1"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale samples per chain (debug)")
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="strong (default): BASELINE's fixed problem sharded over the GPUs; weak: per-GPU copy")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)  # honour the caller; the contract default is >= 3
    if args.impl == "reference":
        return run_reference(args)

    ctx = Ctx(args)
    name = args.workload
    scaling = scaling_mode(args)
    peaks, peak_kind = load_peaks()

    run = timed_device_run(ctx, name, scaling, args.steps, args.warmup, args.scale, sample_clocks=True)
    ev = run["ev"]
    result = {"ln_evidence_inv": float(ev.ln_evidence_inv), "ln_evidence_inv_var": float(ev.ln_evidence_inv_var),
              "n_eff": float(ev.n_eff), "guard_trips": int(ev.guard_trips)}

    e2e = None
    if not args.no_e2e:
        e2e = e2e_runs(ctx, name, run, args.steps)
    roof = roofline_block(name, args.path, run, peaks, peak_kind)
    main_keep = {k: run[k] for k in ("value", "ms_per_step", "kernel_ms", "launches", "clocks", "l2_flushed")}
    del run, ev
    ctx.torch.cuda.empty_cache()

    weak = None
    if ctx.world > 1 and scaling == "strong":
        w = timed_device_run(ctx, name, "weak", max(3, min(args.steps, 8)), 3, args.scale)
        weak = {"value": w["value"], "ms_per_step": w["ms_per_step"], "kernel_ms": w["kernel_ms"],
                "samples_per_gpu": w["n_local"]}
        del w
        ctx.torch.cuda.empty_cache()

    secondary = None
    if not args.no_secondary and name == "cfg4":
        secondary = {}
        for sname in ("cfg5", "cfg3", "cfg2"):
            s = timed_device_run(ctx, sname, scaling, max(3, min(args.steps, 8)), 3, args.scale)
            r = roofline_block(sname, "auto", s, peaks, peak_kind)
            secondary[sname] = {
                "value": s["value"], "unit": "samples/s", "ms_per_step": s["ms_per_step"], "kernel": r["kernel"],
                "kernel_ms": s["kernel_ms"], "samples_total": s["total"], "samples_per_gpu": s["n_local"],
                "roofline_bound": r["bound"], "roofline_frac": r["frac"], "l2_flushed_between_steps": s["l2_flushed"],
                "ln_evidence_inv": float(s["ev"].ln_evidence_inv)}
            if sname == "cfg5":
                # step-level: whole-step samples/s of ONE GPU's share against the HBM bound 8 B/sample
                secondary[sname]["step_frac_of_hbm_bound"] = (
                    s["n_local"] * 8.0 / (s["ms_per_step"] * 1e-3) / 1e9 / float(peaks["hbm_gbs"]))
            del s
            ctx.torch.cuda.empty_cache()

    parity = mgpu_parity(ctx) if ctx.world > 1 else None

    if ctx.rank != 0:
        if ctx.world > 1:
            ctx.dist.destroy_process_group()
        return 0

    cpu = None
    if not args.no_cpu and ctx.world == 1:
        threads = os.cpu_count() or 1
        cn, cp = cpu_sample_shape(name)
        model = None if name == "cfg5" else make_model(ctx, name)
        v, dt, _ = cpu_reference_pass(name, model, cn, cp, seed=0, threads=threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "%d chains x %d samples (%.1f s) of the same workload, float32 restatement in oracle/"
                         % (cn, cp, dt)}

    tc = name in ("cfg4", "cfg3", "cfg2", "cfg2x") and args.path != "simt"
    line = {
        "metric": METRIC, "value": main_keep["value"], "unit": "samples/s", "n_gpus": ctx.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main_keep["ms_per_step"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None,
        "dtype": "f16 3-product split on tcgen05 (f32-equivalent products, f32 accumulate)" if tc else "f32",
        "data": "synthetic", "config": workload_config(name, ctx.world, scaling, args.scale),
        "clocks": main_keep["clocks"], "e2e": e2e["e2e"] if e2e else None,
        "e2e_pageable_f64": e2e.get("e2e_pageable_f64") if e2e else None,
        "host_copy_gbs": e2e.get("host_copy_gbs") if e2e else None,
        "gpu_launches": main_keep["launches"], "roofline": roof, "cpu_baseline": cpu,
        "value_weak": weak, "secondary": secondary, "mgpu_parity": parity, "result": result,
    }
    emit(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
