#!/usr/bin/env python
"""bench.py -- samples/sec folded into ln Z (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--workload cfg4|cfg2|cfg2x|cfg3|cfg5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic chains: evaluate the flow on
every sample, fold ln phi - ln posterior into the per-chain state, (N > 1: one all-gather of the
per-chain partials), finalise ln(1/Z) and its error estimates.  Default workload: cfg4, the
configuration the north_star target is quoted on (RealNVP 2+4 coupling layers, ndim 64, 1e8 samples
per GPU as 1000 chains x 1e5).  Weak scaling: every rank owns its own 1000 chains.

`value`        device-resident inputs, CUDA-event timed, max over ranks.
`e2e`          the public API (Evidence.add_chains) on pinned HOST chains: H2D inside the timed region.
`roofline`     the dominant kernel timed live with CUDA events inside the library.
`cpu_baseline` the float32 CPU restatement (oracle/) on a bounded sample, all host cores.
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
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, args.gpus, *wl.default_shape(name)),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


def workload_config(name, n_gpus, nchains, per_chain):
    desc = {
        "cfg4": "cfg4: RealNVP 2 scaled + 4 unscaled coupling layers (hidden 128), ndim 64, T=0.8, "
                "Gaussian target, predict fused with harmonic-mean accumulation",
        "cfg3": "cfg3: RealNVP 6+2, ndim 5, T=0.9",
        "cfg1": "cfg1: RealNVP 2+4, ndim 5, T=0.8",
        "cfg2": "cfg2: RQSpline 8 layers x 8 bins, hidden [64,64], ndim 2 (Rosenbrock), T=0.9",
        "cfg2x": "cfg2 with the example's architecture: RQSpline 3 layers x 8 bins, hidden [32,32], ndim 2, T=0.9",
        "cfg5": "cfg5: accumulation only, precomputed ln phi and ln posterior (float32)",
    }[name]
    return {"workload": desc, "nchains_per_gpu": nchains, "samples_per_chain": per_chain,
            "samples_per_gpu": nchains * per_chain, "sharding": "chains sharded across %d GPU(s), one "
            "all-gather of per-chain partials" % n_gpus,
            "l2": "inputs (%.1f GB per GPU) are far larger than the 126 MB L2; no flush needed" % (
                nchains * per_chain * {"cfg4": 260, "cfg3": 24, "cfg1": 24, "cfg2": 12, "cfg2x": 12, "cfg5": 8}[name] / 1e9)}


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


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg3", "cfg2", "cfg2x", "cfg5", "cfg1"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale samples per chain (debug)")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)  # honour the caller; the contract default is >= 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import bench_workloads as wl
    import harmonic_b200 as hm
    from harmonic_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.require_gpu()
    name = args.workload
    nchains, per_chain = wl.default_shape(name)
    per_chain = max(1, int(per_chain * args.scale))
    n_local = nchains * per_chain
    total_chains = nchains * world
    comm = True if world > 1 else None

    class Precomputed:  # cfg5: a fitted model is only needed for ndim / is_fitted
        ndim = 1
        flow = object()
        device = local

        def is_fitted(self):
            return True

    if name == "cfg5":
        model = Precomputed()
    else:
        model = wl.make_model(name)
        model.device = local
        model.kernel_path = {"auto": _lib.HB_PATH_AUTO, "simt": _lib.HB_PATH_SIMT,
                             "tcgen05": _lib.HB_PATH_TCGEN05}[args.path]

    a, b = wl.device_data(name, nchains, per_chain, dev, seed=rank)
    start = [i * per_chain for i in range(nchains + 1)]
    torch.cuda.synchronize()

    def one_pass(ev_holder):
        ev = hm.Evidence(total_chains, model)
        ev.device = local
        if name == "cfg5":
            ev.add_lnpred(a, b, start, chain_offset=rank * nchains, comm=comm)
        else:
            ev.add_chains(hm.Chains.from_arrays(a, b, start), chain_offset=rank * nchains, comm=comm)
        ev_holder[:] = [ev]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    holder = []
    for _ in range(args.warmup):
        one_pass(holder)
    lib.hb_set_kernel_timing(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = lib.hb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record()
    for _ in range(args.steps):
        one_pass(holder)
        kernel_ms.append(lib.hb_last_kernel_ms())
    e1.record()
    barrier()
    launches = lib.hb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    lib.hb_set_kernel_timing(0)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ev = holder[0]
    value = world * n_local * args.steps / (ms * 1e-3)

    # ---- e2e: public API on pinned HOST chains, copies inside the timed region ------------
    e2e = None
    if not args.no_e2e:
        rows = min(n_local, int(1.0e8 // world))
        rows = (rows // per_chain) * per_chain or per_chain
        ch_e = rows // per_chain
        ha = torch.empty(a[:rows].shape, dtype=a.dtype).pin_memory()
        hb_ = torch.empty(b[:rows].shape, dtype=b.dtype).pin_memory()
        ha.copy_(a[:rows])
        hb_.copy_(b[:rows])
        xa, xb = ha.numpy(), hb_.numpy()
        st_e = start[:ch_e + 1]

        def e2e_pass():
            evh = hm.Evidence(ch_e * world, model)
            evh.device = local
            if name == "cfg5":
                evh.add_lnpred(xa, xb, st_e, chain_offset=rank * ch_e, comm=comm)
            else:
                evh.add_chains(hm.Chains.from_arrays(xa, xb, st_e), chain_offset=rank * ch_e, comm=comm)
            return evh.compute_ln_evidence()  # D2H read of the step's result

        e2e_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            res = e2e_pass()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        nst = max(1, min(args.steps, 3))
        e2e = {"value": world * rows * nst / dt, "unit": "samples/s",
               "h2d_bytes_per_step": int(xa.nbytes + xb.nbytes), "d2h_bytes_per_step": int(8 * (4 * ch_e * world + 40)),
               "samples_per_step_per_gpu": int(rows), "ln_evidence": float(res[0])}
        del ha, hb_

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = load_peaks()
    kms = float(np.mean([k for k in kernel_ms if k > 0])) if any(k > 0 for k in kernel_ms) else None
    if name == "cfg5":
        ach = wl.BYTES_PER_SAMPLE[name] * n_local / (kms * 1e-3) / 1e9 if kms else None
        peak = float(peaks["hbm_gbs"])
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None,
                # ncu --set full (profiles/r1_prof_acc_r1_raw_summary.csv): dram read+write per sample
                "traffic": 8.0053 * n_local, "traffic_unit": "bytes per launch (ncu dram bytes/sample x samples)",
                "kernel": "accumulate_precomputed_kernel", "kernel_ms": kms,
                "peak_source": "HBM copy bandwidth, of %s" % peak_kind}
    elif name == "cfg4" and args.path != "simt":
        ach = wl.FLOP_PER_SAMPLE[name] * n_local / (kms * 1e-3) / 1e12 if kms else None
        peak = float(peaks["bf16_tflops"]) / 2.0
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": (ach / peak) if ach else None,
                # ncu --set full (profiles/r1_prof_tc_r1l_raw_summary.csv): 262.5 B of DRAM traffic per sample
                # (read 260.4 + write 2.1; algorithmic 260 B): the kernel is not memory bound
                "traffic": 262.5 * n_local, "traffic_unit": "DRAM bytes per launch (ncu bytes/sample x samples)",
                "kernel": "realnvp_tc_kernel", "kernel_ms": kms,
                "peak_source": "TF32 dense = bf16 burst / 2, of %s (the roofline BASELINE.json names; the kernel issues "
                               "kind::f16 MMAs, against the measured 16-bit peak the fraction is half of frac)" % peak_kind,
                "frac_of_16bit_peak": (ach / (2.0 * peak)) if ach else None,
                # issued: 3.43x the algorithmic FLOP, all as kind::f16 (three FP16 products on compensated splits of
                # activations and weights, GEMM1 K 32 -> 40); a tcgen05.mma costs max(~45, N/2) cycles whatever its kind
                # (profiles/r1_mma_probe.txt).  Three tiles in flight leave 32 accumulator columns per tile, so the 2
                # scaled layers run GEMM2 in two N = 32 passes: 8 x 64.4 + (2 x 48 + 4 x 24)/6 x 45 = 1955 tensor-pipe
                # cycles per 128-sample tile and layer => pipe bound 33.5 ms for 1e8 samples at 1.85 GHz, i.e.
                # frac <= 0.42 for this instruction mix
                "issued_mma_flop_factor": 3.43, "issued_tf32_time_factor": 1.71,
                "mma_instructions_per_tile_layer": 40, "tensor_pipe_cycles_per_tile_layer": 1955,
                "algorithmic_flop_per_sample": wl.FLOP_PER_SAMPLE[name]}
    else:
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
            # SURVEY.md 8(d) counts the reference's dense conditioner (all D rows of the (D, 3K+1) output, two
            # separate last Denses).  The kernel pre-multiplies the last MLP Dense with the output Dense (no
            # activation between them, harmonic/flows.py:357,434-438) and evaluates only the transformed rows,
            # so it EXECUTES fewer FLOPs than the algorithmic count: frac can exceed 1; frac_executed is the
            # fraction of the FP32 peak actually sustained (the spline's MUFU work is not counted in either).
            L, h = (8, 64) if name == "cfg2" else (3, 32)
            ex = L * 2.0 * (2 * 2 + 2 * h + h * 25)
            roof["executed_flop_per_sample"] = ex
            roof["frac_executed"] = roof["frac"] * ex / wl.FLOP_PER_SAMPLE[name]

    cpu = None
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        cn, cp = cpu_sample_shape(name)
        v, dt, _ = cpu_reference_pass(name, None if name == "cfg5" else model, cn, cp, seed=0, threads=threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "%d chains x %d samples (%.1f s) of the same workload, float32 restatement in oracle/"
                         % (cn, cp, dt)}

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 3-product split on tcgen05 (f32-equivalent products, f32 accumulate)" if (name == "cfg4" and args.path != "simt") else "f32",
        "data": "synthetic", "config": workload_config(name, world, nchains, per_chain),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        "result": {"ln_evidence_inv": float(ev.ln_evidence_inv), "ln_evidence_inv_var": float(ev.ln_evidence_inv_var),
                   "n_eff": float(ev.n_eff)},
    }
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
