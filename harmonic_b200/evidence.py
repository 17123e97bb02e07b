"""Drop-in mirror of ``harmonic.evidence`` (harmonic/evidence.py:11-664) running on a B200.

``Evidence.add_chains`` hands the chains to ONE fused C-ABI call (flow evaluation + lnarg + per-chain
max-rescaled logsumexp + NaN counts + extrema, ``hb_accumulate_flow``) and one finalise kernel
(``hb_finalize`` = ``process_run``).  The per-chain state is kept shift-free as (m, s) pairs, which
merge exactly across calls (streaming / checkpointing, harmonic/evidence.py:499-531) and across
GPUs (one all-gather of O(nchains) doubles).  Public attributes, signatures and exceptions follow
the reference.  There is no CPU path.
"""
from enum import Enum

import cloudpickle
import numpy as np

from . import _lib
from . import logs as lg


class Shifting(Enum):
    """harmonic/evidence.py:11-20."""

    MEAN_SHIFT = 1
    MAX_SHIFT = 2
    MIN_SHIFT = 3
    ABS_MAX_SHIFT = 4


class Evidence:
    def __init__(self, nchains: int, model, shift=Shifting.MEAN_SHIFT):
        if nchains < 1:
            raise ValueError("nchains must be greater than 0.")
        if model.ndim < 1:
            raise ValueError("ndim must be greater than 0.")
        if not model.is_fitted():
            raise ValueError("Model not fitted.")

        self.nchains = nchains
        # running_sum, nsamples_per_chain, nsamples_eff_per_chain, ln_evidence_inv_per_chain are properties:
        # after add_chains they stay on the device until first read (only the 128-byte result crosses PCIe
        # per call)
        self._pc = {"running_sum": np.zeros(nchains), "nsamples_per_chain": np.zeros(nchains),
                    "nsamples_eff_per_chain": np.zeros(nchains), "ln_evidence_inv_per_chain": None}
        self._pc_src = None  # accumulator handle holding newer per-chain arrays than _pc
        self._mean_nsamples_eff = 0.0
        self.guard_trips = 0

        self.ndim = model.ndim
        self.kurtosis = 0.0
        self.n_eff = 0

        self.ln_evidence_inv = 0.0
        self.ln_evidence_inv_var = 0.0
        self.ln_evidence_inv_var_var = 0.0
        self.ln_kurtosis = 0.0

        self.shift = shift
        self.shift_set = False
        self.shift_value = 0.0

        self.chains_added = False

        self.model = model
        self.batch_calculation = hasattr(self.model, "flow")

        self.lnargmax = -np.inf
        self.lnargmin = np.inf
        self.lnprobmax = -np.inf
        self.lnprobmin = np.inf
        self.lnpredictmax = -np.inf
        self.lnpredictmin = np.inf

        # device-side streaming state (never pickled) and its host image (pickled)
        self.device = getattr(model, "device", 0)
        self._acc = None
        self._acc_global = None
        self._partials = None  # (nchains, 4): m, s, n, n_nan of THIS process' contributions
        self._gather = None    # (send view, recv buffer, key) of the NCCL all-gather, kept across calls

    # ------------------------------------------------------------------ lazily fetched per-chain arrays
    def _fetch_per_chain(self):
        src = self._pc_src
        if src is None:
            return
        self._pc_src = None
        arrs = [np.empty(self.nchains, dtype=np.float64) for _ in range(4)]
        _lib.check(_lib.load().hb_finalize_arrays(src, *[a.ctypes.data for a in arrs], self._stream()))
        for k, a in zip(("ln_evidence_inv_per_chain", "running_sum", "nsamples_per_chain",
                         "nsamples_eff_per_chain"), arrs):
            self._pc[k] = a

    def _stream(self):
        return _current_stream(self.device)

    # ------------------------------------------------------------------ harmonic/evidence.py:100-123
    def set_shift(self, shift_value: float):
        if not np.isfinite(shift_value):
            raise ValueError("Shift must be a number")
        if self.shift_set:
            raise ValueError("Cannot define multiple shifts!")
        self.shift_value = shift_value
        self.shift_set = True

    # ------------------------------------------------------------------ handles / pickling
    def _get_acc(self):
        if self._acc is None:
            lib = _lib.require_gpu()
            h = _lib.C.c_void_p()
            _lib.check(lib.hb_accum_create(self.nchains, self.device, _lib.C.byref(h)))
            self._acc = h
            if self._partials is not None:  # restored from a checkpoint
                p = np.ascontiguousarray(self._partials, dtype=np.float64)
                _lib.check(lib.hb_accum_merge(h, p.ctypes.data, None, self._stream()))
        return self._acc

    def _release(self):
        self._pc_src = None
        self._gather = None
        for name in ("_acc", "_acc_global"):
            h = getattr(self, name, None)
            if h is not None:
                try:
                    _lib.load().hb_accum_destroy(h)
                except Exception:
                    pass
            setattr(self, name, None)

    def __del__(self):
        self._release()

    def partials(self):
        """(nchains, 4) float64 array of this process' per-chain state (m, s, n, n_nan): what is
        pickled, and what crosses NVLink in the multi-GPU all-gather."""
        if self._acc is not None:
            part = np.empty((self.nchains, 4), dtype=np.float64)
            _lib.check(_lib.load().hb_accum_export(self._acc, part.ctypes.data, None, self._stream()))
            self._partials = part
        return self._partials

    def __getstate__(self):
        self.partials()  # pull the device-side streaming state into the picklable image
        self._fetch_per_chain()
        st = dict(self.__dict__)
        st["_acc"] = None
        st["_acc_global"] = None
        st["_gather"] = None
        st["_pc_src"] = None
        return st

    # ------------------------------------------------------------------ harmonic/evidence.py:125-189
    def process_run(self):
        """Recompute the statistics from the public running totals (``running_sum``,
        ``nsamples_per_chain``, ``shift_value``) with the finalise kernel."""
        lib = _lib.require_gpu()
        S = np.ascontiguousarray(self.running_sum, dtype=np.float64)
        n = np.ascontiguousarray(self.nsamples_per_chain, dtype=np.float64)
        res = _lib.Result()
        per = np.empty(self.nchains, dtype=np.float64)
        _lib.check(lib.hb_finalize_realspace(self.device, self.nchains, S.ctypes.data, n.ctypes.data,
                                             float(self.shift_value), _lib.C.byref(res), per.ctypes.data))
        self._take(res)
        self.ln_evidence_inv_per_chain = per
        self._mean_nsamples_eff = float(np.mean(self.nsamples_eff_per_chain))

    def _take(self, res):
        self.ln_evidence_inv = res.ln_evidence_inv
        self.ln_evidence_inv_var = res.ln_evidence_inv_var
        self.ln_evidence_inv_var_var = res.ln_evidence_inv_var_var
        self.ln_kurtosis = res.ln_kurtosis
        self.kurtosis = float(np.exp(res.ln_kurtosis))
        self.n_eff = res.n_eff

    # ------------------------------------------------------------------ harmonic/evidence.py:215-358
    def _shard_start(self, start_indices, nloc, sharded, chain_offset):
        """Global start_indices array (nchains + 1, int64) with this shard's offsets in place [lo, hi]."""
        lo = int(chain_offset) if sharded else 0
        hi = lo + nloc
        loc = start_indices if isinstance(start_indices, np.ndarray) and start_indices.dtype == np.int64 \
            else np.asarray(start_indices, dtype=np.int64)
        if lo == 0 and hi == self.nchains:
            return np.ascontiguousarray(loc), lo, hi
        start = np.zeros(self.nchains + 1, dtype=np.int64)
        start[lo:hi + 1] = loc
        start[hi + 1:] = start[hi]
        return start, lo, hi

    def add_chains(self, chains, num_slices=None, chain_offset=0, comm=None):
        """Add chains and update the inverse-evidence statistics.

        ``chains`` / ``num_slices`` as in the reference (``num_slices`` only bounded memory there; it
        is validated and otherwise unused because the kernel streams tiles).

        Multi-GPU (additive): every rank builds ``Evidence(nchains_total, model)`` and calls
        ``add_chains(local_chains, chain_offset=first_chain_of_this_rank, comm=group_or_True)`` with
        only ITS chains; per-chain partials are all-gathered (torch.distributed) and every rank ends
        with identical statistics.

        Everything the call launches -- flow + fold kernel(s), merge, (all-gather, merge of the gathered
        blocks,) finalise -- is ordered on ONE CUDA stream (torch's current stream when torch is in use);
        the call ends with a single stream synchronisation that brings back the 128-byte result.
        """
        sharded = comm is not None and comm is not False
        if sharded:
            if chain_offset < 0 or chain_offset + chains.nchains > self.nchains:
                raise ValueError("nchains do not match")
        elif chains.nchains != self.nchains:
            raise ValueError("nchains do not match")
        if chains.ndim != self.ndim:
            raise ValueError("Chains ndim inconsistent")

        X = chains.samples
        Y = chains.ln_posterior
        if num_slices is not None:
            if num_slices > X.shape[0]:
                raise ValueError("Can't split chains into more blocks than there are samples.")

        lib = _lib.require_gpu()
        acc = self._get_acc()
        start_loc = chains.start_array() if hasattr(chains, "start_array") else chains.start_indices
        start, lo, hi = self._shard_start(start_loc, chains.nchains, sharded, chain_offset)
        stream = self._stream()

        if self.batch_calculation:
            if hasattr(self.model, "_check_temperature"):
                self.model._check_temperature()
            xr = _lib.ArrayRef(X, want_2d=True)
            yr = _lib.ArrayRef(Y)
            if xr.mem != yr.mem:
                raise ValueError("samples and ln_posterior must live in the same memory space")
            h = self.model._get_handle(self.device)
            _lib.check(lib.hb_accumulate_flow(
                acc, h, float(self.model.temperature), xr.ptr, xr.dtype, xr.ld, yr.ptr, yr.dtype, xr.mem,
                start.ctypes.data, lo, hi, None, _lib.HB_ACC_NEW_CALL, stream))
        else:
            # user-supplied model without a flow: per-sample predict on the host
            # (harmonic/evidence.py:285-303), then the accumulation-only kernel; float64 travels as float64
            # and lnpred - ln_posterior is formed in double on the device, as numpy does there
            Xh, Yh = np.asarray(X), np.ascontiguousarray(Y, dtype=np.float64)
            lnpred = np.empty(Xh.shape[0], dtype=np.float64)
            for i in range(Xh.shape[0]):
                lnpred[i] = self.model.predict(Xh[i, :])
            lnpred[np.isinf(lnpred)] = np.nan
            _lib.check(lib.hb_accumulate_precomputed(
                acc, lnpred.ctypes.data, _lib.HB_F64, Yh.ctypes.data, _lib.HB_F64, _lib.HB_MEM_HOST,
                start.ctypes.data, lo, hi, _lib.HB_ACC_NEW_CALL, stream))

        self._finish_call(lib, acc, comm if sharded else None, stream)
        self.chains_added = True
        self.check_basic_diagnostic()

    def add_lnpred(self, lnpred, ln_posterior, start_indices, chain_offset=0, comm=None):
        """Accumulation-only entry (additive): fold a precomputed ln phi (float32/float64, host or
        device) with ln_posterior; everything after ``predict`` in harmonic/evidence.py:278-358."""
        lib = _lib.require_gpu()
        acc = self._get_acc()
        sharded = comm is not None and comm is not False
        nloc = len(start_indices) - 1
        if chain_offset < 0 or chain_offset + nloc > self.nchains or (not sharded and nloc != self.nchains):
            raise ValueError("nchains do not match")
        start, lo, hi = self._shard_start(start_indices, nloc, sharded, chain_offset)
        pr, yr = _lib.ArrayRef(lnpred), _lib.ArrayRef(ln_posterior)
        if pr.mem != yr.mem:
            raise ValueError("lnpred and ln_posterior must live in the same memory space")
        stream = self._stream()
        _lib.check(lib.hb_accumulate_precomputed(
            acc, pr.ptr, pr.dtype, yr.ptr, yr.dtype, pr.mem, start.ctypes.data, lo, hi,
            _lib.HB_ACC_NEW_CALL, stream))
        self._finish_call(lib, acc, comm if sharded else None, stream)
        self.chains_added = True
        self.check_basic_diagnostic()

    def _finish_call(self, lib, acc, comm, stream):
        fin = acc
        if comm is not None:
            # the one collective of the path: all-gather of [4*nchains partials | 8 scalars] per rank,
            # device to device over NCCL (host staging only for CPU backends), then one merge kernel that
            # REPLACES the state of the gather target -- all on `stream`, no host synchronisation
            if self._acc_global is None:
                h = _lib.C.c_void_p()
                _lib.check(lib.hb_accum_create(self.nchains, self.device, _lib.C.byref(h)))
                self._acc_global = h
            fin = self._acc_global
            n = 4 * self.nchains + 8
            import torch.distributed as dist
            group = None if comm is True else comm
            world = dist.get_world_size(group)
            if dist.get_backend(group) == "nccl":
                send, recv = self._gather_buffers(lib, acc, n, world)
                dist.all_gather_into_tensor(recv, send, group=group)
                _lib.check(lib.hb_accum_merge_packed(fin, world, recv.data_ptr(), n, _lib.HB_MEM_DEVICE,
                                                     _lib.HB_MERGE_OVERWRITE, stream))
            else:
                send = np.empty(n, dtype=np.float64)
                _lib.check(lib.hb_accum_pack(acc, send.ctypes.data, _lib.HB_MEM_HOST, stream))
                recv = np.ascontiguousarray(_allgather(send, comm))
                _lib.check(lib.hb_accum_merge_packed(fin, world, recv.ctypes.data, n, _lib.HB_MEM_HOST,
                                                     _lib.HB_MERGE_OVERWRITE, stream))
        res = _lib.Result()
        shift_in = float(self.shift_value) if self.shift_set else float("nan")
        # finalise on the same stream; only the 128-byte result crosses PCIe, the per-chain arrays stay on
        # the device until somebody reads them (_fetch_per_chain)
        _lib.check(lib.hb_finalize(fin, self.shift.value, shift_in, _lib.C.byref(res), None, None, None, None,
                                   stream))
        self._pc_src = fin
        if not self.shift_set:
            self.set_shift(res.shift_value)
        self._take(res)
        self._mean_nsamples_eff = res.mean_nsamples_eff
        self.lnargmax, self.lnargmin = res.lnargmax, res.lnargmin
        self.lnprobmax, self.lnprobmin = res.lnprobmax, res.lnprobmin
        self.lnpredictmax, self.lnpredictmin = res.lnpredictmax, res.lnpredictmin
        trips = int(res.guard_trips)
        if trips > self.guard_trips:
            lg.warning_log(
                "RealNVP tensor-core path: {} launch(es) had inputs outside its guaranteed-accuracy range and "
                "were re-evaluated on the CUDA cores (float32 FMA arithmetic).".format(trips - self.guard_trips))
        self.guard_trips = trips

    def _gather_buffers(self, lib, acc, n, world):
        """Send view on the accumulator's packed device block (zero copy) and a persistent receive buffer."""
        import torch
        ptr = _lib.C.c_void_p()
        _lib.check(lib.hb_accum_packed_ptr(acc, _lib.C.byref(ptr), None))
        key = (ptr.value, n, world, self.device)
        if self._gather is None or self._gather[0] != key:
            dev = torch.device("cuda", self.device)
            send = torch.as_tensor(_DevView(ptr.value, n), device=dev)
            recv = _recv_buffer(n * world, dev)
            self._gather = (key, send, recv)
        return self._gather[1], self._gather[2]

    # ------------------------------------------------------------------ harmonic/evidence.py:360-399
    def check_basic_diagnostic(self):
        NSAMPLES_EFF_WARNING_LEVEL = 30
        LNARG_WARNING_LEVEL = 1000.0
        tests_pass = True
        mean_eff = self._mean_nsamples_eff if self._pc_src is not None else np.mean(self.nsamples_eff_per_chain)
        if mean_eff <= NSAMPLES_EFF_WARNING_LEVEL:
            lg.warning_log(
                "Evidence may not be accurate due to low number of effective samples (mean number of "
                "effective samples per chain is {}). Use more samples.".format(mean_eff))
            tests_pass = False
        if (self.lnargmax - self.lnargmin) >= LNARG_WARNING_LEVEL:
            lg.warning_log(
                "Evidence may not be accurate due to large dynamic range. Use model with smaller "
                "support and/or better predictive accuracy.")
            tests_pass = False
        return tests_pass

    # ------------------------------------------------------------------ harmonic/evidence.py:401-436
    def compute_evidence(self):
        self.check_basic_diagnostic()
        with np.errstate(over="ignore"):
            evidence_inv = np.exp(self.ln_evidence_inv)
            evidence_inv_var = np.exp(self.ln_evidence_inv_var)
        if np.isinf(np.nan_to_num(evidence_inv, nan=np.inf)) or np.isinf(
                np.nan_to_num(evidence_inv_var, nan=np.inf)):
            raise ValueError(
                "Evidence is too large to represent in non-log space. Use log-space values instead.")
        common_factor = 1.0 + evidence_inv_var / (evidence_inv**2)
        return (common_factor / evidence_inv, np.sqrt(evidence_inv_var) / (evidence_inv**2))

    # ------------------------------------------------------------------ harmonic/evidence.py:438-459
    def compute_ln_evidence(self):
        self.check_basic_diagnostic()
        ln_x = self.ln_evidence_inv_var - 2.0 * self.ln_evidence_inv
        x = np.exp(ln_x)
        ln_evidence = np.log(1.0 + x) - self.ln_evidence_inv
        ln_evidence_std = 0.5 * self.ln_evidence_inv_var - 2.0 * self.ln_evidence_inv
        return (ln_evidence, ln_evidence_std)

    # ------------------------------------------------------------------ harmonic/evidence.py:461-497
    def compute_ln_inv_evidence_errors(self):
        ln_ratio = 0.5 * self.ln_evidence_inv_var - self.ln_evidence_inv
        ratio = np.exp(ln_ratio)
        if np.abs(ratio - 1.0) > 1e-8:
            with np.errstate(invalid="ignore", divide="ignore"):
                ln_evidence_err_neg = np.log(1.0 - ratio)
        else:
            ln_evidence_err_neg = -np.inf
        ln_evidence_err_pos = np.log(1.0 + ratio)
        return (ln_evidence_err_neg, ln_evidence_err_pos)

    # ------------------------------------------------------------------ harmonic/evidence.py:499-531
    def serialize(self, filename):
        with open(filename, "wb") as fh:
            cloudpickle.dump(self, fh)

    @classmethod
    def deserialize(cls, filename):
        with open(filename, "rb") as fh:
            return cloudpickle.load(fh)


def _per_chain_property(name):
    def get(self):
        self._fetch_per_chain()
        return self._pc[name]

    def set_(self, value):
        self._fetch_per_chain()
        self._pc[name] = value

    return property(get, set_)


for _name in ("running_sum", "nsamples_per_chain", "nsamples_eff_per_chain", "ln_evidence_inv_per_chain"):
    setattr(Evidence, _name, _per_chain_property(_name))


class _DevView:
    """``n`` float64 values at a raw device pointer, importable by torch.as_tensor without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


_RECV = {}


def _recv_buffer(n, dev):
    """Receive buffers of the all-gather are shared by all Evidence objects of the process (one per size)."""
    import torch
    key = (n, dev.index)
    buf = _RECV.get(key)
    if buf is None:
        if len(_RECV) > 8:
            _RECV.clear()
        buf = _RECV[key] = torch.empty(n, dtype=torch.float64, device=dev)
    return buf


def _current_stream(device):
    """torch's current stream on ``device`` when the process uses torch, else the legacy default stream."""
    import sys
    torch = sys.modules.get("torch")
    if torch is None:
        return None
    try:
        return torch.cuda.current_stream(device).cuda_stream
    except Exception:
        return None


def _allgather(vec, comm):
    """All-gather one float64 vector per rank; the only collective of the path (O(nchains))."""
    import torch
    import torch.distributed as dist

    group = None if comm is True else comm
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64))
    if backend == "nccl":
        t = t.cuda()
    out = torch.empty(world * t.numel(), dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.view(world, t.numel()).cpu().numpy()


# ---------------------------------------------------------------------- harmonic/evidence.py:534-664
def _inv_and_var(ev, which, nan_only):
    with np.errstate(over="ignore"):
        inv, var = np.exp(ev.ln_evidence_inv), np.exp(ev.ln_evidence_inv_var)
    bad = (np.isnan(inv) or np.isnan(var)) if nan_only else (
        np.isinf(np.nan_to_num(inv, nan=np.inf)) or np.isinf(np.nan_to_num(var, nan=np.inf)))
    if bad:
        raise ValueError("Evidence for model %d is too large to represent in non-log space. "
                         "Use log-space values instead." % which)
    return inv, var


def _bf_prologue(ev1, ev2):
    if not ev1.chains_added:
        raise ValueError("Evidence for model 1 does not have chains added")
    if not ev2.chains_added:
        raise ValueError("Evidence for model 2 does not have chains added")
    ev1.check_basic_diagnostic()
    ev2.check_basic_diagnostic()


def compute_bayes_factor(ev1, ev2):
    _bf_prologue(ev1, ev2)
    i1, v1 = _inv_and_var(ev1, 1, False)
    i2, v2 = _inv_and_var(ev2, 2, False)
    common_factor = 1.0 + v1 / (i1**2)
    bf12 = i2 / i1 * common_factor
    bf12_std = np.sqrt(i1**2 * v2 + i2**2 * v1) / (i1**2)
    return (bf12, bf12_std)


def compute_ln_bayes_factor(ev1, ev2):
    _bf_prologue(ev1, ev2)
    i1, v1 = _inv_and_var(ev1, 1, True)
    i2, v2 = _inv_and_var(ev2, 2, True)
    common_factor = 1.0 + v1 / (i1**2)
    ln_bf12 = np.log(i2) - np.log(i1) + np.log(common_factor)
    factor = i1**2 * v2 + i2**2 * v1
    ln_bf12_std = 0.5 * np.log(factor) - 2.0 * np.log(i1)
    return (ln_bf12, ln_bf12_std)


# ---------------------------------------------------------------------- multi-model batch (additive)
def add_chains_batch(pairs):
    """Fold several (Evidence, Chains) pairs -- e.g. the two models of a Bayes factor
    (harmonic/evidence.py:534-664 takes two Evidence objects whose chains were added one after the other) -- at the
    same time: every pair gets its own CUDA stream and host thread, so the flow + fold kernels, merges and
    finalises of the models overlap on the GPU wherever one model's launches do not fill it (small-ndim flows,
    short chains), and the host waits once for all of them instead of once per model.  Results are those of the
    separate ``ev.add_chains(chains)`` calls, bit for bit.

    ``pairs``: iterable of ``(evidence, chains)`` or ``(evidence, chains, kwargs_for_add_chains)``.
    """
    import concurrent.futures

    import torch

    jobs = []
    for p in pairs:
        ev, ch = p[0], p[1]
        kw = p[2] if len(p) > 2 else {}
        jobs.append((ev, ch, kw))
    if len({id(j[0]) for j in jobs}) != len(jobs):
        raise ValueError("every Evidence object may appear once per batch")
    if not jobs:
        return

    def work(ev, ch, kw):
        dev = ev.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)  # device-resident chains may still be in flight on the caller's stream
            with torch.cuda.stream(side):
                ev.add_chains(ch, **kw)
            side.synchronize()

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(jobs)) as pool:
        futs = [pool.submit(work, *j) for j in jobs]
        for f in futs:
            f.result()  # re-raises the reference's ValueErrors etc. from the worker


def compute_ln_bayes_factor_batch(ev1, chains1, ev2, chains2):
    """``add_chains`` of both models at once (add_chains_batch), then harmonic's ``compute_ln_bayes_factor``."""
    add_chains_batch([(ev1, chains1), (ev2, chains2)])
    return compute_ln_bayes_factor(ev1, ev2)
