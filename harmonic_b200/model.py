"""Posterior-model "plugin" side of the hot path: drop-in mirrors of
``harmonic.model.RealNVPModel`` / ``RQSplineModel`` (harmonic/model.py:97-404).

Same constructor signatures, same ``predict`` contract and error behaviour
(harmonic/model.py:196-237), same duck-typed surface that ``Evidence`` relies on (``ndim``,
``is_fitted()``, ``predict``, attribute ``flow``; harmonic/evidence.py:56-63,90).  ``predict`` runs
on the B200 through libharmonic_b200 (ctypes); there is no CPU path.  Training stays with the
reference's JAX/flax code: ``fit`` delegates to it when it is importable, otherwise load trained
weights with ``load_params`` / ``from_reference``.
"""
import abc
from typing import Sequence

import cloudpickle
import numpy as np

from . import _lib
from . import export


class Model(metaclass=abc.ABCMeta):
    """harmonic/model_abstract.py:6-100."""

    @abc.abstractmethod
    def __init__(self, ndim: int):
        pass

    @abc.abstractmethod
    def fit(self, X, Y):
        pass

    @abc.abstractmethod
    def predict(self, x):
        pass

    def is_fitted(self) -> bool:
        return self.fitted

    def serialize(self, filename):
        with open(filename, "wb") as fh:
            cloudpickle.dump(self, fh)

    @classmethod
    def deserialize(cls, filename):
        with open(filename, "rb") as fh:
            return cloudpickle.load(fh)


class _State:
    """Stand-in for flax TrainState: only ``params`` is consumed by the hot path."""

    def __init__(self, params):
        self.params = params


class FlowSpec:
    """What ``model.flow`` holds here: the hyper-parameters that shape the network (the reference
    holds the flax module).  Its presence selects Evidence's batched branch."""

    def __init__(self, kind, **hyper):
        self.kind = kind
        self.hyper = hyper

    def __repr__(self):
        return "FlowSpec(%s, %s)" % (self.kind, self.hyper)


class FlowModel(Model):
    def __init__(self, ndim_in: int, learning_rate: float = 0.001, momentum: float = 0.9,
                 standardize: bool = False, temperature: float = 0.8):
        if ndim_in < 1:
            raise ValueError("Dimension must be greater than 0.")
        self.ndim = ndim_in
        self.fitted = False
        self.state = None
        self.learning_rate = learning_rate
        self.momentum = momentum
        self.standardize = standardize
        self.temperature = temperature
        self.flow = None
        self.pre_offset = None
        self.pre_amp = None
        self.device = 0
        self.kernel_path = _lib.HB_PATH_AUTO
        self._handle = None
        self._handle_key = None

    # ------------------------------------------------------------------ weights in
    def load_params(self, params, pre_offset=None, pre_amp=None):
        """Install trained weights: ``params`` is ``model.state.params`` of the reference
        (nested dict); ``pre_offset`` / ``pre_amp`` as set by the reference
        ``fit`` when ``standardize`` (harmonic/model.py:174-183)."""
        if self.flow is None:
            raise NotImplementedError(
                "This method cannot be used in the FlowModel class directly. Use a class with a "
                "specific flow implemented (RealNVPModel, RQSplineModel).")
        if self.standardize and (pre_offset is None or pre_amp is None):
            raise ValueError("standardize=True needs pre_offset and pre_amp")
        self.state = _State(_to_numpy_tree(params))
        self.pre_offset = None if pre_offset is None else np.asarray(pre_offset, dtype=np.float32)
        self.pre_amp = None if pre_amp is None else np.asarray(pre_amp, dtype=np.float32)
        self._flat_weights()  # validates names and shapes
        self.fitted = True
        self._release()
        return self

    @classmethod
    def from_reference(cls, ref_model):
        """Build from a fitted ``harmonic.model.RealNVPModel`` / ``RQSplineModel`` (weights exported
        once; harmonic/model.py:186-192)."""
        raise NotImplementedError

    def fit(self, X, batch_size: int = 64, epochs: int = 3, key=None, verbose: bool = False):
        """Training is NOT part of this library (it stays in the reference's JAX/flax code).  When
        ``harmonic`` + ``jax`` are importable this trains the reference model and exports its
        weights; otherwise it raises."""
        if self.flow is None:
            raise NotImplementedError(
                "This method cannot be used in the FlowModel class directly. Use a class with a "
                "specific flow implemented (RealNVPModel, RQSplineModel).")
        if X.shape[1] != self.ndim:
            raise ValueError("X second dimension not the same as ndim.")
        try:
            import jax  # noqa: F401
            import harmonic.model as ref_md
        except Exception as exc:  # pragma: no cover - depends on the environment
            raise ImportError(
                "harmonic_b200 accelerates predict/add_chains only; flow training stays in the "
                "reference harmonic (JAX/flax), which is not importable here (%s). Train there and "
                "call load_params(model.state.params, model.pre_offset, model.pre_amp)." % exc)
        ref = self._make_reference(ref_md)  # pragma: no cover
        kw = {} if key is None else {"key": key}  # pragma: no cover
        ref.fit(X, batch_size=batch_size, epochs=epochs, verbose=verbose, **kw)  # pragma: no cover
        self.load_params(ref.state.params, getattr(ref, "pre_offset", None),
                         getattr(ref, "pre_amp", None))  # pragma: no cover

    # ------------------------------------------------------------------ handle management
    def _desc(self):
        raise NotImplementedError

    def _flat_weights(self):
        raise NotImplementedError

    def _get_handle(self, device=None):
        device = self.device if device is None else device
        key = (device, self.kernel_path)
        if self._handle is not None and self._handle_key == key:
            return self._handle
        self._release()
        lib = _lib.require_gpu()
        w = self._flat_weights()
        desc = self._desc()
        h = _lib.C.c_void_p()
        _lib.check(lib.hb_flow_create(_lib.C.byref(desc), w.ctypes.data, w.size, device, _lib.C.byref(h)))
        if self.kernel_path != _lib.HB_PATH_AUTO:
            try:
                _lib.check(lib.hb_flow_set_path(h, self.kernel_path))
            except Exception:
                lib.hb_flow_destroy(h)
                raise
        self._handle, self._handle_key = h, key
        self._guard_rows = self._guard_calls = 0
        return h

    def _warn_guard(self, h):
        """The tensor path's range guard (include/harmonic_b200.h, hb_flow_guard_stats): say so when a call was
        evaluated again on the CUDA cores because samples lay outside the range the FP16-split arithmetic serves."""
        if self.kernel_path != _lib.HB_PATH_AUTO:
            return
        rows, calls = _lib.C.c_int64(0), _lib.C.c_int64(0)
        _lib.check(_lib.load().hb_flow_guard_stats(h, _lib.C.byref(rows), _lib.C.byref(calls)))
        if calls.value > self._guard_calls:
            from . import logs
            logs.warning_log(
                "harmonic_b200: %d sample rows outside the tensor-core path's range; the call was evaluated on the "
                "CUDA cores (float32 FMA arithmetic) instead." % (rows.value - self._guard_rows))
        self._guard_rows, self._guard_calls = rows.value, calls.value

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().hb_flow_destroy(self._handle)
            except Exception:
                pass
        self._handle = None
        self._handle_key = None

    def __del__(self):
        self._release()

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_handle"] = None  # ctypes handles never travel; re-created lazily
        st["_handle_key"] = None
        return st

    def _check_temperature(self):
        if self.temperature > 1:
            raise ValueError("Scaling must not be greater than 1.")
        if self.temperature <= 0:
            raise ValueError("Scaling must be positive.")

    # ------------------------------------------------------------------ harmonic/model.py:196-237
    def predict(self, x):
        """ln phi at batched x (batch, ndim) -> (batch,) float32; a 1-D x gives a scalar
        (harmonic/model.py:229-231).  numpy in -> numpy out; CUDA tensor in -> torch CUDA tensor out."""
        self._check_temperature()
        if not self.fitted:
            raise ValueError("Model not fitted.")
        lib = _lib.require_gpu()
        is_dev = hasattr(x, "__cuda_array_interface__")
        if is_dev:
            one_d = len(x.shape) == 1
            x2 = x.reshape(1, -1) if one_d else x
            ref = _lib.ArrayRef(x2, want_2d=True)
            if ref.shape[1] != self.ndim:
                raise ValueError("x second dimension not the same as ndim.")
            import torch
            dev = ref.device if ref.device is not None else self.device
            out = torch.empty(ref.shape[0], dtype=torch.float32, device="cuda:%d" % dev)
            h = self._get_handle(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.hb_flow_predict(h, self.temperature, ref.ptr, ref.dtype, _lib.HB_MEM_DEVICE,
                                           ref.shape[0], ref.ld, out.data_ptr(), _lib.HB_MEM_DEVICE, stream))
            self._warn_guard(h)
            return out[0] if one_d else out
        xa = np.asarray(x)
        one_d = xa.ndim == 1
        if one_d:
            xa = xa[None, :]
        ref = _lib.ArrayRef(xa, want_2d=True)
        if ref.shape[1] != self.ndim:
            raise ValueError("x second dimension not the same as ndim.")
        out = np.empty(ref.shape[0], dtype=np.float32)
        h = self._get_handle()
        _lib.check(lib.hb_flow_predict(h, self.temperature, ref.ptr, ref.dtype, _lib.HB_MEM_HOST,
                                       ref.shape[0], ref.ld, out.ctypes.data, _lib.HB_MEM_HOST, None))
        self._warn_guard(h)
        return out[0] if one_d else out

    # ------------------------------------------------------------------ harmonic/model.py:239-275
    def sample(self, n_sample, rng_key=None, eps=None):
        """Samples from the trained flow at ``self.temperature`` -> (n_sample, ndim) float32.

        The reference draws the base Gaussian with a JAX PRNG key; here the standard-normal draws come from
        torch's Philox generator on the GPU, seeded from ``rng_key`` (an int, or any array whose integer
        contents are folded into a seed; default 0), so streams differ from JAX's but the transform is the
        same.  ``eps`` (n_sample, ndim) supplies the standard-normal draws explicitly (numpy -> numpy out,
        CUDA tensor -> torch CUDA tensor out): the call is then deterministic."""
        self._check_temperature()
        if not self.fitted:
            raise ValueError("Model not fitted.")
        lib = _lib.require_gpu()
        n_sample = int(n_sample)
        if n_sample < 0:
            raise ValueError("n_sample must be non-negative.")
        if eps is not None and not hasattr(eps, "__cuda_array_interface__"):
            ea = np.asarray(eps)
            if ea.ndim != 2 or ea.shape[1] != self.ndim or ea.shape[0] != n_sample:
                raise ValueError("eps must have shape (n_sample, ndim).")
            ref = _lib.ArrayRef(ea, want_2d=True)
            out = np.empty((n_sample, self.ndim), dtype=np.float32)
            _lib.check(lib.hb_flow_sample(self._get_handle(), self.temperature, ref.ptr, ref.dtype, n_sample, ref.ld,
                                          out.ctypes.data, _lib.HB_MEM_HOST, None))
            return out
        import torch
        if eps is None:
            dev = self.device
            seed = 0 if rng_key is None else int(np.sum(np.asarray(rng_key, dtype=np.uint64).ravel()
                                                        * np.uint64(0x9E3779B97F4A7C15)) % np.uint64(2**63))
            g = torch.Generator(device="cuda:%d" % dev)
            g.manual_seed(seed)
            e = torch.randn((n_sample, self.ndim), generator=g, device="cuda:%d" % dev, dtype=torch.float32)
            to_host = True
        else:
            e, to_host = eps, False
            if tuple(e.shape) != (n_sample, self.ndim):
                raise ValueError("eps must have shape (n_sample, ndim).")
        ref = _lib.ArrayRef(e, want_2d=True)
        dev = ref.device if ref.device is not None else self.device
        out = torch.empty((n_sample, self.ndim), dtype=torch.float32, device="cuda:%d" % dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.hb_flow_sample(self._get_handle(dev), self.temperature, ref.ptr, ref.dtype, n_sample, ref.ld,
                                      out.data_ptr(), _lib.HB_MEM_DEVICE, stream))
        return out.cpu().numpy() if to_host else out


def _to_numpy_tree(tree):
    if isinstance(tree, dict) or hasattr(tree, "items"):
        return {k: _to_numpy_tree(v) for k, v in tree.items()}
    return np.asarray(tree, dtype=np.float32)


class RealNVPModel(FlowModel):
    """harmonic/model.py:283-337."""

    def __init__(self, ndim_in: int, n_scaled_layers: int = 2, n_unscaled_layers: int = 4,
                 learning_rate: float = 0.001, momentum: float = 0.9, standardize: bool = False,
                 temperature: float = 0.8):
        if n_scaled_layers <= 0:
            raise ValueError("Number of scaled layers must be greater than 0.")
        FlowModel.__init__(self, ndim_in, learning_rate, momentum, standardize, temperature)
        self.n_scaled_layers = n_scaled_layers
        self.n_unscaled_layers = n_unscaled_layers
        self.flow = FlowSpec("realnvp", n_features=ndim_in, n_scaled_layers=n_scaled_layers,
                             n_unscaled_layers=n_unscaled_layers)

    def _desc(self):
        d = _lib.FlowDesc()
        d.kind = _lib.HB_FLOW_REALNVP
        d.ndim = self.ndim
        d.n_scaled = self.n_scaled_layers
        d.n_unscaled = self.n_unscaled_layers
        d.standardize = 1 if self.standardize else 0
        return d

    def _flat_weights(self):
        return export.flatten_realnvp(self.state.params, self.ndim, self.n_scaled_layers,
                                      self.n_unscaled_layers, self.pre_offset, self.pre_amp)

    def _make_reference(self, ref_md):  # pragma: no cover
        return ref_md.RealNVPModel(self.ndim, self.n_scaled_layers, self.n_unscaled_layers,
                                   self.learning_rate, self.momentum, self.standardize, self.temperature)

    @classmethod
    def from_reference(cls, ref_model):
        m = cls(ref_model.ndim, ref_model.n_scaled_layers, ref_model.n_unscaled_layers,
                ref_model.learning_rate, ref_model.momentum, ref_model.standardize, ref_model.temperature)
        return m.load_params(ref_model.state.params, getattr(ref_model, "pre_offset", None),
                             getattr(ref_model, "pre_amp", None))

    def oracle_spec(self):
        """The dict oracle.flows.predict understands (tests only)."""
        return dict(kind="realnvp", params=self.state.params, temperature=self.temperature,
                    standardize=self.standardize, pre_offset=self.pre_offset, pre_amp=self.pre_amp,
                    n_scaled_layers=self.n_scaled_layers, n_unscaled_layers=self.n_unscaled_layers)


class RQSplineModel(FlowModel):
    """harmonic/model.py:345-404."""

    def __init__(self, ndim_in: int, n_layers: int = 8, n_bins: int = 8,
                 hidden_size: Sequence[int] = [64, 64], spline_range: Sequence[float] = (-10.0, 10.0),
                 standardize: bool = False, learning_rate: float = 0.001, momentum: float = 0.9,
                 temperature: float = 0.8):
        FlowModel.__init__(self, ndim_in, learning_rate, momentum, standardize, temperature)
        self.n_layers = n_layers
        self.hidden_size = hidden_size
        self.n_bins = n_bins
        self.spline_range = spline_range
        self.flow = FlowSpec("rqspline", n_features=ndim_in, num_layers=n_layers,
                             hidden_size=list(hidden_size), num_bins=n_bins, spline_range=tuple(spline_range))

    def _desc(self):
        d = _lib.FlowDesc()
        d.kind = _lib.HB_FLOW_RQSPLINE
        d.ndim = self.ndim
        d.n_layers = self.n_layers
        d.n_bins = self.n_bins
        if len(self.hidden_size) > _lib.HB_MAX_HIDDEN:
            raise NotImplementedError("at most %d hidden layers are supported" % _lib.HB_MAX_HIDDEN)
        d.n_hidden = len(self.hidden_size)
        for i, h in enumerate(self.hidden_size):
            d.hidden[i] = int(h)
        d.range_min = float(self.spline_range[0])
        d.range_max = float(self.spline_range[1])
        d.standardize = 1 if self.standardize else 0
        return d

    def _flat_weights(self):
        return export.flatten_rqspline(self.state.params, self.ndim, self.n_layers, self.n_bins,
                                       self.hidden_size, self.pre_offset, self.pre_amp)

    def _make_reference(self, ref_md):  # pragma: no cover
        return ref_md.RQSplineModel(self.ndim, self.n_layers, self.n_bins, self.hidden_size,
                                    self.spline_range, self.standardize, self.learning_rate,
                                    self.momentum, self.temperature)

    @classmethod
    def from_reference(cls, ref_model):
        m = cls(ref_model.ndim, ref_model.n_layers, ref_model.n_bins, ref_model.hidden_size,
                ref_model.spline_range, ref_model.standardize, ref_model.learning_rate,
                ref_model.momentum, ref_model.temperature)
        return m.load_params(ref_model.state.params, getattr(ref_model, "pre_offset", None),
                             getattr(ref_model, "pre_amp", None))

    def oracle_spec(self):
        return dict(kind="rqspline", params=self.state.params, temperature=self.temperature,
                    standardize=self.standardize, pre_offset=self.pre_offset, pre_amp=self.pre_amp,
                    n_layers=self.n_layers, n_bins=self.n_bins, hidden_size=list(self.hidden_size),
                    spline_range=tuple(self.spline_range))
