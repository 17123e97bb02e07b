"""ctypes binding of libharmonic_b200.so (the C ABI in include/harmonic_b200.h).

There is no CPU fallback: every compute entry point raises if the library or a B200 is
missing.  The library is built in-tree by ``harmonic_b200/build.py`` (nvcc, sm_100a).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# HARMONIC_B200_LIB selects another build of the same library (kernel A/B measurements); always a CUDA build
LIB_PATH = os.environ.get("HARMONIC_B200_LIB") or os.path.join(HERE, "libharmonic_b200.so")

HB_FLOW_REALNVP, HB_FLOW_RQSPLINE = 1, 2
HB_F32, HB_F64 = 0, 1
HB_MEM_HOST, HB_MEM_DEVICE = 0, 1
HB_PATH_AUTO, HB_PATH_SIMT, HB_PATH_TCGEN05, HB_PATH_SIMT_GENERIC = 0, 1, 2, 3
HB_ACC_NEW_CALL = 1
HB_MERGE_OVERWRITE = 1
HB_MAX_HIDDEN = 8
HB_ERR_INVALID, HB_ERR_CUDA, HB_ERR_UNSUPPORTED, HB_ERR_NO_DEVICE = -1, -2, -3, -4


class FlowDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("ndim", C.c_int32), ("n_scaled", C.c_int32), ("n_unscaled", C.c_int32),
        ("n_layers", C.c_int32), ("n_bins", C.c_int32), ("n_hidden", C.c_int32),
        ("hidden", C.c_int32 * HB_MAX_HIDDEN), ("range_min", C.c_float), ("range_max", C.c_float),
        ("standardize", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff",
        "shift_value", "lnargmax", "lnargmin", "lnprobmax", "lnprobmin", "lnpredictmax",
        "lnpredictmin", "nsamples_total", "mean_nsamples_eff", "guard_trips", "reserved")]


# name -> (restype, argtypes); mirrors include/harmonic_b200.h one to one
SIGNATURES = {
    "hb_version": (C.c_int, []),
    "hb_last_error": (C.c_char_p, []),
    "hb_device_count": (C.c_int, []),
    "hb_flow_weight_count": (C.c_size_t, [C.POINTER(FlowDesc)]),
    "hb_flow_create": (C.c_int, [C.POINTER(FlowDesc), C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "hb_flow_destroy": (C.c_int, [C.c_void_p]),
    "hb_flow_set_path": (C.c_int, [C.c_void_p, C.c_int]),
    "hb_flow_uses_tcgen05": (C.c_int, [C.c_void_p]),
    "hb_flow_guard_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "hb_flow_predict": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_int, C.c_void_p]),
    "hb_flow_sample": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                 C.c_int, C.c_void_p]),
    "hb_accum_create": (C.c_int, [C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "hb_accum_destroy": (C.c_int, [C.c_void_p]),
    "hb_accum_reset": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hb_pool_trim": (C.c_int, []),
    "hb_accumulate_flow": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int64,
                                     C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "hb_accumulate_precomputed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "hb_accum_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_accum_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_accum_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "hb_accum_packed_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "hb_accum_merge_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "hb_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.POINTER(Result), C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_finalize_arrays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_finalize_realspace": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                        C.POINTER(Result), C.c_void_p]),
    "hb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "hb_host_free": (C.c_int, [C.c_void_p]),
    "hb_device_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "hb_device_free": (C.c_int, [C.c_void_p]),
    "hb_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hb_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hb_stream_sync": (C.c_int, [C.c_void_p]),
    "hb_launch_count": (C.c_int64, []),
    "hb_set_kernel_timing": (C.c_int, [C.c_int]),
    "hb_last_kernel_ms": (C.c_double, [C.c_int]),
    "hb_bind_host_to_device_node": (C.c_int, [C.c_int]),
    "hb_host_copy_gbs": (C.c_double, [C.c_size_t, C.c_int]),
}

_lib = None


class HarmonicB200Error(RuntimeError):
    pass


def load():
    """dlopen the in-tree library (never a CPU substitute)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HarmonicB200Error(
            "libharmonic_b200.so is not built (run `python harmonic_b200/build.py`); "
            "harmonic_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == 0:
        return
    msg = load().hb_last_error().decode("utf-8", "replace")
    if rc == HB_ERR_INVALID:
        raise ValueError(msg)
    if rc == HB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise HarmonicB200Error(msg)


_gpu_ok = False


def require_gpu():
    global _gpu_ok
    lib = load()
    if _gpu_ok:
        return lib
    if lib.hb_device_count() < 1:
        raise HarmonicB200Error(
            "no sm_100 (B200) device visible: harmonic_b200 has no CPU fallback")
    _gpu_ok = True
    return lib


# ------------------------------------------------------------------ array ingestion helpers
class ArrayRef:
    """Pointer + dtype + location of a numpy array or of any CUDA tensor exposing
    ``__cuda_array_interface__`` (torch, cupy): zero-copy for device data."""

    __slots__ = ("ptr", "dtype", "mem", "shape", "ld", "keep", "device")

    def __init__(self, a, want_2d=False, allow_f64=True):
        self.device = None
        if hasattr(a, "__cuda_array_interface__"):
            cai = a.__cuda_array_interface__
            ts = cai["typestr"]
            if ts not in ("<f4", "<f8"):
                raise ValueError("device arrays must be float32 or float64")
            shape = tuple(cai["shape"])
            strides = cai.get("strides")
            item = 4 if ts == "<f4" else 8
            if want_2d:
                if len(shape) != 2:
                    raise ValueError("expected a 2-D array")
                if strides is None:
                    ld = shape[1]
                else:
                    if strides[1] != item or strides[0] % item:
                        raise ValueError("device samples must be row-major")
                    ld = strides[0] // item
            else:
                if strides is not None and len(shape) == 1 and shape[0] > 1 and strides[0] != item:
                    raise ValueError("device vector must be contiguous")
                ld = 1
            self.ptr = cai["data"][0]
            self.dtype = HB_F32 if ts == "<f4" else HB_F64
            self.mem = HB_MEM_DEVICE
            self.shape = shape
            self.ld = ld
            self.keep = a
            dev = getattr(a, "device", None)
            self.device = getattr(dev, "index", None)
        else:
            arr = np.asarray(a)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            arr = np.ascontiguousarray(arr)
            if want_2d and arr.ndim != 2:
                raise ValueError("expected a 2-D array")
            self.ptr = arr.ctypes.data
            self.dtype = HB_F32 if arr.dtype == np.float32 else HB_F64
            self.mem = HB_MEM_HOST
            self.shape = arr.shape
            self.ld = arr.shape[1] if arr.ndim == 2 else 1
            self.keep = arr
