"""Weight export: flax parameter pytrees -> the flat float32 buffer of include/harmonic_b200.h.

The reference keeps trained weights in ``model.state.params`` (harmonic/model.py:186-192), a nested
dict whose names are fixed by flax.  ``flatten_*`` walk that tree, assert
every shape and emit the canonical order the C ABI documents.  ``random_*`` build synthetic trees
with the same names/shapes (JAX is not available in the build image, so tests and benchmarks use
seeded synthetic weights).
"""
import numpy as np

HIDDEN = 128  # nn.Dense(128), harmonic/flows.py:171


def num_masked(ndim):
    """tfp RealNVP(fraction_masked=0.5): int(np.round(ndim * 0.5)) (round-half-even)."""
    return int(np.round(ndim * 0.5))


def _arr(node, name, shape):
    a = np.asarray(node[name], dtype=np.float32)
    if tuple(a.shape) != tuple(shape):
        raise ValueError("parameter %s has shape %s, expected %s" % (name, a.shape, shape))
    return a.reshape(-1)


def _pre(ndim, pre_offset, pre_amp):
    off = np.zeros(ndim, np.float32) if pre_offset is None else np.asarray(pre_offset, np.float32).reshape(ndim)
    amp = np.ones(ndim, np.float32) if pre_amp is None else np.asarray(pre_amp, np.float32).reshape(ndim)
    return [off, amp]


def flatten_realnvp(params, ndim, n_scaled, n_unscaled, pre_offset=None, pre_amp=None):
    d = num_masked(ndim)
    u = ndim - d
    out = _pre(ndim, pre_offset, pre_amp)
    names = ["scaled_layers_%d" % i for i in range(n_scaled)] + [
        "unscaled_layers_%d" % i for i in range(n_unscaled)]
    for li, name in enumerate(names):
        p = params[name]
        out.append(_arr(p["Dense_0"], "kernel", (d, HIDDEN)))
        out.append(_arr(p["Dense_0"], "bias", (HIDDEN,)))
        out.append(_arr(p["Dense_1"], "kernel", (HIDDEN, u)))
        out.append(_arr(p["Dense_1"], "bias", (u,)))
        if li < n_scaled:
            out.append(_arr(p["Dense_2"], "kernel", (HIDDEN, u)))
            out.append(_arr(p["Dense_2"], "bias", (u,)))
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)


def flatten_rqspline(params, ndim, n_layers, n_bins, hidden_size, pre_offset=None, pre_amp=None):
    npar = 3 * n_bins + 1
    widths = [ndim, ndim] + list(hidden_size)
    out = _pre(ndim, pre_offset, pre_amp)
    for i in range(n_layers):
        cond = params["conditioner_%d" % i]["conditioner"]
        mlp = cond["layers_0"]
        for m in range(len(widths) - 1):
            out.append(_arr(mlp["layers_%d" % m], "kernel", (widths[m], widths[m + 1])))
            out.append(_arr(mlp["layers_%d" % m], "bias", (widths[m + 1],)))
        out.append(_arr(cond["layers_1"], "kernel", (widths[-1], ndim * npar)))
        out.append(_arr(cond["layers_1"], "bias", (ndim * npar,)))
        sc = params["scalar_%d" % i]
        out.append(_arr(sc, "scales", (ndim,)))
        out.append(_arr(sc, "shifts", (ndim,)))
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)


def random_realnvp_params(ndim, n_scaled=2, n_unscaled=4, seed=0, weight_scale=1.0):
    """Synthetic, non-trivial RealNVP weights with flax names (lecun-normal-like magnitudes)."""
    rng = np.random.default_rng(seed)
    d = num_masked(ndim)
    u = ndim - d

    def dense(n_in, n_out, s):
        return {"kernel": (rng.standard_normal((n_in, n_out)) * s / np.sqrt(n_in)).astype(np.float32),
                "bias": (rng.standard_normal(n_out) * 0.1).astype(np.float32)}

    params = {}
    for i in range(n_scaled):
        params["scaled_layers_%d" % i] = {
            "Dense_0": dense(d, HIDDEN, weight_scale), "Dense_1": dense(HIDDEN, u, 0.5 * weight_scale),
            "Dense_2": dense(HIDDEN, u, 0.5 * weight_scale)}
    for i in range(n_unscaled):
        params["unscaled_layers_%d" % i] = {
            "Dense_0": dense(d, HIDDEN, weight_scale), "Dense_1": dense(HIDDEN, u, 0.5 * weight_scale)}
    return params


def random_rqspline_params(ndim, n_layers=8, n_bins=8, hidden_size=(64, 64), seed=0, weight_scale=1.0):
    """Synthetic RQSpline weights with flax names; the output layer is non-zero so the splines
    are not the identity (the reference initialises it to zero, flows.py:361-365)."""
    rng = np.random.default_rng(seed)
    npar = 3 * n_bins + 1
    widths = [ndim, ndim] + list(hidden_size)

    def dense(n_in, n_out, s):
        return {"kernel": (rng.standard_normal((n_in, n_out)) * s / np.sqrt(n_in)).astype(np.float32),
                "bias": (rng.standard_normal(n_out) * 0.1).astype(np.float32)}

    params = {}
    for i in range(n_layers):
        mlp = {"layers_%d" % m: dense(widths[m], widths[m + 1], weight_scale) for m in range(len(widths) - 1)}
        params["conditioner_%d" % i] = {"conditioner": {
            "layers_0": mlp, "layers_1": dense(widths[-1], ndim * npar, 0.7 * weight_scale)}}
        params["scalar_%d" % i] = {
            "scales": (1.0 + 0.2 * rng.standard_normal(ndim)).astype(np.float32),
            "shifts": (0.1 * rng.standard_normal(ndim)).astype(np.float32)}
    return params
