"""Numpy restatement of the flow half of the harmonic hot path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: tfp-jax / distrax / flax are absent from this image, so this file
restates their published arithmetic; see ``oracle/__init__.py``.

What is restated and where it comes from (paths into /root/reference):

* ``predict``            -> ``harmonic/model.py:196-237`` (temperature guards, optional
  standardisation, ``- sum(log(pre_amp))``).
* ``realnvp_log_prob``   -> ``harmonic/flows.py:41-99`` (chain wiring, permutation,
  base ``MultivariateNormalDiag(scale_diag=T)``), ``:159-180`` (AffineCoupling MLP),
  plus tfp-jax ``RealNVP(fraction_masked=0.5)``, ``Permute``, ``Chain``, ``Shift``,
  ``Scale`` and ``TransformedDistribution.log_prob`` (inverse pass + inverse log-det).
* ``sample_transform``   -> ``harmonic/model.py:239-275``, ``harmonic/flows.py:118-138,290-313`` (the
  sampling direction: tfb.Chain.forward / distrax Inverse(Chain).forward, RQ spline inverse).
* ``rqspline_log_prob``  -> ``harmonic/flows.py:234-288`` (layer wiring, masks, base
  ``MultivariateNormalFullCovariance(cov=T*I)``), ``:335-438`` (Conditioner / MLP /
  Reshape / Scalar), plus distrax ``RationalQuadraticSpline`` (forward),
  ``MaskedCoupling``, ``ScalarAffine``, ``Chain``, ``Inverse``, ``Transformed``.

Parameters are flax-style nested dicts exactly as ``model.state.params`` holds them
(names in SURVEY.md Appendix B).  ``dtype`` selects the arithmetic type: float64 is the
oracle proper, float32 emulates the reference's JAX default (x64 disabled).
"""
import math

import numpy as np

# --- named "ordering facts" (hard part 1): one flip fixes all ---
PERMUTE_INVERSE_IS_ROTATE_LEFT = True  # tfb.Permute([D-1,0..D-2]).inverse == roll(-1)
MASK_TRUE_IS_PASSTHROUGH = True  # distrax.MaskedCoupling: mask True -> unchanged
RESHAPE_ROW_MAJOR = True  # Reshape((D, 3K+1)) of a 1-D vector is row-major
HIDDEN = 128  # nn.Dense(128) in AffineCoupling (flows.py:171)
LEAKY_SLOPE = 0.01  # flax nn.leaky_relu default
MIN_BIN_SIZE = 1e-4  # distrax RationalQuadraticSpline default
MIN_KNOT_SLOPE = 1e-4  # distrax RationalQuadraticSpline default


def num_masked(ndim):
    """tfp RealNVP: int(np.round(input_depth * fraction_masked)), fraction 0.5."""
    return int(np.round(ndim * 0.5))


def _softplus(a):
    # jax.nn.softplus == logaddexp(a, 0)
    return np.logaddexp(a, np.zeros((), dtype=a.dtype))


def _dense(a, layer, dtype):
    return a @ np.asarray(layer["kernel"], dtype=dtype) + np.asarray(
        layer["bias"], dtype=dtype
    )


def realnvp_log_prob(params, x, temperature, n_scaled, n_unscaled, dtype=np.float64):
    """log_prob of harmonic.flows.RealNVP for a batch x (N, D).  flows.py:41-116,140-156."""
    x = np.asarray(x, dtype=dtype)
    n, D = x.shape
    d = num_masked(D)
    T = dtype(temperature)
    y = x.copy()
    ldj = np.zeros(n, dtype=dtype)

    layers = [("scaled_layers_%d" % i, True) for i in range(n_scaled)] + [
        ("unscaled_layers_%d" % i, False) for i in range(n_unscaled)
    ]
    with np.errstate(all="ignore"):
        for li, (name, scaled) in enumerate(layers):
            if li > 0:
                # tfb.Permute(perm).inverse, perm = [D-1, 0, ..., D-2] (flows.py:65-66)
                y = np.roll(y, -1 if PERMUTE_INVERSE_IS_ROTATE_LEFT else 1, axis=1)
            p = params[name]
            y0 = y[:, :d]
            h = _dense(y0, p["Dense_0"], dtype)
            h = np.where(h >= 0, h, dtype(LEAKY_SLOPE) * h)  # flows.py:171
            shift = _dense(h, p["Dense_1"], dtype)  # flows.py:174
            y1 = y[:, d:] - shift
            if scaled:
                scale = np.clip(
                    _softplus(_dense(h, p["Dense_2"], dtype)), dtype(1e-3), dtype(1e3)
                )  # flows.py:177
                y1 = y1 / scale
                ldj = ldj - np.sum(np.log(scale), axis=1, dtype=dtype)
            y = np.concatenate([y0, y1], axis=1)
        # tfd.MultivariateNormalDiag(loc=0, scale_diag=T) (flows.py:91-95): std = T
        z = y / T
        lp = (
            dtype(-0.5) * np.sum(z * z, axis=1, dtype=dtype)
            - dtype(D) * dtype(math.log(temperature))
            - dtype(0.5 * D * math.log(2.0 * math.pi))
        )
    return (lp + ldj).astype(dtype)


def _mlp_conditioner(hin, cp, n_hidden, dtype):
    """Conditioner = Sequential[MLP([D]+hidden, tanh), Dense(D*(3K+1)), Reshape] (flows.py:353-368).

    MLP.__call__ applies the activation after every Dense but the last (flows.py:434-438).
    """
    mlp = cp["conditioner"]["layers_0"]
    a = hin
    nl = n_hidden + 1  # Dense(D) + one per hidden size
    for m in range(nl):
        a = _dense(a, mlp["layers_%d" % m], dtype)
        if m < nl - 1:
            a = np.tanh(a)
    return _dense(a, cp["conditioner"]["layers_1"], dtype)


def _rqs_forward(x, P, K, lo, hi, dtype):
    """distrax.RationalQuadraticSpline(P, lo, hi).forward_and_log_det for x (N,), P (N, 3K+1)."""
    R = dtype(hi - lo)
    mbs = dtype(MIN_BIN_SIZE)
    mks = dtype(MIN_KNOT_SLOPE)

    def softmax(u):
        u = u - np.max(u, axis=-1, keepdims=True)
        e = np.exp(u)
        return e / np.sum(e, axis=-1, keepdims=True)

    w = softmax(P[:, 0:K]) * (R - dtype(K) * mbs) + mbs
    hgt = softmax(P[:, K : 2 * K]) * (R - dtype(K) * mbs) + mbs
    n = x.shape[0]
    lo_col = np.full((n, 1), lo, dtype=dtype)
    hi_col = np.full((n, 1), hi, dtype=dtype)
    xk = np.concatenate(
        [lo_col, dtype(lo) + np.cumsum(w[:, :-1], axis=1, dtype=dtype), hi_col], axis=1
    )
    yk = np.concatenate(
        [lo_col, dtype(lo) + np.cumsum(hgt[:, :-1], axis=1, dtype=dtype), hi_col], axis=1
    )
    offset = dtype(math.log(math.exp(1.0 - MIN_KNOT_SLOPE) - 1.0))
    s = _softplus(P[:, 2 * K :] + offset) + mks  # K+1 slopes, 'unconstrained'

    below = x <= xk[:, 0]
    above = x >= xk[:, -1]
    correct = (x[:, None] >= xk[:, :-1]) & (x[:, None] < xk[:, 1:])
    anyb = np.any(correct, axis=1)
    first = np.zeros_like(correct)
    first[:, 0] = True
    correct = np.where(anyb[:, None], correct, first)
    k = np.argmax(correct, axis=1)
    rows = np.arange(n)
    xl, xr = xk[rows, k], xk[rows, k + 1]
    yl, yr = yk[rows, k], yk[rows, k + 1]
    sl, sr = s[rows, k], s[rows, k + 1]
    bw = xr - xl
    bh = yr - yl
    delta = bh / bw
    z = np.clip((x - xl) / bw, dtype(0), dtype(1))
    sq_z = z * z
    z1mz = z - sq_z
    sq_1mz = (dtype(1) - z) ** 2
    slopes_term = sr + sl - dtype(2) * delta
    numerator = bh * (delta * sq_z + sl * z1mz)
    denominator = delta + slopes_term * z1mz
    y = yl + numerator / denominator
    logdet = (
        dtype(2) * np.log(delta)
        + np.log(sr * sq_z + dtype(2) * delta * z1mz + sl * sq_1mz)
        - dtype(2) * np.log(denominator)
    )
    y = np.where(below, (x - xk[:, 0]) * s[:, 0] + yk[:, 0], y)
    y = np.where(above, (x - xk[:, -1]) * s[:, -1] + yk[:, -1], y)
    logdet = np.where(below, np.log(s[:, 0]), logdet)
    logdet = np.where(above, np.log(s[:, -1]), logdet)
    return y.astype(dtype), logdet.astype(dtype)


def rqspline_log_prob(
    params, x, temperature, n_layers, n_bins, hidden_size, spline_range, dtype=np.float64
):
    """log_prob of harmonic.flows.RQSpline for a batch x (N, D).  flows.py:234-288,315-332."""
    x = np.asarray(x, dtype=dtype)
    n, D = x.shape
    K = n_bins
    npar = 3 * K + 1
    lo, hi = float(spline_range[0]), float(spline_range[1])
    y = x.copy()
    ld = np.zeros(n, dtype=dtype)
    mask0 = (np.arange(D) % 2).astype(bool)  # flows.py:245
    with np.errstate(all="ignore"):
        # Inverse(Chain(layers)).inverse == Chain.forward: LAST list element first.
        for i in range(n_layers - 1, -1, -1):
            mask = mask0 if i % 2 == 0 else ~mask0  # negated once per layer (flows.py:261)
            if not MASK_TRUE_IS_PASSTHROUGH:
                mask = ~mask
            # (1) spline coupling i, forward
            hin = np.where(mask[None, :], y, dtype(0))
            P = _mlp_conditioner(hin, params["conditioner_%d" % i], len(hidden_size), dtype)
            if RESHAPE_ROW_MAJOR:
                P = P.reshape(n, D, npar)
            else:
                P = P.reshape(n, npar, D).transpose(0, 2, 1)
            for j in range(D):
                if mask[j]:
                    continue
                yj, lj = _rqs_forward(y[:, j], P[:, j, :], K, lo, hi, dtype)
                y[:, j] = yj
                ld = ld + lj
            # (2) scalar affine i, forward, all dims (mask_all = False)
            sc = np.asarray(params["scalar_%d" % i]["scales"], dtype=dtype)
            sh = np.asarray(params["scalar_%d" % i]["shifts"], dtype=dtype)
            y = y * sc[None, :] + sh[None, :]
            ld = ld + np.sum(np.log(np.abs(sc)), dtype=dtype)
        # MultivariateNormalFullCovariance(0, T*I): VARIANCE = T (flows.py:264-269)
        T = dtype(temperature)
        lp = (
            dtype(-0.5) * np.sum(y * y, axis=1, dtype=dtype) / T
            - dtype(0.5 * D * math.log(2.0 * math.pi))
            - dtype(0.5 * D) * dtype(math.log(temperature))
        )
    return (lp + ld).astype(dtype)


def realnvp_sample_transform(params, z, n_scaled, n_unscaled, dtype=np.float64):
    """Sampling direction of harmonic.flows.RealNVP (flows.py:118-138): x = bijector.forward(z) for base
    draws z (N, D) ~ N(0, T^2 I).  tfb.Chain.forward applies the list [S0, P, S1, ..] right to left."""
    y = np.asarray(z, dtype=dtype).copy()
    n, D = y.shape
    d = num_masked(D)
    layers = [("scaled_layers_%d" % i, True) for i in range(n_scaled)] + [
        ("unscaled_layers_%d" % i, False) for i in range(n_unscaled)
    ]
    with np.errstate(all="ignore"):
        for li in range(len(layers) - 1, -1, -1):
            name, scaled = layers[li]
            p = params[name]
            y0 = y[:, :d]
            h = _dense(y0, p["Dense_0"], dtype)
            h = np.where(h >= 0, h, dtype(LEAKY_SLOPE) * h)
            shift = _dense(h, p["Dense_1"], dtype)
            y1 = y[:, d:]
            if scaled:
                scale = np.clip(_softplus(_dense(h, p["Dense_2"], dtype)), dtype(1e-3), dtype(1e3))
                y1 = y1 * scale
            y = np.concatenate([y0, y1 + shift], axis=1)  # Chain([Shift, Scale]).forward = scale * x + shift
            if li > 0:
                # tfb.Permute(perm).forward = gather with perm = [D-1, 0, ..., D-2] (flows.py:65-66)
                y = np.roll(y, 1 if PERMUTE_INVERSE_IS_ROTATE_LEFT else -1, axis=1)
    return y.astype(dtype)


def _rqs_inverse(y, P, K, lo, hi, dtype):
    """distrax.RationalQuadraticSpline(P, lo, hi).inverse for y (N,), P (N, 3K+1) (quadratic root)."""
    R = dtype(hi - lo)
    mbs = dtype(MIN_BIN_SIZE)
    mks = dtype(MIN_KNOT_SLOPE)

    def softmax(u):
        u = u - np.max(u, axis=-1, keepdims=True)
        e = np.exp(u)
        return e / np.sum(e, axis=-1, keepdims=True)

    w = softmax(P[:, 0:K]) * (R - dtype(K) * mbs) + mbs
    hgt = softmax(P[:, K : 2 * K]) * (R - dtype(K) * mbs) + mbs
    n = y.shape[0]
    lo_col = np.full((n, 1), lo, dtype=dtype)
    hi_col = np.full((n, 1), hi, dtype=dtype)
    xk = np.concatenate([lo_col, dtype(lo) + np.cumsum(w[:, :-1], axis=1, dtype=dtype), hi_col], axis=1)
    yk = np.concatenate([lo_col, dtype(lo) + np.cumsum(hgt[:, :-1], axis=1, dtype=dtype), hi_col], axis=1)
    offset = dtype(math.log(math.exp(1.0 - MIN_KNOT_SLOPE) - 1.0))
    s = _softplus(P[:, 2 * K :] + offset) + mks
    below = y <= yk[:, 0]
    above = y >= yk[:, -1]
    correct = (y[:, None] >= yk[:, :-1]) & (y[:, None] < yk[:, 1:])
    anyb = np.any(correct, axis=1)
    first = np.zeros_like(correct)
    first[:, 0] = True
    correct = np.where(anyb[:, None], correct, first)
    k = np.argmax(correct, axis=1)
    rows = np.arange(n)
    xl, xr = xk[rows, k], xk[rows, k + 1]
    yl, yr = yk[rows, k], yk[rows, k + 1]
    sl, sr = s[rows, k], s[rows, k + 1]
    bw = xr - xl
    bh = yr - yl
    delta = bh / bw
    wv = np.clip((y - yl) / bh, dtype(0), dtype(1))
    slopes_term = sr + sl - dtype(2) * delta
    c = -delta * wv
    b = sl - slopes_term * wv
    a = delta - b
    z = -dtype(2) * c / (b + np.sqrt(b * b - dtype(4) * a * c))
    z = np.clip(z, dtype(0), dtype(1))
    x = bw * z + xl
    x = np.where(below, (y - yk[:, 0]) / s[:, 0] + xk[:, 0], x)
    x = np.where(above, (y - yk[:, -1]) / s[:, -1] + xk[:, -1], x)
    return x.astype(dtype)


def rqspline_sample_transform(params, z, n_layers, n_bins, hidden_size, spline_range, dtype=np.float64):
    """Sampling direction of harmonic.flows.RQSpline (flows.py:290-313): Inverse(Chain(layers)).forward =
    Chain.inverse, i.e. the inverses of [affine_0, spline_0, affine_1, ..] in LIST order, for base draws
    z (N, D) ~ N(0, T I)."""
    y = np.asarray(z, dtype=dtype).copy()
    n, D = y.shape
    K = n_bins
    npar = 3 * K + 1
    lo, hi = float(spline_range[0]), float(spline_range[1])
    mask0 = (np.arange(D) % 2).astype(bool)
    with np.errstate(all="ignore"):
        for i in range(n_layers):
            mask = mask0 if i % 2 == 0 else ~mask0
            if not MASK_TRUE_IS_PASSTHROUGH:
                mask = ~mask
            sc = np.asarray(params["scalar_%d" % i]["scales"], dtype=dtype)
            sh = np.asarray(params["scalar_%d" % i]["shifts"], dtype=dtype)
            y = (y - sh[None, :]) / sc[None, :]
            hin = np.where(mask[None, :], y, dtype(0))
            P = _mlp_conditioner(hin, params["conditioner_%d" % i], len(hidden_size), dtype)
            if RESHAPE_ROW_MAJOR:
                P = P.reshape(n, D, npar)
            else:
                P = P.reshape(n, npar, D).transpose(0, 2, 1)
            for j in range(D):
                if mask[j]:
                    continue
                y[:, j] = _rqs_inverse(y[:, j], P[:, j, :], K, lo, hi, dtype)
    return y.astype(dtype)


def sample_transform(spec, eps, dtype=np.float64):
    """FlowModel.sample (model.py:239-275) with the standard-normal draws ``eps`` (N, D) supplied by the
    caller: base draw = T * eps (RealNVP, scale_diag = T) or sqrt(T) * eps (RQSpline, covariance T I),
    flow in the sampling direction, then ``x * pre_amp + pre_offset`` if standardised."""
    T = spec["temperature"]
    if T > 1:
        raise ValueError("Scaling must not be greater than 1.")
    if T <= 0:
        raise ValueError("Scaling must be positive.")
    eps = np.asarray(eps, dtype=dtype)
    if spec["kind"] == "realnvp":
        x = realnvp_sample_transform(spec["params"], eps * dtype(T), spec["n_scaled_layers"],
                                     spec["n_unscaled_layers"], dtype)
    elif spec["kind"] == "rqspline":
        x = rqspline_sample_transform(spec["params"], eps * dtype(math.sqrt(T)), spec["n_layers"], spec["n_bins"],
                                      spec["hidden_size"], spec["spline_range"], dtype)
    else:
        raise ValueError("unknown flow kind")
    if spec.get("standardize", False):
        x = x * np.asarray(spec["pre_amp"], dtype=dtype) + np.asarray(spec["pre_offset"], dtype=dtype)
    return x


def predict(spec, x, dtype=np.float64):
    """FlowModel.predict (model.py:196-237) for a model ``spec`` dict:

    {"kind": "realnvp"|"rqspline", "params": pytree, "temperature": T, "standardize": bool,
     "pre_offset": (D,), "pre_amp": (D,), + hyper-parameters of the kind}
    """
    T = spec["temperature"]
    if T > 1:
        raise ValueError("Scaling must not be greater than 1.")
    if T <= 0:
        raise ValueError("Scaling must be positive.")
    x = np.asarray(x, dtype=dtype)
    one_d = x.ndim == 1
    if one_d:
        x = x[None, :]
    if spec.get("standardize", False):
        off = np.asarray(spec["pre_offset"], dtype=dtype)
        amp = np.asarray(spec["pre_amp"], dtype=dtype)
        x = (x - off) / amp
    if spec["kind"] == "realnvp":
        lp = realnvp_log_prob(
            spec["params"], x, T, spec["n_scaled_layers"], spec["n_unscaled_layers"], dtype
        )
    elif spec["kind"] == "rqspline":
        lp = rqspline_log_prob(
            spec["params"],
            x,
            T,
            spec["n_layers"],
            spec["n_bins"],
            spec["hidden_size"],
            spec["spline_range"],
            dtype,
        )
    else:
        raise ValueError("unknown flow kind")
    if spec.get("standardize", False):
        lp = lp - np.sum(np.log(amp), dtype=dtype)
    return lp[0] if one_d else lp
