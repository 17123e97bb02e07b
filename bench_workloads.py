"""Synthetic workloads for bench.py and the large-size tests.

cfg4 (north-star): D=64 Gaussian with the tridiagonal covariance pattern of
examples/gaussian_nondiagcov.py:49-70 (np.random.seed(10)), samples x = L eps drawn on the device,
lnP = -1/2 x^T Sigma^-1 x = -1/2 |eps|^2 (examples/gaussian_nondiagcov.py:31-46), RealNVP 2 scaled +
4 unscaled coupling layers (harmonic/model.py:289-290), T = 0.8, 1000 chains x 1e5 samples per GPU.
cfg2: Rosenbrock D=2 (examples/rosenbrock.py:11-34), RQSpline defaults, T = 0.9.
cfg3: D=5 RealNVP 6+2 (examples/pima_indian.py:179-182), T = 0.9.
cfg5: accumulation only, precomputed ln phi / ln posterior.
Weights are seeded synthetic (no JAX in the image to train with).
"""
import os

import numpy as np

import harmonic_b200 as hm
from harmonic_b200 import export

ROOT = os.path.dirname(os.path.abspath(__file__))


def make_realnvp(ndim, n_scaled=2, n_unscaled=4, seed=0, standardize=False, temperature=0.8, ws=1.0):
    """RealNVPModel with seeded synthetic weights (flax naming) installed through load_params."""
    m = hm.model.RealNVPModel(ndim, n_scaled, n_unscaled, standardize=standardize, temperature=temperature)
    rng = np.random.default_rng(seed + 1000)
    off = rng.standard_normal(ndim).astype(np.float32) * 0.3 if standardize else None
    amp = (0.5 + rng.random(ndim)).astype(np.float32) if standardize else None
    m.load_params(export.random_realnvp_params(ndim, n_scaled, n_unscaled, seed=seed, weight_scale=ws), off, amp)
    return m


def make_rqspline(ndim, n_layers=3, n_bins=8, hidden=(32, 32), seed=0, standardize=False,
                  temperature=0.9, spline_range=(-10.0, 10.0), ws=1.0):
    m = hm.model.RQSplineModel(ndim, n_layers, n_bins, list(hidden), spline_range,
                               standardize=standardize, temperature=temperature)
    rng = np.random.default_rng(seed + 2000)
    off = rng.standard_normal(ndim).astype(np.float32) * 0.3 if standardize else None
    amp = (0.5 + rng.random(ndim)).astype(np.float32) if standardize else None
    m.load_params(export.random_rqspline_params(ndim, n_layers, n_bins, hidden, seed=seed, weight_scale=ws), off, amp)
    return m


def load_cfg1_model(temperature=0.8):
    """The trained cfg1 fixture (tests/golden/make_cfg1_flow.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_realnvp_flow.npz"))
    params = {}
    for k in z.files:
        if "/" not in k:
            continue
        layer, dense, leaf = k.split("/")
        params.setdefault(layer, {}).setdefault(dense, {})[leaf] = z[k]
    m = hm.model.RealNVPModel(5, 2, 4, standardize=True, temperature=temperature)
    m.load_params(params, z["pre_offset"], z["pre_amp"])
    return m

FLOP_PER_SAMPLE = {"cfg4": 114688.0, "cfg3": 14848.0, "cfg1": 9216.0, "cfg2": 118848.0, "cfg2x": 16152.0}
BYTES_PER_SAMPLE = {"cfg4": 260.0, "cfg3": 24.0, "cfg1": 24.0, "cfg2": 12.0, "cfg2x": 12.0, "cfg5": 8.0}


def init_cov(ndim, seed=10):
    rs = np.random.RandomState(seed)
    cov = np.zeros((ndim, ndim))
    np.fill_diagonal(cov, np.ones(ndim) + rs.randn(ndim) * 0.1)
    for i in range(ndim - 1):
        cov[i, i + 1] = (-1) ** i * 0.5 * np.sqrt(cov[i, i] * cov[i + 1, i + 1])
        cov[i + 1, i] = cov[i, i + 1]
    return cov


def make_model(name):
    if name == "cfg4":
        return make_realnvp(64, 2, 4, seed=4, standardize=True, temperature=0.8, ws=0.3)
    if name == "cfg3":
        return make_realnvp(5, 6, 2, seed=3, standardize=True, temperature=0.9, ws=0.3)
    if name == "cfg1":
        return load_cfg1_model(0.8)
    if name == "cfg2":
        return make_rqspline(2, 8, 8, (64, 64), seed=2, standardize=False, temperature=0.9, ws=0.15)
    if name == "cfg2x":  # the example's architecture (examples/rosenbrock.py:175-178): 3 layers, hidden [32, 32]
        return make_rqspline(2, 3, 8, (32, 32), seed=2, standardize=False, temperature=0.9, ws=0.15)
    raise ValueError(name)


def default_shape(name):
    """(nchains, samples per chain) per GPU."""
    return {"cfg4": (1000, 100000), "cfg3": (1000, 10000), "cfg1": (100, 5000), "cfg2": (1000, 100000), "cfg2x": (1000, 100000),
            "cfg5": (10000, 100000)}[name]


def gaussian_host(ndim, n, seed):
    """Host samples of the cfg4/cfg3/cfg1 Gaussian target (float32) + ln posterior."""
    rng = np.random.default_rng(seed)
    L = np.linalg.cholesky(init_cov(ndim))
    eps = rng.standard_normal((n, ndim))
    x = (eps @ L.T).astype(np.float32)
    lnp = (-0.5 * np.sum(eps * eps, axis=1)).astype(np.float32)
    return x, lnp


def rosenbrock_host(n, seed):
    rng = np.random.default_rng(seed)
    x1 = np.clip(1.0 + 0.7 * rng.standard_normal(n), -10, 10)
    x2 = np.clip(x1 * x1 + 0.07 * rng.standard_normal(n), -5, 15)
    x = np.stack([x1, x2], axis=1).astype(np.float32)
    lnp = (-(100.0 * (x2 - x1 * x1) ** 2 + (1.0 - x1) ** 2) + np.log(1.0 / 400.0)).astype(np.float32)
    return x, lnp


def host_sample(name, n, seed=0):
    if name in ("cfg2", "cfg2x"):
        return rosenbrock_host(n, seed)
    ndim = {"cfg4": 64, "cfg3": 5, "cfg1": 5}[name]
    return gaussian_host(ndim, n, seed)


def device_data(name, nchains, per_chain, device, seed=0):
    """Device-resident synthetic chains (float32), generated in blocks with torch's Philox."""
    import torch

    n = nchains * per_chain
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if name == "cfg5":
        lnphi = torch.empty(n, dtype=torch.float32, device=device)
        lnp = torch.empty(n, dtype=torch.float32, device=device)
        blk = 1 << 26
        for a in range(0, n, blk):
            b = min(n, a + blk)
            lnphi[a:b] = torch.randn(b - a, generator=g, device=device) - 20.0
            lnp[a:b] = lnphi[a:b] + 10.0 + 0.05 * torch.randn(b - a, generator=g, device=device)
            bad = torch.rand(b - a, generator=g, device=device) < 1e-4  # non-finite path
            lnphi[a:b][bad] = float("nan")
        return lnphi, lnp
    if name in ("cfg2", "cfg2x"):
        x = torch.empty((n, 2), dtype=torch.float32, device=device)
        lnp = torch.empty(n, dtype=torch.float32, device=device)
        blk = 1 << 24
        for a in range(0, n, blk):
            b = min(n, a + blk)
            x1 = (1.0 + 0.7 * torch.randn(b - a, generator=g, device=device)).clamp(-10, 10)
            x2 = (x1 * x1 + 0.07 * torch.randn(b - a, generator=g, device=device)).clamp(-5, 15)
            x[a:b, 0] = x1
            x[a:b, 1] = x2
            lnp[a:b] = -(100.0 * (x2 - x1 * x1) ** 2 + (1.0 - x1) ** 2) + float(np.log(1.0 / 400.0))
        return x, lnp
    ndim = {"cfg4": 64, "cfg3": 5, "cfg1": 5}[name]
    Lt = torch.from_numpy(np.linalg.cholesky(init_cov(ndim)).T.astype(np.float32)).to(device)
    x = torch.empty((n, ndim), dtype=torch.float32, device=device)
    lnp = torch.empty(n, dtype=torch.float32, device=device)
    blk = 1 << 22
    for a in range(0, n, blk):
        b = min(n, a + blk)
        eps = torch.randn((b - a, ndim), generator=g, device=device)
        torch.matmul(eps, Lt, out=x[a:b])
        lnp[a:b] = -0.5 * (eps * eps).sum(dim=1)
    return x, lnp
