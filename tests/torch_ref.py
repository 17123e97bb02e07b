"""float64 torch restatement of oracle/flows.py::realnvp_log_prob for LARGE-size GPU tests (test
infrastructure only; checked against the numpy oracle in test_gpu_fullsize.py)."""
import math

import numpy as np
import torch


def realnvp_log_prob_torch(model, x, chunk=1 << 20):
    """model: harmonic_b200 RealNVPModel with loaded params; x: (N, D) CUDA tensor -> float64 (N,)."""
    dev = x.device
    D = model.ndim
    d = int(np.round(D * 0.5))
    T = float(model.temperature)
    names = ["scaled_layers_%d" % i for i in range(model.n_scaled_layers)] + [
        "unscaled_layers_%d" % i for i in range(model.n_unscaled_layers)]
    P = model.state.params
    W = [{k: {kk: torch.from_numpy(np.asarray(vv, dtype=np.float64)).to(dev) for kk, vv in v.items()}
          for k, v in P[n].items()} for n in names]
    off = amp = None
    if model.standardize:
        off = torch.from_numpy(np.asarray(model.pre_offset, dtype=np.float64)).to(dev)
        amp = torch.from_numpy(np.asarray(model.pre_amp, dtype=np.float64)).to(dev)
    out = torch.empty(x.shape[0], dtype=torch.float64, device=dev)
    for a in range(0, x.shape[0], chunk):
        y = x[a:a + chunk].to(torch.float64)
        if off is not None:
            y = (y - off) / amp
        ldj = torch.zeros(y.shape[0], dtype=torch.float64, device=dev)
        for li, p in enumerate(W):
            if li > 0:
                y = torch.roll(y, -1, dims=1)
            y0 = y[:, :d]
            h = y0 @ p["Dense_0"]["kernel"] + p["Dense_0"]["bias"]
            h = torch.where(h >= 0, h, 0.01 * h)
            y1 = y[:, d:] - (h @ p["Dense_1"]["kernel"] + p["Dense_1"]["bias"])
            if li < model.n_scaled_layers:
                s = torch.nn.functional.softplus(h @ p["Dense_2"]["kernel"] + p["Dense_2"]["bias"]).clamp(1e-3, 1e3)
                y1 = y1 / s
                ldj = ldj - torch.log(s).sum(dim=1)
            y = torch.cat([y0, y1], dim=1)
        lp = -0.5 * ((y / T) ** 2).sum(dim=1) - D * math.log(T) - 0.5 * D * math.log(2 * math.pi) + ldj
        if amp is not None:
            lp = lp - torch.log(amp).sum()
        out[a:a + chunk] = lp
    return out


def evidence_from_lnarg(lnarg64, start):
    """float64 device reference of ln rho_hat and the per-chain logsumexp."""
    N = lnarg64.numel()
    fin = torch.isfinite(lnarg64)
    la = torch.where(fin, lnarg64, torch.full_like(lnarg64, -float("inf")))
    total = torch.logsumexp(la, dim=0) - math.log(N)
    return float(total), int((~fin).sum())
