"""Train the cfg1 flow fixture (BASELINE.json configs[0]: examples/gaussian_nondiagcov.py, ndim 5,
RealNVP 2 scaled + 4 unscaled layers, standardize=True) with a small torch (CPU) stand-in for the
reference's JAX training loop (harmonic/model.py:11-94,129-194; jax/flax are not installed here), and
save the weights in flax naming as tests/golden/cfg1_realnvp_flow.npz.

The torch model computes exactly oracle/flows.py::realnvp_log_prob at T = 1 (the training objective
of the reference is -mean(log_prob(x, T=1)), model.py:25-40).  Run: python tests/golden/make_cfg1_flow.py
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import bench_workloads as wl  # noqa: E402

NDIM, NS, NU, HID = 5, 2, 4, 128


class Coupling(torch.nn.Module):
    def __init__(self, d, u, scaled):
        super().__init__()
        self.l0 = torch.nn.Linear(d, HID)
        self.l1 = torch.nn.Linear(HID, u)
        self.l2 = torch.nn.Linear(HID, u) if scaled else None


class Flow(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.d = int(np.round(NDIM * 0.5))
        self.u = NDIM - self.d
        self.layers = torch.nn.ModuleList([Coupling(self.d, self.u, i < NS) for i in range(NS + NU)])

    def log_prob(self, x, T=1.0):
        y = x
        ldj = torch.zeros(x.shape[0], dtype=x.dtype)
        for li, L in enumerate(self.layers):
            if li > 0:
                y = torch.roll(y, -1, dims=1)
            y0, y1 = y[:, :self.d], y[:, self.d:]
            h = torch.nn.functional.leaky_relu(L.l0(y0), 0.01)
            y1 = y1 - L.l1(h)
            if L.l2 is not None:
                s = torch.nn.functional.softplus(L.l2(h)).clamp(1e-3, 1e3)
                y1 = y1 / s
                ldj = ldj - torch.log(s).sum(dim=1)
            y = torch.cat([y0, y1], dim=1)
        return -0.5 * ((y / T) ** 2).sum(dim=1) - NDIM * math.log(T) - 0.5 * NDIM * math.log(2 * math.pi) + ldj


def main():
    torch.manual_seed(0)
    x, _ = wl.gaussian_host(NDIM, 60000, seed=123)  # training half (the example trains on half the chains)
    x = torch.from_numpy(x.astype(np.float64))
    off = x.mean(dim=0)
    amp = torch.sqrt(torch.diag(torch.cov(x.T)))  # harmonic/model.py:176-181
    xs = ((x - off) / amp).float()
    flow = Flow()
    opt = torch.optim.Adam(flow.parameters(), lr=1e-3, betas=(0.9, 0.999))
    for epoch in range(12):
        perm = torch.randperm(xs.shape[0])
        tot = 0.0
        for i in range(0, xs.shape[0], 256):
            b = xs[perm[i:i + 256]]
            loss = -flow.log_prob(b).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            tot += float(loss) * b.shape[0]
        print("epoch", epoch, "loss", tot / xs.shape[0])
    out = {"pre_offset": off.numpy().astype(np.float32), "pre_amp": amp.numpy().astype(np.float32)}
    for i, L in enumerate(flow.layers):
        name = ("scaled_layers_%d" % i) if i < NS else ("unscaled_layers_%d" % (i - NS))
        for k, lin in (("Dense_0", L.l0), ("Dense_1", L.l1), ("Dense_2", L.l2)):
            if lin is None:
                continue
            out["%s/%s/kernel" % (name, k)] = lin.weight.detach().numpy().T.astype(np.float32).copy()  # flax: (in, out)
            out["%s/%s/bias" % (name, k)] = lin.bias.detach().numpy().astype(np.float32).copy()
    np.savez_compressed(os.path.join(HERE, "cfg1_realnvp_flow.npz"), **out)
    print("saved", len(out), "arrays")


if __name__ == "__main__":
    main()
