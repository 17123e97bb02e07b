"""Per-sample goldens captured from the REAL reference (tools/capture_reference_golden.py, needs a
JAX machine).  None are committed yet -- jax / flax / distrax / tfp are not installable in the build
image -- so these tests SKIP until tests/golden/reference_*.npz exist; with the files present they pin
oracle/flows.py (CPU) and the CUDA kernels (GPU) to the reference's own ln phi at the north_star
tolerance (1e-5 relative)."""
import glob
import os

import numpy as np
import pytest

import harmonic_b200 as hm
from oracle import flows as of
from tests import util

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "reference_*.npz")))


def unflatten(z, prefix):
    tree = {}
    for k in z.files:
        if not k.startswith(prefix):
            continue
        node = tree
        parts = k[len(prefix):].split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = z[k]
    return tree


def build(z):
    kind = str(z["kind"])
    ndim = int(z["ndim"])
    std = bool(z["standardize"])
    if kind == "RealNVPModel":
        m = hm.model.RealNVPModel(ndim, int(z["hp/n_scaled_layers"]), int(z["hp/n_unscaled_layers"]),
                                  standardize=std, temperature=float(z["temperature"]))
    else:
        m = hm.model.RQSplineModel(ndim, int(z["hp/n_layers"]), int(z["hp/n_bins"]),
                                   [int(h) for h in z["hp/hidden_size"]], tuple(float(v) for v in z["hp/spline_range"]),
                                   standardize=std, temperature=float(z["temperature"]))
    m.load_params(unflatten(z, "params/"), z["pre_offset"] if std else None, z["pre_amp"] if std else None)
    return m


@pytest.mark.skipif(not FILES, reason="no reference goldens captured yet (tools/capture_reference_golden.py)")
@pytest.mark.parametrize("path", FILES or ["none"])
def test_oracle_matches_reference_goldens(path):
    z = np.load(path)
    m = build(z)
    got = of.predict(m.oracle_spec(), z["x"], np.float32)
    util.assert_lnphi_close(got, z["lnphi"], rel=2e-5)  # both sides float32
    assert float(of.predict(m.oracle_spec(), z["x"][0], np.float32)) == pytest.approx(float(z["lnphi_one"]), rel=2e-5)


@pytest.mark.gpu
@pytest.mark.skipif(not FILES, reason="no reference goldens captured yet (tools/capture_reference_golden.py)")
@pytest.mark.parametrize("path", FILES or ["none"])
def test_gpu_matches_reference_goldens(path):
    z = np.load(path)
    m = build(z)
    util.assert_lnphi_close(m.predict(z["x"]), z["lnphi"], rel=2e-5)
