#!/usr/bin/env python
"""Capture per-sample golden vectors from the REAL reference (run on any machine with
`pip install harmonic` + jax/flax/distrax/tfp; none of these exist in the build image).

For each flow family this trains a small reference model, then dumps
    {hyper-parameters, flattened model.state.params, pre_offset, pre_amp, temperature, x, predict(x)}
to tests/golden/reference_<name>.npz.  tests/test_reference_golden.py picks the files up
automatically: the numpy oracle (oracle/flows.py) and, on a B200, the CUDA kernels are then checked
against the reference's own per-sample ln phi, which pins the six "ordering facts" of SURVEY.md
section 7 (permute direction, num_masked rounding, chain order, mask polarity, reshape order, flax
parameter names).  This closes the "parity unpinned" gap of the flow half of the path.

usage: python tools/capture_reference_golden.py [outdir]
"""
import os
import sys

import numpy as np


def flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if hasattr(v, "items"):
            out.update(flatten(v, prefix + k + "/"))
        else:
            out[prefix + k] = np.asarray(v)
    return out


def main():
    import harmonic as hm  # the reference
    import jax

    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    rng = np.random.default_rng(0)
    cases = [
        ("realnvp_d5", lambda: hm.model.RealNVPModel(5, standardize=True, temperature=0.8), 5),
        ("realnvp_d2_s3u1", lambda: hm.model.RealNVPModel(2, n_scaled_layers=3, n_unscaled_layers=1, temperature=0.6), 2),
        ("realnvp_d64", lambda: hm.model.RealNVPModel(64, standardize=True, temperature=0.8), 64),
        ("rqspline_d2", lambda: hm.model.RQSplineModel(2, temperature=0.9), 2),
        ("rqspline_d5_small", lambda: hm.model.RQSplineModel(5, n_layers=3, n_bins=8, hidden_size=[32, 32],
                                                             standardize=True, temperature=0.8), 5),
    ]
    for name, make, ndim in cases:
        model = make()
        cov = np.eye(ndim) + 0.3 * rng.standard_normal((ndim, ndim))
        cov = cov @ cov.T
        X = rng.multivariate_normal(np.zeros(ndim), cov, size=4000)
        model.fit(X, epochs=3, key=jax.random.PRNGKey(1))
        x = rng.multivariate_normal(np.zeros(ndim), cov, size=512) * 1.5
        x[:8] *= 8.0  # tails of the spline range
        lnphi = np.asarray(model.predict(x))
        one = np.asarray(model.predict(x[0]))
        arrays = {"params/" + k: v for k, v in flatten(model.state.params).items()}
        arrays.update(x=x, lnphi=lnphi, lnphi_one=one, temperature=np.float64(model.temperature),
                      standardize=np.bool_(model.standardize), ndim=np.int64(ndim), kind=np.str_(type(model).__name__))
        if model.standardize:
            arrays.update(pre_offset=np.asarray(model.pre_offset), pre_amp=np.asarray(model.pre_amp))
        for attr in ("n_scaled_layers", "n_unscaled_layers", "n_layers", "n_bins", "hidden_size", "spline_range"):
            if hasattr(model, attr):
                arrays["hp/" + attr] = np.asarray(getattr(model, attr))
        np.savez_compressed(os.path.join(out, "reference_%s.npz" % name), **arrays)
        print("wrote", name, lnphi[:3])


if __name__ == "__main__":
    main()
