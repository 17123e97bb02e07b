"""cfg1 (BASELINE.json configs[0], examples/gaussian_nondiagcov.py): ndim 5 Gaussian with a
non-diagonal covariance, 200 chains x 5000 samples, second half of the chains used for inference,
learned RealNVP flow at T = 0.8 -> ln Z against the ANALYTIC evidence.

The flow weights are a committed fixture trained by tests/golden/make_cfg1_flow.py (torch stand-in
for the reference's JAX training; jax is not installable here)."""
import numpy as np
import pytest

import bench_workloads as wl
from bench_workloads import load_cfg1_model
import harmonic_b200 as hm
from tests import util



def cfg1_chains():
    """examples/gaussian_nondiagcov.py:138-171 with direct sampling (MCMC=False branch)."""
    ndim, nchains, per = 5, 200, 5000
    x, lnp = wl.gaussian_host(ndim, nchains * per, seed=0)
    ch = hm.Chains(ndim)
    ch.add_chains_2d(x.astype(np.float64), lnp.astype(np.float64), nchains)
    infer = ch.get_sub_chains(list(range(100, 200)))  # second half of the chains (:169-171)
    cov = wl.init_cov(ndim)
    ln_z_analytic = 0.5 * ndim * np.log(2 * np.pi) + 0.5 * np.log(np.linalg.det(cov))  # :11-27
    return infer, ln_z_analytic


def test_cfg1_oracle_recovers_analytic_evidence():
    m = load_cfg1_model()
    ch, ln_z = cfg1_chains()
    ev = util.oracle_evidence(m, ch)
    est, _ = ev.ln_evidence()
    neg, pos = ev.ln_inv_evidence_errors()
    # ln(1/Z) +- errors brackets the analytic value within 4 sigma-equivalents
    assert abs(est - ln_z) < 4 * max(abs(neg), abs(pos)), (est, ln_z, neg, pos)
    assert abs(est - ln_z) < 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("temperature", [0.8, 0.9])
def test_cfg1_gpu_matches_oracle_and_analytic(temperature):
    m = load_cfg1_model(temperature)
    ch, ln_z = cfg1_chains()
    ev = hm.Evidence(ch.nchains, m)
    ev.add_chains(ch)
    ref = util.oracle_evidence(m, ch)
    util.assert_evidence_close(ev, ref)  # ln Z 1e-4 abs, errors 1e-3 rel vs the CPU oracle
    est, _ = ev.compute_ln_evidence()
    neg, pos = ev.compute_ln_inv_evidence_errors()
    assert abs(est - ln_z) < 4 * max(abs(neg), abs(pos))
    x = ch.samples[:20000]
    util.check_predict(m, x, m.predict(x))
    assert ev.check_basic_diagnostic()
