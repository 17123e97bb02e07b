import numpy as np

import harmonic_b200 as hm
from bench_workloads import load_cfg1_model, make_realnvp, make_rqspline  # noqa: F401  (re-exported)
from oracle import evidence as oe
from oracle import flows as of


def oracle_predict(model, x, dtype=np.float64):
    return of.predict(model.oracle_spec(), x, dtype)


def oracle_evidence(model, chains, shift=1, calls=None):
    """Float64 oracle of Evidence.add_chains for host chains (optionally several calls)."""
    ev = oe.EvidenceOracle(chains.nchains if calls is None else calls[0].nchains, shift)
    for ch in (calls if calls is not None else [chains]):
        lnpred = oracle_predict(model, np.asarray(ch.samples, dtype=np.float64))
        # the reference forms lnargs in float32 from float32 lnpred; the oracle keeps float64
        ev.add(lnpred, np.asarray(ch.ln_posterior, dtype=np.float64), ch.start_indices)
    return ev


def gaussian_chains(ndim, nchains, nsamples, seed=0, ragged=False):
    rng = np.random.default_rng(seed)
    ch = hm.Chains(ndim)
    for c in range(nchains):
        n = nsamples if not ragged else int(nsamples * (0.5 + rng.random()))
        X = rng.standard_normal((n, ndim))
        Y = -0.5 * np.sum(X * X, axis=1)
        ch.add_chain(X, Y)
    return ch


def oracle_sample(model, eps, dtype=np.float64):
    return of.sample_transform(model.oracle_spec(), eps, dtype)


def check_sample(model, eps, got, rel=1e-5):
    """got (n, ndim) vs the float64 oracle of FlowModel.sample for the same standard-normal draws: 1e-5
    relative per coordinate (floor |x| at 1), widened by 3x the deviation of the float32 restatement."""
    want = oracle_sample(model, eps, np.float64)
    w32 = oracle_sample(model, eps, np.float32).astype(np.float64)
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == want.shape
    nan_w = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan_w), "NaN pattern differs"
    ok = ~nan_w
    den = np.maximum(np.abs(want[ok]), 1.0)
    err = np.abs(got[ok] - want[ok]) / den
    tol = rel + 3.0 * np.abs(w32[ok] - want[ok]) / den
    bad = err > tol
    assert not bad.any(), "max rel err %.3g, %d coordinates over tolerance" % (err.max(), int(bad.sum()))
    return 0.0 if err.size == 0 else float(err.max())


def check_predict(model, x, got, rel=1e-5, slack_mult=1.0):
    """got vs the float64 oracle at the north_star tolerance (1e-5 relative), widened per sample
    where float32 evaluation is ill-conditioned: by 3x the deviation of the float32 restatement (the
    arithmetic the reference itself runs in) and by 2x the change of ln phi under +-1 ulp(float32)
    relative perturbations of the input."""
    want = oracle_predict(model, x, np.float64)
    with np.errstate(all="ignore"):
        den = np.maximum(np.abs(want), 1.0)
        w32 = oracle_predict(model, x, np.float32).astype(np.float64)
        slack = 3.0 * np.abs(w32 - want) / den
        # conditioning: change of ln phi under a one-ulp (float32) relative perturbation of the input
        x64 = np.asarray(x, dtype=np.float64)
        for sgn in (1.0, -1.0):
            wp = oracle_predict(model, x64 * (1.0 + sgn * 2.0**-23), np.float64)
            slack = slack + 2.0 * np.abs(wp - want) / den
    slack[~np.isfinite(slack)] = 0.0
    return assert_lnphi_close(got, want, rel, slack_mult * slack)


def assert_lnphi_close(got, want, rel=1e-5, slack=None):
    """north_star tolerance: per-sample ln phi within 1e-5 relative (floor |ln phi| at 1)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    nan_w = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan_w), "NaN pattern differs"
    inf_w = np.isinf(want)
    assert np.array_equal(got[inf_w], want[inf_w])
    ok = ~(nan_w | inf_w)
    err = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1.0)
    tol = rel if slack is None else rel + np.asarray(slack)[ok]
    bad = err > tol
    assert not bad.any(), "max rel err %.3g (at %d), %d samples over tolerance" % (
        err.max(), int(np.argmax(err)), int(bad.sum()))
    return 0.0 if err.size == 0 else float(err.max())


def assert_evidence_close(ev, ref, ln_abs=1e-4, err_rel=1e-3):
    """ln Z within 1e-4 absolute; reported error estimates within 1e-3 relative."""
    lnz, lnz_std = ev.compute_ln_evidence()
    rz, rz_std = ref.ln_evidence()
    assert abs(lnz - rz) <= ln_abs, (lnz, rz)
    assert abs(ev.ln_evidence_inv - ref.ln_evidence_inv) <= ln_abs
    # error estimates: sigma = exp(0.5 ln var) relative 1e-3  <=> |d ln var| <= 2e-3
    assert abs(ev.ln_evidence_inv_var - ref.ln_evidence_inv_var) <= 2 * err_rel
    assert abs(lnz_std - rz_std) <= 2 * err_rel
    assert abs(ev.ln_evidence_inv_var_var - ref.ln_evidence_inv_var_var) <= 4 * err_rel
    neg, pos = ev.compute_ln_inv_evidence_errors()
    rneg, rpos = ref.ln_inv_evidence_errors()
    np.testing.assert_allclose([neg, pos], [rneg, rpos], rtol=err_rel, atol=1e-12)
    np.testing.assert_allclose(ev.nsamples_per_chain, ref.nsamples_per_chain)
    np.testing.assert_allclose(ev.nsamples_eff_per_chain, ref.nsamples_eff_per_chain)
