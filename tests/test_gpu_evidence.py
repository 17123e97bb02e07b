"""GPU parity: accumulation + finalise kernels vs the golden vectors produced by the unmodified
reference (tests/golden/evidence_golden.npz) and vs the float64 oracle."""
import os
import pickle

import numpy as np
import pytest

import harmonic_b200 as hm
from oracle import evidence as oe
from tests import util

pytestmark = pytest.mark.gpu

MODES = {"MEAN_SHIFT": hm.Shifting.MEAN_SHIFT, "MAX_SHIFT": hm.Shifting.MAX_SHIFT,
         "MIN_SHIFT": hm.Shifting.MIN_SHIFT, "ABS_MAX_SHIFT": hm.Shifting.ABS_MAX_SHIFT}
KEYS = ["running_sum", "nsamples_per_chain", "nsamples_eff_per_chain", "shift_value", "ln_evidence_inv",
        "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff",
        "ln_evidence_inv_per_chain", "lnargmax", "lnargmin", "lnprobmax", "lnprobmin", "lnpredictmax",
        "lnpredictmin"]


class Dummy:
    """fitted model with a flow attribute (only ndim / is_fitted are touched by add_lnpred)."""
    ndim = 2
    flow = object()

    def is_fitted(self):
        return True


def check(ev, g, prefix, rtol=2e-6):
    # inputs are folded in float32 on the device (like the reference's JAX path); golden is float64
    for k in KEYS:
        np.testing.assert_allclose(np.asarray(getattr(ev, k), dtype=np.float64), g[prefix + "/" + k],
                                   rtol=rtol, atol=2e-6, err_msg=prefix + "/" + k)
    np.testing.assert_allclose(np.array(ev.compute_ln_evidence()), g[prefix + "/ln_evidence"], rtol=rtol, atol=2e-6)
    np.testing.assert_allclose(np.array(ev.compute_ln_inv_evidence_errors()), g[prefix + "/ln_inv_errors"], rtol=1e-4)


def case_a(golden):
    np.random.seed(30)
    X = np.random.randn(200, 500, 2)
    Y = -np.sum(X * X, axis=2) / 2.0
    hs = (golden["hs/centre"], golden["hs/inv_covariance"], float(golden["hs/R"]),
          float(golden["hs/ln_one_over_volume"]))
    return X, Y, hs


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_case_a_golden(golden, mode, dtype):
    X, Y, hs = case_a(golden)
    lnpred = oe.hypersphere_lnpred(X.reshape(-1, 2), *hs)
    ev = hm.Evidence(200, Dummy(), MODES[mode])
    ev.add_lnpred(lnpred.astype(dtype), Y.reshape(-1).astype(dtype), np.arange(201) * 500)
    check(ev, golden, "A_batched_" + mode)
    if mode == "MEAN_SHIFT":  # the reference's pinned constants, tests/test_evidence.py:177-181
        assert np.exp(ev.ln_evidence_inv) == pytest.approx(0.159438606)
        assert np.exp(ev.ln_evidence_inv_var) == pytest.approx(1.164628268e-07, rel=1e-5)
        assert np.exp(ev.ln_evidence_inv_var_var) ** 0.5 == pytest.approx(1.142786462e-08, rel=1e-5)


def test_case_a_streaming_and_pickle(golden, tmp_path):
    X, Y, hs = case_a(golden)
    ev = hm.Evidence(200, Dummy())
    Xs, Ys = X[:, :300, :], Y[:, :300]
    ev.add_lnpred(oe.hypersphere_lnpred(Xs.reshape(-1, 2), *hs), Ys.reshape(-1), np.arange(201) * 300)
    # checkpoint between the two calls (notebooks/checkpointing.ipynb pattern)
    fn = os.path.join(tmp_path, "ev.pkl")
    ev.serialize(fn)
    ev2 = hm.Evidence.deserialize(fn)
    Xs, Ys = X[:, 300:, :], Y[:, 300:]
    for e in (ev, ev2):
        e.add_lnpred(oe.hypersphere_lnpred(Xs.reshape(-1, 2), *hs), Ys.reshape(-1), np.arange(201) * 200)
        check(e, golden, "A_stream")
        assert np.exp(e.ln_evidence_inv) == pytest.approx(0.159438606)
    assert ev.ln_evidence_inv == ev2.ln_evidence_inv


def test_case_a_user_model_per_sample_branch(golden):
    """A model WITHOUT a flow goes through per-sample predict (harmonic/evidence.py:285-303)."""
    X, Y, hs = case_a(golden)

    class Sphere:
        ndim = 2

        def is_fitted(self):
            return True

        def predict(self, x):
            return float(oe.hypersphere_lnpred(x[None, :], *hs)[0])

    ch = hm.Chains(2)
    ch.add_chains_3d(X[:40], Y[:40])
    ev = hm.Evidence(40, Sphere())
    assert not ev.batch_calculation
    ev.add_chains(ch)
    ref = oe.EvidenceOracle(40)
    ref.add(oe.hypersphere_lnpred(ch.samples, *hs), ch.ln_posterior, ch.start_indices)
    util.assert_evidence_close(ev, ref)
    # reference sets inf predictions to NaN in this branch -> predict extrema are finite
    assert np.isfinite(ev.lnpredictmax) and np.isfinite(ev.lnpredictmin)


@pytest.mark.parametrize("mode", list(MODES))
def test_case_b_ragged_nonfinite(golden, mode):
    lens = golden["B/lens"]
    start = np.concatenate([[0], np.cumsum(lens)])
    m = Dummy()
    m.ndim = 3
    ev = hm.Evidence(len(lens), m, MODES[mode])
    ev.add_lnpred(golden["B/lnpred"], golden["B/ln_posterior"], start)
    check(ev, golden, "B_" + mode)
    assert ev.nsamples_per_chain.sum() == lens.sum()
    assert ev.nsamples_eff_per_chain.sum() == lens.sum() - 6


def test_process_run_kat(golden):
    """tests/test_evidence.py:95-144 through Evidence.process_run (finalise kernel, real-space)."""
    from scipy.stats import kurtosis
    nchains, n_samples = 10, 20
    rho = hm.Evidence(nchains, Dummy())
    np.random.seed(1)
    samples = np.random.randn(nchains, n_samples)
    rho.running_sum = np.sum(samples, axis=1)
    rho.nsamples_per_chain = np.ones(nchains, dtype=int) * n_samples
    rho.process_run()
    var = np.std(np.sum(samples, axis=1) / n_samples) ** 2 / (nchains - 1)
    varvar = var**2 * (kurtosis(np.sum(samples, axis=1) / n_samples) + 2 + 2.0 / (nchains - 1)) \
        * (nchains - 1) ** 2 / nchains**3
    assert np.exp(rho.ln_evidence_inv) == pytest.approx(np.mean(samples), abs=1e-7)
    assert np.exp(rho.ln_evidence_inv_var) == pytest.approx(var)
    assert np.exp(rho.ln_evidence_inv_var_var) == pytest.approx(varvar)
    for k in ("ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff"):
        np.testing.assert_allclose(getattr(rho, k), golden["C1/" + k], rtol=1e-10)
    rho = hm.Evidence(nchains, Dummy())
    rho.running_sum = golden["C2/running_sum"].copy()
    rho.nsamples_per_chain = np.ones(nchains) * n_samples
    rho.shift_value = float(golden["C2/shift_value"])
    rho.process_run()
    for k in ("ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff"):
        np.testing.assert_allclose(getattr(rho, k), golden["C2/" + k], rtol=1e-10)


def test_many_tiny_and_empty_chains():
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 7, size=3000)
    lens[::97] = 5000  # a few long ones spanning several CTAs
    start = np.concatenate([[0], np.cumsum(lens)])
    n = start[-1]
    lnpred = rng.standard_normal(n) * 3
    lnpost = rng.standard_normal(n) * 3
    lnpred[rng.integers(0, n, 50)] = np.nan
    m = Dummy()
    ev = hm.Evidence(len(lens), m)
    ev.add_lnpred(lnpred, lnpost, start)
    part = ev.partials()
    # per-chain logsumexp vs numpy float64 on the float32-rounded inputs
    la = lnpred.astype(np.float32) - lnpost.astype(np.float32)
    for c in rng.integers(0, len(lens), 300):
        seg = la[start[c]:start[c + 1]].astype(np.float64)
        seg = seg[np.isfinite(seg)]
        want = -np.inf if seg.size == 0 else np.log(np.sum(np.exp(seg - seg.max()))) + seg.max()
        got = part[c, 0] + np.log(part[c, 1]) if part[c, 1] > 0 else -np.inf
        assert got == pytest.approx(want, rel=1e-6, abs=1e-6)
        assert part[c, 2] == lens[c]
    assert part[:, 2].sum() == n


def test_device_resident_inputs():
    import torch
    rng = np.random.default_rng(3)
    start = np.arange(0, 64 * 1000 + 1, 1000)
    lnpred = rng.standard_normal(64000)
    lnpost = rng.standard_normal(64000)
    a = hm.Evidence(64, Dummy())
    a.add_lnpred(lnpred.astype(np.float32), lnpost.astype(np.float32), start)
    b = hm.Evidence(64, Dummy())
    b.add_lnpred(torch.from_numpy(lnpred.astype(np.float32)).cuda(), torch.from_numpy(lnpost.astype(np.float32)).cuda(), start)
    assert a.ln_evidence_inv == b.ln_evidence_inv
    assert a.ln_evidence_inv_var == b.ln_evidence_inv_var


def test_errors():
    with pytest.raises(ValueError):
        hm.Evidence(0, Dummy())
    ev = hm.Evidence(4, Dummy())
    with pytest.raises(ValueError):
        ev.set_shift(np.nan)
    ev.set_shift(1.0)
    with pytest.raises(ValueError):
        ev.set_shift(2.0)
    ch = util.gaussian_chains(2, 5, 10)
    with pytest.raises(ValueError):
        ev.add_chains(ch)  # nchains mismatch


def test_split_data_device_resident_chains():
    """utils.split_data on device-resident chains is index arithmetic; the evidence of the test half equals the
    host-chains result."""
    import torch
    from tests import util as u
    ndim = 5
    m = u.make_realnvp(ndim, seed=2, standardize=True, ws=0.3)
    ch = u.gaussian_chains(ndim, 20, 400, seed=5, ragged=True)
    _, te = hm.utils.split_data(ch, 0.5)
    ev_h = hm.Evidence(te.nchains, m)
    ev_h.add_chains(te)
    xs = torch.from_numpy(ch.samples.astype(np.float32)).cuda()
    ys = torch.from_numpy(ch.ln_posterior.astype(np.float32)).cuda()
    dch = hm.Chains.from_arrays(xs, ys, ch.start_indices)
    _, dte = hm.utils.split_data(dch, 0.5)
    assert dte.samples.is_cuda and dte.samples.data_ptr() == xs[ch.start_indices[10]:].data_ptr()
    ev_d = hm.Evidence(dte.nchains, m)
    ev_d.add_chains(dte)
    assert ev_d.ln_evidence_inv == pytest.approx(ev_h.ln_evidence_inv, abs=1e-5)
    np.testing.assert_allclose(ev_d.nsamples_per_chain, ev_h.nsamples_per_chain)


def test_add_chains_batch_equals_separate_calls():
    """Two models of a Bayes factor (different flows, different chains) folded at the same time on two streams /
    host threads (evidence.add_chains_batch) give, bit for bit, what the reference's order --
    one add_chains after the other (harmonic/evidence.py:534-664 then compares the two objects) -- gives; a third
    pair on the tensor path rides along."""
    m1 = util.make_realnvp(5, seed=31, standardize=True, ws=0.4, temperature=0.8)
    m2 = util.make_rqspline(5, 3, 8, (32, 32), seed=32, standardize=True, ws=0.4, temperature=0.9)
    m3 = util.make_realnvp(64, seed=33, standardize=True, ws=0.3)
    c1 = util.gaussian_chains(5, 40, 3000, seed=41, ragged=True)
    c2 = util.gaussian_chains(5, 25, 4000, seed=42, ragged=True)
    c3 = util.gaussian_chains(64, 16, 900, seed=43, ragged=True)
    seq = [hm.Evidence(c.nchains, m) for m, c in ((m1, c1), (m2, c2), (m3, c3))]
    for ev, c in zip(seq, (c1, c2, c3)):
        ev.add_chains(c)
    bat = [hm.Evidence(c.nchains, m) for m, c in ((m1, c1), (m2, c2), (m3, c3))]
    hm.evidence.add_chains_batch(list(zip(bat, (c1, c2, c3))))
    for a, b in zip(seq, bat):
        assert a.ln_evidence_inv == b.ln_evidence_inv
        assert a.ln_evidence_inv_var == b.ln_evidence_inv_var
        assert a.ln_evidence_inv_var_var == b.ln_evidence_inv_var_var
        np.testing.assert_array_equal(a.ln_evidence_inv_per_chain, b.ln_evidence_inv_per_chain)
        np.testing.assert_array_equal(a.nsamples_eff_per_chain, b.nsamples_eff_per_chain)
    # a second streaming call, and the Bayes factor of the first two
    more1 = util.gaussian_chains(5, 40, 1000, seed=44)
    more2 = util.gaussian_chains(5, 25, 1000, seed=45)
    seq[0].add_chains(more1)
    seq[1].add_chains(more2)
    got = hm.evidence.compute_ln_bayes_factor_batch(bat[0], more1, bat[1], more2)
    assert got == hm.evidence.compute_ln_bayes_factor(seq[0], seq[1])
    with pytest.raises(ValueError):
        hm.evidence.add_chains_batch([(bat[0], more1), (bat[0], more1)])
    with pytest.raises(ValueError):  # the reference's own check, raised through the worker thread
        hm.evidence.add_chains_batch([(bat[0], c2)])
