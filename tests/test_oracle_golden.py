"""Pin the CPU oracle (accumulate + finalise half) against the reference.

Golden values come from the UNMODIFIED reference evidence.py / chains.py / HyperSphere run in
the build container (tests/golden/make_golden.py), and from the constants pinned in the
reference's own tests (tests/test_evidence.py:177-181, 95-144, 283-329).
"""
import numpy as np
import pytest
from scipy.stats import kurtosis

from oracle import evidence as oe

STATE_KEYS = [
    "running_sum", "nsamples_per_chain", "nsamples_eff_per_chain", "shift_value",
    "ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis",
    "n_eff", "ln_evidence_inv_per_chain", "lnargmax", "lnargmin", "lnprobmax", "lnprobmin",
    "lnpredictmax", "lnpredictmin",
]
MODES = {"MEAN_SHIFT": 1, "MAX_SHIFT": 2, "MIN_SHIFT": 3, "ABS_MAX_SHIFT": 4}


def check_state(ev, g, prefix, rtol=1e-10):
    for k in STATE_KEYS:
        np.testing.assert_allclose(
            np.asarray(getattr(ev, k), dtype=np.float64), g[prefix + "/" + k], rtol=rtol, atol=1e-12,
            err_msg=prefix + "/" + k,
        )
    np.testing.assert_allclose(np.array(ev.ln_evidence()), g[prefix + "/ln_evidence"], rtol=rtol)
    np.testing.assert_allclose(np.array(ev.ln_inv_evidence_errors()), g[prefix + "/ln_inv_errors"], rtol=rtol)


def case_a():
    np.random.seed(30)
    X = np.random.randn(200, 500, 2)
    Y = -np.sum(X * X, axis=2) / 2.0
    return X, Y


def test_hypersphere_reference_constants(golden):
    # reference pins (tests/test_evidence.py:177-181) reproduced by the reference run itself
    assert np.exp(golden["A_legacy/ln_evidence_inv"]) == pytest.approx(0.159438606)
    assert np.exp(golden["A_legacy/ln_evidence_inv_var"]) == pytest.approx(1.164628268e-07)
    assert np.exp(golden["A_legacy/ln_evidence_inv_var_var"]) ** 0.5 == pytest.approx(1.142786462e-08)


@pytest.mark.parametrize("mode", list(MODES))
def test_oracle_case_a_all_shifts(golden, mode):
    X, Y = case_a()
    lnpred = oe.hypersphere_lnpred(
        X.reshape(-1, 2), golden["hs/centre"], golden["hs/inv_covariance"],
        float(golden["hs/R"]), float(golden["hs/ln_one_over_volume"]),
    )
    ev = oe.EvidenceOracle(200, MODES[mode])
    ev.add(lnpred, Y.reshape(-1), np.arange(201) * 500)
    check_state(ev, golden, "A_batched_" + mode)
    if mode == "MEAN_SHIFT":
        assert np.exp(ev.ln_evidence_inv) == pytest.approx(0.159438606)
        assert np.exp(ev.ln_evidence_inv_var) == pytest.approx(1.164628268e-07)
        assert np.exp(ev.ln_evidence_inv_var_var) ** 0.5 == pytest.approx(1.142786462e-08)
        # float32 emulation of the reference's JAX arithmetic stays inside its own tolerance
        ev32 = oe.EvidenceOracle(200, MODES[mode], f32=True)
        ev32.add(lnpred, Y.reshape(-1), np.arange(201) * 500)
        assert np.exp(ev32.ln_evidence_inv) == pytest.approx(0.159438606, rel=1e-5)


def test_oracle_streaming(golden):
    X, Y = case_a()
    hs = (golden["hs/centre"], golden["hs/inv_covariance"], float(golden["hs/R"]),
          float(golden["hs/ln_one_over_volume"]))
    ev = oe.EvidenceOracle(200)
    for sl in (slice(0, 300), slice(300, 500)):
        Xs, Ys = X[:, sl, :], Y[:, sl]
        n = Xs.shape[1]
        ev.add(oe.hypersphere_lnpred(Xs.reshape(-1, 2), *hs), Ys.reshape(-1), np.arange(201) * n)
    check_state(ev, golden, "A_stream")
    assert np.exp(ev.ln_evidence_inv) == pytest.approx(0.159438606)


@pytest.mark.parametrize("mode", list(MODES))
def test_oracle_case_b_ragged_nonfinite(golden, mode):
    lens = golden["B/lens"]
    start = np.concatenate([[0], np.cumsum(lens)])
    ev = oe.EvidenceOracle(len(lens), MODES[mode])
    ev.add(golden["B/lnpred"], golden["B/ln_posterior"], start)
    check_state(ev, golden, "B_" + mode)
    # NaN samples decrement the effective count only (tests/test_evidence.py:478-516)
    assert ev.nsamples_per_chain.sum() == lens.sum()
    assert ev.nsamples_eff_per_chain.sum() == lens.sum() - 6


def test_oracle_process_run_kat(golden):
    # tests/test_evidence.py:95-144 (inputs regenerated from the seeds, answers from scipy)
    nchains, n_samples = 10, 20
    np.random.seed(1)
    samples = np.random.randn(nchains, n_samples)
    ev = oe.EvidenceOracle(nchains)
    ev.running_sum = np.sum(samples, axis=1)
    ev.nsamples_per_chain = np.ones(nchains) * n_samples
    ev.process_run()
    var = np.std(np.sum(samples, axis=1) / n_samples) ** 2 / (nchains - 1)
    varvar = var**2 * (kurtosis(np.sum(samples, axis=1) / n_samples) + 2 + 2.0 / (nchains - 1)) \
        * (nchains - 1) ** 2 / nchains**3
    assert np.exp(ev.ln_evidence_inv) == pytest.approx(np.mean(samples), abs=1e-7)
    assert np.exp(ev.ln_evidence_inv_var) == pytest.approx(var)
    assert np.exp(ev.ln_evidence_inv_var_var) == pytest.approx(varvar)
    for k in ("ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff"):
        np.testing.assert_allclose(getattr(ev, k), golden["C1/" + k], rtol=1e-12)
    ev = oe.EvidenceOracle(nchains)
    ev.running_sum = golden["C2/running_sum"].copy()
    ev.nsamples_per_chain = np.ones(nchains) * n_samples
    ev.shift_value = float(golden["C2/shift_value"])
    ev.process_run()
    for k in ("ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff"):
        np.testing.assert_allclose(getattr(ev, k), golden["C2/" + k], rtol=1e-12)


def test_oracle_error_accessor_edge():
    # tests/test_evidence.py:283-329: ratio == 1 -> -inf
    ev = oe.EvidenceOracle(3)
    ev.ln_evidence_inv, ev.ln_evidence_inv_var = 1.0, 2.0
    neg, pos = ev.ln_inv_evidence_errors()
    assert neg == -np.inf and pos == pytest.approx(np.log(2.0))
    ev.ln_evidence_inv, ev.ln_evidence_inv_var = np.log(10.0), np.log(4.0)
    neg, pos = ev.ln_inv_evidence_errors()
    assert neg == pytest.approx(np.log(1 - 0.2)) and pos == pytest.approx(np.log(1.2))
