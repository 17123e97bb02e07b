"""Property checks of the flow oracle (parity unpinned: tfp/distrax/jax are absent; see oracle/__init__.py).
The reference pins predict only statistically: the flow density integrates to one at several
temperatures (tests/test_flow_model.py:155-206)."""
import numpy as np
import pytest

from harmonic_b200 import export
from oracle import flows as of


def grid2d(lim, n):
    g = np.linspace(-lim, lim, n)
    X, Y = np.meshgrid(g, g, indexing="ij")
    return np.stack([X.ravel(), Y.ravel()], axis=1), (g[1] - g[0]) ** 2


@pytest.mark.parametrize("T", [1.0, 0.8, 0.3])
def test_realnvp_density_normalised(T):
    spec = dict(kind="realnvp", params=export.random_realnvp_params(2, 2, 4, seed=3, weight_scale=0.7),
                temperature=T, standardize=True, pre_offset=np.array([0.3, -0.2]), pre_amp=np.array([1.3, 0.7]),
                n_scaled_layers=2, n_unscaled_layers=4)
    x, dA = grid2d(16.0, 801)
    total = np.sum(np.exp(of.predict(spec, x))) * dA
    assert total == pytest.approx(1.0, abs=2e-3)


@pytest.mark.parametrize("T", [1.0, 0.9, 0.3])
def test_rqspline_density_normalised(T):
    spec = dict(kind="rqspline", params=export.random_rqspline_params(2, 3, 8, (32, 32), seed=5),
                temperature=T, standardize=False, n_layers=3, n_bins=8, hidden_size=[32, 32],
                spline_range=(-10.0, 10.0))
    x, dA = grid2d(14.0, 901)
    total = np.sum(np.exp(of.predict(spec, x))) * dA
    assert total == pytest.approx(1.0, abs=2e-3)


def test_f32_mode_close_to_f64_and_nan_propagates():
    for spec in (
        dict(kind="realnvp", params=export.random_realnvp_params(5, 2, 4, seed=1), temperature=0.8,
             n_scaled_layers=2, n_unscaled_layers=4),
        dict(kind="rqspline", params=export.random_rqspline_params(5, 3, 8, (32, 32), seed=1),
             temperature=0.9, n_layers=3, n_bins=8, hidden_size=[32, 32], spline_range=(-10.0, 10.0)),
    ):
        x = np.random.default_rng(0).standard_normal((500, 5)) * 2
        a = of.predict(spec, x, np.float64)
        b = of.predict(spec, x, np.float32)
        assert b.dtype == np.float32
        assert np.max(np.abs(a - b) / np.maximum(np.abs(a), 1)) < 2e-5
        x[3, 2] = np.nan
        assert np.isnan(of.predict(spec, x)[3])
        one = of.predict(spec, x[0])
        assert np.ndim(one) == 0 and one == pytest.approx(a[0])


def test_num_masked_rounding():
    # tfp: int(np.round(D * 0.5)), round-half-even
    assert [of.num_masked(d) for d in (2, 3, 5, 7, 64)] == [1, 2, 2, 4, 32]
    assert [export.num_masked(d) for d in (2, 3, 5, 7, 64)] == [1, 2, 2, 4, 32]


def test_predict_temperature_guards():
    spec = dict(kind="realnvp", params=export.random_realnvp_params(2), temperature=1.5,
                n_scaled_layers=2, n_unscaled_layers=4)
    with pytest.raises(ValueError):
        of.predict(spec, np.zeros((1, 2)))
    spec["temperature"] = 0.0
    with pytest.raises(ValueError):
        of.predict(spec, np.zeros((1, 2)))


# ------------------------------------------------------------------ sampling direction (model.py:239-275)
def _fd_logdet(spec, e0):
    D = e0.size
    J = np.zeros((D, D))
    h = 1e-6
    for i in range(D):
        de = np.zeros(D)
        de[i] = h
        J[:, i] = (of.sample_transform(spec, (e0 + de)[None, :])[0] - of.sample_transform(spec, (e0 - de)[None, :])[0]) / (2 * h)
    return np.linalg.slogdet(J)[1]


@pytest.mark.parametrize("kind,ndim", [("realnvp", 2), ("realnvp", 5), ("realnvp", 16), ("rqspline", 2), ("rqspline", 5)])
def test_sample_direction_is_the_inverse_of_log_prob(kind, ndim):
    """Change of variables: log_prob(sample(eps)) = ln N(eps; 0, I) - ln |det d sample / d eps| with the Jacobian
    taken by finite differences, i.e. the sampling direction inverts the log_prob direction including every
    log-det term, temperature and (de)standardisation."""
    rng = np.random.default_rng(7)
    common = dict(standardize=True, pre_offset=rng.standard_normal(ndim) * 0.3, pre_amp=0.5 + rng.random(ndim))
    if kind == "realnvp":
        spec = dict(kind=kind, params=export.random_realnvp_params(ndim, 2, 4, seed=1, weight_scale=0.5),
                    temperature=0.8, n_scaled_layers=2, n_unscaled_layers=4, **common)
    else:
        spec = dict(kind=kind, params=export.random_rqspline_params(ndim, 3, 8, (16, 16), seed=2, weight_scale=0.6),
                    temperature=0.7, n_layers=3, n_bins=8, hidden_size=[16, 16], spline_range=(-10.0, 10.0), **common)
    eps = rng.standard_normal((4, ndim)) * 1.5
    x = of.sample_transform(spec, eps)
    lp = of.predict(spec, x)
    for r in range(eps.shape[0]):
        want = -0.5 * np.sum(eps[r] ** 2) - 0.5 * ndim * np.log(2 * np.pi) - _fd_logdet(spec, eps[r])
        assert lp[r] == pytest.approx(want, abs=2e-6 * max(1.0, abs(want)))


def test_rqspline_inverse_covers_tails_and_nan():
    """Linear tails outside the spline range and NaN propagation of the inverse spline."""
    rng = np.random.default_rng(3)
    K = 8
    P = rng.standard_normal((6, 3 * K + 1))
    x = np.array([-25.0, -10.0, -3.0, 4.0, 10.0, 31.0])
    y, _ = of._rqs_forward(x, P, K, -10.0, 10.0, np.float64)
    back = of._rqs_inverse(y, P, K, -10.0, 10.0, np.float64)
    np.testing.assert_allclose(back, x, rtol=1e-10, atol=1e-10)
    assert np.isnan(of._rqs_inverse(np.array([np.nan]), P[:1], K, -10.0, 10.0, np.float64))[0]


def test_sample_temperature_guards():
    spec = dict(kind="realnvp", params=export.random_realnvp_params(2, 1, 1, seed=1), temperature=1.2,
                n_scaled_layers=1, n_unscaled_layers=1)
    with pytest.raises(ValueError):
        of.sample_transform(spec, np.zeros((1, 2)))
    spec["temperature"] = 0.0
    with pytest.raises(ValueError):
        of.sample_transform(spec, np.zeros((1, 2)))
