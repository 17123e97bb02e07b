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
