"""Stand-in for harmonic/model.py: same constructor signatures and attributes (harmonic/model.py:277-404), synthetic
weights under the reference's flax names, predict through the numpy oracle on the parameter tree."""
import numpy as np

from harmonic_b200 import export
from oracle import flows as of


class _State:
    def __init__(self, params):
        self.params = params


class _FlowModel:
    def __init__(self, ndim_in, standardize, temperature):
        self.ndim = ndim_in
        self.standardize = standardize
        self.temperature = temperature
        self.fitted = False
        self.state = None
        self.pre_offset = None
        self.pre_amp = None

    def _random_params(self, seed):
        raise NotImplementedError

    def _spec(self):
        raise NotImplementedError

    def fit(self, X, batch_size=64, epochs=3, key=0, verbose=False):
        X = np.asarray(X)
        if self.standardize:  # harmonic/model.py:150-155
            self.pre_amp = np.sqrt(np.diag(np.atleast_2d(np.cov(X.T))))
            self.pre_offset = np.mean(X, axis=0)
        self.state = _State(self._random_params(int(key) + 7))
        self.fitted = True

    def predict(self, x):
        spec = dict(self._spec(), params=self.state.params, temperature=self.temperature,
                    standardize=self.standardize, pre_offset=self.pre_offset, pre_amp=self.pre_amp)
        return of.predict(spec, np.asarray(x), np.float32)


class RealNVPModel(_FlowModel):
    def __init__(self, ndim_in, n_scaled_layers=2, n_unscaled_layers=4, learning_rate=0.001, momentum=0.9,
                 standardize=False, temperature=0.8):
        _FlowModel.__init__(self, ndim_in, standardize, temperature)
        self.n_scaled_layers = n_scaled_layers
        self.n_unscaled_layers = n_unscaled_layers

    def _random_params(self, seed):
        return export.random_realnvp_params(self.ndim, self.n_scaled_layers, self.n_unscaled_layers, seed=seed,
                                            weight_scale=0.4)

    def _spec(self):
        return dict(kind="realnvp", n_scaled_layers=self.n_scaled_layers, n_unscaled_layers=self.n_unscaled_layers)


class RQSplineModel(_FlowModel):
    def __init__(self, ndim_in, n_layers=8, n_bins=8, hidden_size=[64, 64], spline_range=(-10.0, 10.0),
                 standardize=False, learning_rate=0.001, momentum=0.9, temperature=0.8):
        _FlowModel.__init__(self, ndim_in, standardize, temperature)
        self.n_layers = n_layers
        self.n_bins = n_bins
        self.hidden_size = hidden_size
        self.spline_range = spline_range

    def _random_params(self, seed):
        return export.random_rqspline_params(self.ndim, self.n_layers, self.n_bins, tuple(self.hidden_size), seed=seed,
                                             weight_scale=0.4)

    def _spec(self):
        return dict(kind="rqspline", n_layers=self.n_layers, n_bins=self.n_bins, hidden_size=list(self.hidden_size),
                    spline_range=tuple(self.spline_range))
