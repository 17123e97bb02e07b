"""Numpy restatement of the accumulate + finalise half of the hot path (TEST INFRASTRUCTURE ONLY).

PINNED against the reference's own known-answer tests and against the unmodified
reference ``evidence.py`` run through a numpy stand-in for jax.numpy (see
``tests/golden/make_golden.py`` and ``tests/test_oracle_golden.py``).

Follows (paths into /root/reference):
* ``EvidenceOracle.add``                 -> ``harmonic/evidence.py:278-358`` (batched branch)
* ``EvidenceOracle.process_run``         -> ``harmonic/evidence.py:125-189``
* ``ln_evidence`` / ``ln_inv_evidence_errors`` / ``evidence`` -> ``:401-497``

``f32=True`` emulates the reference's JAX default arithmetic (x64 disabled): lnargs, the
shift, exp and the per-chain sums are float32, added into float64 running totals.
``f32=False`` (default) is the float64 oracle proper.
"""
import numpy as np
import scipy.special as sp

MEAN_SHIFT, MAX_SHIFT, MIN_SHIFT, ABS_MAX_SHIFT = 1, 2, 3, 4


class EvidenceOracle:
    def __init__(self, nchains, shift=MEAN_SHIFT, f32=False):
        if nchains < 1:
            raise ValueError("nchains must be greater than 0.")
        self.nchains = nchains
        self.shift = shift
        self.f32 = f32
        self.ft = np.float32 if f32 else np.float64
        self.running_sum = np.zeros(nchains)
        self.nsamples_per_chain = np.zeros(nchains)
        self.nsamples_eff_per_chain = np.zeros(nchains)
        self.shift_set = False
        self.shift_value = 0.0
        self.chains_added = False

    # evidence.py:100-123
    def set_shift(self, v):
        if not np.isfinite(v):
            raise ValueError("Shift must be a number")
        if self.shift_set:
            raise ValueError("Cannot define multiple shifts!")
        self.shift_value = v
        self.shift_set = True

    def add(self, lnpred, ln_posterior, start_indices):
        """One add_chains call given the flow's lnpred (N,) (evidence.py:278-358)."""
        ft = self.ft
        start = np.asarray(start_indices, dtype=np.int64)
        if len(start) != self.nchains + 1:
            raise ValueError("nchains do not match")
        lnpred = np.asarray(lnpred, dtype=ft)
        Y = np.asarray(ln_posterior, dtype=ft)
        with np.errstate(all="ignore"):
            lnargs = lnpred - Y
            lnargs = np.where(np.isinf(lnargs), ft(np.nan), lnargs)  # :282-283
            if not self.shift_set:  # :307-319
                if self.shift == MAX_SHIFT:
                    self.set_shift(-np.nanmax(lnargs))
                elif self.shift == MEAN_SHIFT:
                    self.set_shift(-np.nanmean(lnargs))
                elif self.shift == MIN_SHIFT:
                    self.set_shift(-np.nanmin(lnargs))
                elif self.shift == ABS_MAX_SHIFT:
                    self.set_shift(-lnargs[np.nanargmax(np.abs(lnargs))])
            lnargs = lnargs + ft(self.shift_value)  # :329
            terms = np.exp(lnargs)
            nan = np.isnan(lnargs)
            terms = np.where(nan, ft(0), terms)
            for c in range(self.nchains):  # :321-345, masks replaced by the row ranges they select
                a, b = start[c], start[c + 1]
                self.running_sum[c] += np.sum(terms[a:b], dtype=ft)
                self.nsamples_per_chain[c] += b - a
                self.nsamples_eff_per_chain[c] += (b - a) - np.count_nonzero(nan[a:b])
            self.lnargmax = np.nanmax(lnargs) if not nan.all() else np.nan  # :347-352
            self.lnargmin = np.nanmin(lnargs) if not nan.all() else np.nan
            self.lnprobmax = np.nanmax(Y)
            self.lnprobmin = np.nanmin(Y)
            self.lnpredictmax = np.nanmax(lnpred)
            self.lnpredictmin = np.nanmin(lnpred)
        self.process_run()
        self.chains_added = True

    def process_run(self):
        """evidence.py:125-189."""
        ft = self.ft
        n_c = self.nsamples_per_chain
        with np.errstate(all="ignore"):
            evidence_inv = ft(np.sum(self.running_sum, dtype=ft))
            nsamples = ft(np.sum(n_c, dtype=ft))
            evidence_inv = ft(evidence_inv / nsamples)
            self.ln_evidence_inv_per_chain = (
                np.log(self.running_sum.astype(ft)) - ft(self.shift_value) - np.log(n_c.astype(ft))
            )
            if self.f32:
                z = np.abs((self.running_sum / n_c).astype(ft) - evidence_inv).astype(np.float64)
            else:
                z = np.abs(self.running_sum / n_c - evidence_inv)
            y = z * n_c**0.25
            z = z * n_c**0.5
            lnV = sp.logsumexp(2.0 * np.log(z)) - np.log(nsamples)
            kur_ln = sp.logsumexp(4.0 * np.log(y)) - np.log(nsamples) - 2.0 * lnV
            kur = np.exp(kur_ln)
            self.kurtosis = kur
            self.ln_kurtosis = kur_ln
            n_eff = ft(np.sum(np.square(n_c), dtype=ft))
            n_eff = ft(nsamples * nsamples / n_eff)
            self.n_eff = n_eff
            sv = self.shift_value
            self.ln_evidence_inv = np.log(evidence_inv) - sv
            self.ln_evidence_inv_var = lnV - 2 * sv - np.log(n_eff - 1)
            self.ln_evidence_inv_var_var = (
                2.0 * lnV - 3.0 * np.log(n_eff) - 4.0 * sv + np.log((kur - 1) + 2.0 / (n_eff - 1))
            )

    # evidence.py:401-436
    def evidence(self):
        ei = np.exp(self.ln_evidence_inv)
        ev = np.exp(self.ln_evidence_inv_var)
        if np.isinf(np.nan_to_num(ei, nan=np.inf)) or np.isinf(np.nan_to_num(ev, nan=np.inf)):
            raise ValueError(
                "Evidence is too large to represent in non-log space. Use log-space values instead."
            )
        common = 1.0 + ev / ei**2
        return common / ei, np.sqrt(ev) / ei**2

    # evidence.py:438-459
    def ln_evidence(self):
        ln_x = self.ln_evidence_inv_var - 2.0 * self.ln_evidence_inv
        return (
            np.log(1.0 + np.exp(ln_x)) - self.ln_evidence_inv,
            0.5 * self.ln_evidence_inv_var - 2.0 * self.ln_evidence_inv,
        )

    # evidence.py:461-497
    def ln_inv_evidence_errors(self):
        ratio = np.exp(0.5 * self.ln_evidence_inv_var - self.ln_evidence_inv)
        with np.errstate(all="ignore"):
            neg = np.log(1.0 - ratio) if np.abs(ratio - 1.0) > 1e-8 else -np.inf
            pos = np.log(1.0 + ratio)
        return neg, pos


def hypersphere_lnpred(X, centre, inv_cov, R, ln_one_over_volume):
    """HyperSphere.predict for rows of X (harmonic/model_legacy.pyx:309-330):
    ln(1/V) inside the ellipsoid of squared radius R^2, -inf outside."""
    dx = X - centre[None, :]
    d2 = np.sum(dx * dx * inv_cov[None, :], axis=1)
    return np.where(d2 < R * R, ln_one_over_volume, -np.inf)
