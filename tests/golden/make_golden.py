"""Generate tests/golden/evidence_golden.npz by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference and oracle/_ref built by
``python oracle/build_ref.py``).  The committed .npz is what travels; nothing in the
test-suite reads /root/reference.

What runs from the reference:
* ``harmonic/evidence.py`` (class Evidence, loaded by path, untouched).  Its ``jax`` /
  ``jax.numpy`` imports are satisfied by a small numpy stand-in defined below (jax is not
  installed in this image), so the arithmetic is the reference ALGORITHM in float64, not
  JAX's float32 rounding.
* ``harmonic/chains.py`` (class Chains, untouched).
* ``harmonic/model_legacy.pyx`` HyperSphere, compiled from the reference sources into
  ``oracle/_ref`` (the model the reference's own golden numbers are pinned with,
  ``tests/test_evidence.py:147-205``).
The batched branch (``evidence.py:260-283``) is exercised with a stand-in model that has a
``flow`` attribute and returns a precomputed lnpred.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


# ----------------------------------------------------------------- numpy stand-in for jax
class JArr(np.ndarray):
    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def set(self, v):
        out = np.array(self.a, copy=True)
        out[np.asarray(self.idx)] = v
        return out.view(JArr)


def _wrap(fn):
    def f(*a, **k):
        r = fn(*a, **k)
        return r.view(JArr) if isinstance(r, np.ndarray) else r

    return f


def install_fake_jax():
    jnp = types.ModuleType("jax.numpy")
    for name in (
        "sum log square arange array asarray concatenate isinf isnan nansum where exp "
        "nanmax nanmin abs sqrt zeros ones diff"
    ).split():
        setattr(jnp, name, _wrap(getattr(np, name)))
    jnp.nan = np.nan
    jnp.inf = np.inf
    jnp.ndarray = np.ndarray
    jax = types.ModuleType("jax")
    jax.numpy = jnp

    def vmap(fn, in_axes):
        def g(*args):
            n = [a.shape[0] for a, ax in zip(args, in_axes) if ax == 0][0]
            return np.array(
                [fn(*[a[i] if ax == 0 else a for a, ax in zip(args, in_axes)]) for i in range(n)]
            )

        return g

    jax.vmap = vmap
    jax.Array = type("Array", (), {})  # scipy probes sys.modules["jax"].Array
    sys.modules["jax"] = jax
    sys.modules["jax.numpy"] = jnp


def load_reference():
    install_fake_jax()
    pkg = types.ModuleType("harmonic")
    pkg.__path__ = [os.path.join(ROOT, "oracle", "_ref", "harmonic"), os.path.join(REF, "harmonic")]
    pkg.__file__ = os.path.join(REF, "harmonic", "__init__.py")
    sys.modules["harmonic"] = pkg
    mods = {}
    for name in ("logs", "model_abstract", "chains", "evidence"):
        spec = importlib.util.spec_from_file_location(
            "harmonic." + name, os.path.join(REF, "harmonic", name + ".py")
        )
        m = importlib.util.module_from_spec(spec)
        sys.modules["harmonic." + name] = m
        spec.loader.exec_module(m)
        setattr(pkg, name, m)
        mods[name] = m
    import harmonic.model_legacy as mdl  # compiled from reference sources (oracle/_ref)

    mods["model_legacy"] = mdl
    return mods


class PrecomputedFlowModel:
    """Stand-in with a ``flow`` attribute: selects the reference's batched branch."""

    flow = True
    fitted = True

    def __init__(self, ndim, table):
        self.ndim = ndim
        self.table = table  # id(x.ctypes.data) unsafe -> keyed by nrows + first value

    def is_fitted(self):
        return True

    def predict(self, x):
        return self.table[(x.shape[0], float(x[0, 0]))].view(JArr)


def state_of(ev):
    return dict(
        running_sum=np.asarray(ev.running_sum, dtype=np.float64),
        nsamples_per_chain=np.asarray(ev.nsamples_per_chain, dtype=np.float64),
        nsamples_eff_per_chain=np.asarray(ev.nsamples_eff_per_chain, dtype=np.float64),
        shift_value=np.float64(ev.shift_value),
        ln_evidence_inv=np.float64(ev.ln_evidence_inv),
        ln_evidence_inv_var=np.float64(ev.ln_evidence_inv_var),
        ln_evidence_inv_var_var=np.float64(ev.ln_evidence_inv_var_var),
        ln_kurtosis=np.float64(ev.ln_kurtosis),
        n_eff=np.float64(ev.n_eff),
        ln_evidence_inv_per_chain=np.asarray(ev.ln_evidence_inv_per_chain, dtype=np.float64),
        lnargmax=np.float64(ev.lnargmax),
        lnargmin=np.float64(ev.lnargmin),
        lnprobmax=np.float64(ev.lnprobmax),
        lnprobmin=np.float64(ev.lnprobmin),
        lnpredictmax=np.float64(ev.lnpredictmax),
        lnpredictmin=np.float64(ev.lnpredictmin),
        ln_evidence=np.array(ev.compute_ln_evidence(), dtype=np.float64),
        ln_inv_errors=np.array(ev.compute_ln_inv_evidence_errors(), dtype=np.float64),
    )


def main():
    import warnings

    warnings.simplefilter("ignore")
    m = load_reference()
    ch, cbe, mdl = m["chains"], m["evidence"], m["model_legacy"]
    out = {}

    def put(prefix, d):
        for k, v in d.items():
            out[prefix + "/" + k] = v

    # ---- case A: tests/test_evidence.py:147-205 (HyperSphere, seed 30, 200 x 500 x 2)
    nchains, nsamples, ndim = 200, 500, 2
    np.random.seed(30)
    X = np.random.randn(nchains, nsamples, ndim)
    Y = -np.sum(X * X, axis=2) / 2.0
    chain = ch.Chains(ndim)
    chain.add_chains_3d(X, Y)
    sphere = mdl.HyperSphere(ndim, [np.array([1e-1, 1e1])])
    sphere.fit(chain.samples, chain.ln_posterior)
    out["hs/centre"] = np.asarray(sphere.centre)
    out["hs/inv_covariance"] = np.asarray(sphere.inv_covariance)
    out["hs/R"] = np.float64(sphere.R)
    out["hs/ln_one_over_volume"] = np.float64(sphere.ln_one_over_volume)
    ev = cbe.Evidence(nchains, sphere, cbe.Shifting.MEAN_SHIFT)
    ev.add_chains(chain)  # legacy per-sample branch of the reference
    put("A_legacy", state_of(ev))
    lnpred = np.array([sphere.predict(x) for x in chain.samples])
    # same data through the reference's BATCHED branch, all four shift modes
    for mode in cbe.Shifting:
        model = PrecomputedFlowModel(ndim, {(chain.samples.shape[0], float(chain.samples[0, 0])): lnpred})
        ev = cbe.Evidence(nchains, model, mode)
        ev.add_chains(chain)
        put("A_batched_%s" % mode.name, state_of(ev))
    # two-call streaming (300 + 200 per chain), tests/test_evidence.py:183-203
    n1 = 300
    c1, c2 = ch.Chains(ndim), ch.Chains(ndim)
    for i in range(nchains):
        c1.add_chain(X[i, :n1, :], Y[i, :n1])
        c2.add_chain(X[i, n1:, :], Y[i, n1:])
    lp1 = np.array([sphere.predict(x) for x in c1.samples])
    lp2 = np.array([sphere.predict(x) for x in c2.samples])
    model = PrecomputedFlowModel(
        ndim,
        {
            (c1.samples.shape[0], float(c1.samples[0, 0])): lp1,
            (c2.samples.shape[0], float(c2.samples[0, 0])): lp2,
        },
    )
    ev = cbe.Evidence(nchains, model, cbe.Shifting.MEAN_SHIFT)
    ev.add_chains(c1)
    ev.add_chains(c2)
    put("A_stream", state_of(ev))

    # ---- case B: unequal chains + NaN / inf entries (tests/test_evidence.py:432-516 style)
    rng = np.random.default_rng(7)
    lens = [37, 1, 250, 64, 129, 5, 1000, 333]
    ndim = 3
    cB = ch.Chains(ndim)
    for L in lens:
        cB.add_chain(rng.standard_normal((L, ndim)), rng.standard_normal(L) * 3.0 - 5.0)
    lnpredB = rng.standard_normal(cB.nsamples) * 2.0 - 4.0
    lnpredB[[3, 40, 500, 1500]] = [np.nan, np.inf, -np.inf, np.nan]
    cB.ln_posterior[[10, 700]] = [np.inf, -np.inf]
    out["B/lens"] = np.array(lens)
    out["B/lnpred"] = lnpredB
    out["B/ln_posterior"] = cB.ln_posterior.copy()
    for mode in cbe.Shifting:
        model = PrecomputedFlowModel(ndim, {(cB.samples.shape[0], float(cB.samples[0, 0])): lnpredB})
        ev = cbe.Evidence(len(lens), model, mode)
        ev.add_chains(cB)
        put("B_%s" % mode.name, state_of(ev))

    # ---- case C: process_run known-answer inputs (tests/test_evidence.py:95-144)
    model = PrecomputedFlowModel(2, {})
    nchains, n_samples = 10, 20
    rho = cbe.Evidence(nchains, model)
    np.random.seed(1)
    samples = np.random.randn(nchains, n_samples)
    rho.running_sum = np.sum(samples, axis=1)
    rho.nsamples_per_chain = np.ones(nchains, dtype=int) * n_samples
    rho.process_run()
    out["C1/running_sum"] = rho.running_sum.copy()
    for k in ("ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff"):
        out["C1/" + k] = np.float64(getattr(rho, k))
    rho = cbe.Evidence(nchains, model, cbe.Shifting.MEAN_SHIFT)
    np.random.seed(1)
    post = np.random.uniform(high=1e3, size=(nchains, n_samples))
    mean_shift = np.mean(np.log(post))
    rho.running_sum = np.sum(np.exp(mean_shift) / post, axis=1)
    rho.nsamples_per_chain = np.ones(nchains, dtype=int) * n_samples
    rho.shift_value = mean_shift
    rho.process_run()
    out["C2/running_sum"] = rho.running_sum.copy()
    out["C2/shift_value"] = np.float64(mean_shift)
    for k in ("ln_evidence_inv", "ln_evidence_inv_var", "ln_evidence_inv_var_var", "ln_kurtosis", "n_eff"):
        out["C2/" + k] = np.float64(getattr(rho, k))

    np.savez_compressed(os.path.join(HERE, "evidence_golden.npz"), **out)
    print("wrote", len(out), "arrays;",
          "A_legacy rho/var/sqrt(varvar) =",
          np.exp(out["A_legacy/ln_evidence_inv"]),
          np.exp(out["A_legacy/ln_evidence_inv_var"]),
          np.exp(out["A_legacy/ln_evidence_inv_var_var"]) ** 0.5)


if __name__ == "__main__":
    main()
