"""CPU-only checks: the C-ABI library loads and exports every symbol include/harmonic_b200.h
declares, the host-side mirrors validate like the reference, and there is no CPU fallback."""
import os
import re

import numpy as np
import pytest

import harmonic_b200 as hm
from harmonic_b200 import _lib, export

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "harmonic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names  # ctypes table mirrors the header one to one
    assert lib.hb_version() == 200


def test_struct_layouts_match_header():
    assert _lib.C.sizeof(_lib.FlowDesc) == 4 * (7 + 8) + 4 * 2 + 4
    assert _lib.C.sizeof(_lib.Result) == 8 * 16


def test_weight_count_matches_exporter():
    lib = _lib.load()
    for ndim, ns, nu in [(2, 2, 4), (5, 6, 2), (64, 2, 4), (7, 1, 0)]:
        m = hm.model.RealNVPModel(ndim, ns, nu)
        w = export.flatten_realnvp(export.random_realnvp_params(ndim, ns, nu), ndim, ns, nu)
        assert lib.hb_flow_weight_count(_lib.C.byref(m._desc())) == w.size
    for ndim, L, K, hid in [(2, 8, 8, [64, 64]), (5, 3, 4, [16]), (3, 2, 8, [])]:
        m = hm.model.RQSplineModel(ndim, L, K, hid)
        w = export.flatten_rqspline(export.random_rqspline_params(ndim, L, K, hid), ndim, L, K, hid)
        assert lib.hb_flow_weight_count(_lib.C.byref(m._desc())) == w.size


@pytest.mark.skipif(_lib.load().hb_device_count() > 0, reason="needs a box WITHOUT a B200")
def test_no_cpu_fallback():
    m = hm.model.RealNVPModel(4)
    m.load_params(export.random_realnvp_params(4))
    with pytest.raises(_lib.HarmonicB200Error):
        m.predict(np.zeros((3, 4)))
    ev = hm.Evidence(2, m)
    ch = hm.Chains(4)
    ch.add_chains_2d(np.zeros((10, 4)), np.zeros(10), 2)
    with pytest.raises(_lib.HarmonicB200Error):
        ev.add_chains(ch)
    h = _lib.C.c_void_p()
    assert _lib.load().hb_accum_create(3, 0, _lib.C.byref(h)) == _lib.HB_ERR_NO_DEVICE
    assert b"no CPU fallback" in _lib.load().hb_last_error()


def test_model_constructor_errors_match_reference():
    # harmonic/model.py:106-107,311-312 ; tests/test_flow_model.py:86-110
    with pytest.raises(ValueError):
        hm.model.RealNVPModel(0)
    with pytest.raises(ValueError):
        hm.model.RQSplineModel(-1)
    with pytest.raises(ValueError):
        hm.model.RealNVPModel(3, n_scaled_layers=0)
    m = hm.model.RealNVPModel(3)
    assert not m.is_fitted() and hasattr(m, "flow")
    with pytest.raises(ValueError):
        m.fit(np.zeros((10, 4)))  # X second dimension not the same as ndim
    with pytest.raises(NotImplementedError):
        hm.model.FlowModel(3).fit(np.zeros((10, 3)))
    with pytest.raises(ValueError):  # wrong shapes are rejected by the exporter
        m.load_params(export.random_realnvp_params(5))
    with pytest.raises(ValueError):
        hm.model.RealNVPModel(3, standardize=True).load_params(export.random_realnvp_params(3))
    m.load_params(export.random_realnvp_params(3))
    assert m.is_fitted()
    with pytest.raises(ValueError):
        hm.Evidence(0, m)
    with pytest.raises(ValueError):
        hm.Evidence(3, hm.model.RealNVPModel(3))  # not fitted


def test_model_pickle_roundtrip(tmp_path):
    m = hm.model.RQSplineModel(3, 2, 4, [8], standardize=True, temperature=0.7)
    m.load_params(export.random_rqspline_params(3, 2, 4, [8]), np.zeros(3), np.ones(3))
    fn = os.path.join(tmp_path, "m.pkl")
    m.serialize(fn)
    m2 = hm.model.RQSplineModel.deserialize(fn)
    assert m2.is_fitted() and m2.temperature == 0.7 and m2._handle is None
    np.testing.assert_array_equal(m2._flat_weights(), m._flat_weights())


def test_chains_layout_and_helpers():
    # harmonic/chains.py semantics, tests/test_chains.py
    with pytest.raises(ValueError):
        hm.Chains(0)
    rng = np.random.default_rng(0)
    ch = hm.Chains(2)
    X = rng.standard_normal((3, 10, 2))
    Y = rng.standard_normal((3, 10))
    ch.add_chains_3d(X, Y)
    assert ch.nchains == 3 and ch.nsamples == 30 and ch.start_indices == [0, 10, 20, 30]
    assert ch.samples.dtype == np.float64 and ch.samples.flags.c_contiguous
    np.testing.assert_array_equal(ch.samples[10:20], X[1])
    ch.add_chain(rng.standard_normal((7, 2)), rng.standard_normal(7))
    assert ch.start_indices[-1] == 37 and ch.nchains == 4
    with pytest.raises(ValueError):
        ch.add_chain(np.zeros((3, 3)), np.zeros(3))
    with pytest.raises(ValueError):
        ch.add_chains_2d(np.zeros((10, 2)), np.zeros(10), 3)
    ch.add_chains_2d_list(rng.standard_normal((9, 2)), rng.standard_normal(9), 2, [0, 4, 9])
    assert list(ch.nsamples_per_chain()) == [10, 10, 10, 7, 4, 5]
    sub = ch.get_sub_chains([0, 3])
    assert sub.nchains == 2 and sub.nsamples == 17
    assert ch.get_chain_indices(3) == (30, 37)
    c2 = ch.deepcopy()
    c2.remove_burnin(2)
    assert list(c2.nsamples_per_chain()) == [8, 8, 8, 5, 2, 3]
    np.testing.assert_array_equal(c2.samples[:8], ch.samples[2:10])
    with pytest.raises(ValueError):
        c2.remove_burnin(2)
    c3 = hm.Chains(2)
    c3.add_chains_3d(X, Y)
    c3.split_into_blocks(6)
    assert c3.nchains == 6 and c3.start_indices == [0, 5, 10, 15, 20, 25, 30]
    with pytest.raises(ValueError):
        c3.split_into_blocks(3)
    wrapped = hm.Chains.from_arrays(X.reshape(-1, 2).astype(np.float32), Y.reshape(-1).astype(np.float32), [0, 10, 20, 30])
    assert wrapped.samples.dtype == np.float32 and wrapped.nchains == 3


def test_evidence_scalar_accessors():
    """compute_ln_evidence / compute_ln_inv_evidence_errors / compute_evidence / Bayes factors are
    host scalar math (harmonic/evidence.py:401-497,534-664; tests/test_evidence.py:262-379)."""
    m = hm.model.RealNVPModel(2)
    m.load_params(export.random_realnvp_params(2))
    ev = hm.Evidence(3, m)
    ev.nsamples_eff_per_chain = np.ones(3) * 100
    ev.ln_evidence_inv, ev.ln_evidence_inv_var = np.log(10.0), np.log(4.0)
    neg, pos = ev.compute_ln_inv_evidence_errors()
    assert neg == pytest.approx(np.log(0.8)) and pos == pytest.approx(np.log(1.2))
    ev.ln_evidence_inv, ev.ln_evidence_inv_var = 1.0, 2.0  # ratio == 1 -> -inf
    assert ev.compute_ln_inv_evidence_errors()[0] == -np.inf
    ev.ln_evidence_inv, ev.ln_evidence_inv_var = np.log(2.0), np.log(0.04)
    e, s = ev.compute_evidence()
    assert e == pytest.approx((1 + 0.04 / 4) / 2) and s == pytest.approx(0.2 / 4)
    lne, lns = ev.compute_ln_evidence()
    assert lne == pytest.approx(np.log(e)) and lns == pytest.approx(np.log(s))
    # the reference's overflow guard is isinf(nan_to_num(x, nan=inf)): it fires on NaN only
    # (nan_to_num maps +inf to the largest float), harmonic/evidence.py:423-428 -- mirrored as is
    ev.ln_evidence_inv = np.nan
    with pytest.raises(ValueError):
        ev.compute_evidence()
    a, b = hm.Evidence(3, m), hm.Evidence(3, m)
    with pytest.raises(ValueError):
        hm.evidence.compute_bayes_factor(a, b)
    for e_ in (a, b):
        e_.chains_added = True
        e_.nsamples_eff_per_chain = np.ones(3) * 100
    a.ln_evidence_inv, a.ln_evidence_inv_var = np.log(2.0), np.log(0.01)
    b.ln_evidence_inv, b.ln_evidence_inv_var = np.log(4.0), np.log(0.02)
    bf, bfs = hm.evidence.compute_bayes_factor(a, b)
    lbf, lbfs = hm.evidence.compute_ln_bayes_factor(a, b)
    assert bf == pytest.approx(4 / 2 * (1 + 0.01 / 4)) and lbf == pytest.approx(np.log(bf))
    assert lbfs == pytest.approx(np.log(bfs))


def test_evidence_pickle_drops_handles(tmp_path):
    m = hm.model.RealNVPModel(2)
    m.load_params(export.random_realnvp_params(2))
    ev = hm.Evidence(3, m, hm.Shifting.MAX_SHIFT)
    ev._partials = np.arange(12.0).reshape(3, 4)
    fn = os.path.join(tmp_path, "ev.pkl")
    ev.serialize(fn)
    ev2 = hm.Evidence.deserialize(fn)
    assert ev2._acc is None and ev2.shift == hm.Shifting.MAX_SHIFT
    np.testing.assert_array_equal(ev2._partials, ev._partials)


def test_split_data_views_and_guards():
    """harmonic/utils.py:129-200 and tests/test_utils.py:8-120: first chains train, rest test; argument checks;
    here both halves are views of the parent's arrays."""
    import harmonic_b200 as hm
    ndim, nsamples, nchains = 5, 100, 200
    rng = np.random.default_rng(3)
    samples = rng.standard_normal((nchains, nsamples, ndim))
    lnp = -0.5 * np.sum(samples * samples, axis=2)
    ch = hm.Chains(ndim)
    ch.add_chains_3d(samples, lnp)
    for prop in (0.5, 0.75):
        tr, te = hm.utils.split_data(ch, training_proportion=prop)
        ntr = int(nchains * prop)
        assert tr.nchains == ntr and te.nchains == nchains - ntr
        assert tr.nsamples == nsamples * ntr and te.nsamples == nsamples * (nchains - ntr)
        assert tr.start_indices == [i * nsamples for i in range(ntr + 1)]
        assert te.start_indices == [i * nsamples for i in range(nchains - ntr + 1)]
        assert tr.samples.shape == (nsamples * ntr, ndim) and te.ln_posterior.shape == (nsamples * (nchains - ntr),)
        i = int(rng.integers(nsamples * ntr))
        assert tr.samples[i, 3] == samples[i // nsamples, i % nsamples, 3]
        j = int(rng.integers(nsamples * (nchains - ntr)))
        assert te.samples[j, 3] == samples[j // nsamples + ntr, j % nsamples, 3]
        assert te.ln_posterior[j] == lnp[j // nsamples + ntr, j % nsamples]
        assert np.shares_memory(tr.samples, ch.samples) and np.shares_memory(te.samples, ch.samples)
    for bad in (-0.1, 0.0, 1.0, 1.5, 1e-10):
        with pytest.raises(ValueError):
            hm.utils.split_data(ch, training_proportion=bad)
    # ragged chains keep their own lengths
    rg = hm.Chains(2)
    for n in (3, 7, 5, 2):
        rg.add_chain(rng.standard_normal((n, 2)), rng.standard_normal(n))
    a, b = hm.utils.split_data(rg, 0.5)
    assert a.start_indices == [0, 3, 10] and b.start_indices == [0, 5, 7]
    np.testing.assert_array_equal(b.samples, rg.samples[10:])


def test_realnvp_ndim_limit_is_documented_in_the_header():
    """The generic RealNVP kernel keeps its activations in shared memory: flows wider than ndim 163 are rejected when
    the handle is created (HB_ERR_UNSUPPORTED -> NotImplementedError), which the header says."""
    hdr = open(os.path.join(ROOT, "include", "harmonic_b200.h")).read()
    assert "ndim <= 163" in hdr
