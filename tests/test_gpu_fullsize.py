"""Large-size GPU checks at (or near) the BASELINE.json configurations through size-independent
properties: streaming additivity, per-chain sums against float64 references on the device, NaN
counting, path agreement (tcgen05 vs CUDA-core kernels)."""
import math

import numpy as np
import pytest

import bench_workloads as wl
import harmonic_b200 as hm
from harmonic_b200 import _lib
from tests import torch_ref, util

pytestmark = pytest.mark.gpu


class Pre:
    ndim = 1
    flow = object()

    def is_fitted(self):
        return True


def test_cfg5_full_size_accumulation():
    """cfg5 at full size: 1e9 samples x 1e4 chains (8 GB device-resident), 0.01 % non-finite."""
    import torch
    nch, per = 10000, 100000
    lnphi, lnp = wl.device_data("cfg5", nch, per, torch.device("cuda", 0), seed=0)
    start = np.arange(nch + 1, dtype=np.int64) * per
    ev = hm.Evidence(nch, Pre())
    ev.add_lnpred(lnphi, lnp, start)
    # float64 reference on the device, chunked
    n = nch * per
    m, s, nbad = -math.inf, 0.0, 0
    for a in range(0, n, 1 << 27):
        la = (lnphi[a:a + (1 << 27)].double() - lnp[a:a + (1 << 27)].double())
        fin = torch.isfinite(la)
        nbad += int((~fin).sum())
        la = la[fin]
        mm = float(la.max())
        ss = float(torch.exp(la - mm).sum())
        M = max(m, mm)
        s = s * math.exp(m - M) + ss * math.exp(mm - M) if s > 0 else ss * math.exp(mm - M)
        m = M
    want = m + math.log(s) - math.log(n)
    assert ev.ln_evidence_inv == pytest.approx(want, abs=1e-6)
    assert ev.nsamples_per_chain.sum() == n
    assert ev.nsamples_per_chain.sum() - ev.nsamples_eff_per_chain.sum() == nbad and nbad > 0
    part = ev.partials()
    rng = np.random.default_rng(0)
    for c in rng.integers(0, nch, 12):
        seg = (lnphi[c * per:(c + 1) * per].double() - lnp[c * per:(c + 1) * per].double()).cpu().numpy()
        seg = seg[np.isfinite(seg)]
        ref = np.log(np.sum(np.exp(seg - seg.max()))) + seg.max()
        assert part[c, 0] + np.log(part[c, 1]) == pytest.approx(ref, abs=2e-6)
    # streaming: the two halves of every chain in two calls give the same statistics
    ev2 = hm.Evidence(nch, Pre())
    v1, v2 = lnphi.view(nch, per), lnp.view(nch, per)
    h = per // 2
    ev2.add_lnpred(v1[:, :h].contiguous().view(-1), v2[:, :h].contiguous().view(-1), np.arange(nch + 1) * h)
    ev2.add_lnpred(v1[:, h:].contiguous().view(-1), v2[:, h:].contiguous().view(-1), np.arange(nch + 1) * (per - h))
    assert ev2.ln_evidence_inv == pytest.approx(ev.ln_evidence_inv, abs=1e-6)
    assert ev2.ln_evidence_inv_var == pytest.approx(ev.ln_evidence_inv_var, abs=1e-4)


def test_cfg4_large_tcgen05_vs_float64_device_reference():
    """cfg4 shape at 1e7 samples: every per-sample ln phi of the tcgen05 path against a float64
    torch restatement (itself checked against the numpy oracle), then ln Z / errors / path agreement."""
    import torch
    model = wl.make_model("cfg4")
    nch, per = 100, 100000
    x, lnp = wl.device_data("cfg4", nch, per, torch.device("cuda", 0), seed=1)
    start = [i * per for i in range(nch + 1)]
    # the torch restatement agrees with the numpy oracle
    xs = x[:3000].cpu().numpy()
    ref_small = torch_ref.realnvp_log_prob_torch(model, x[:3000]).cpu().numpy()
    np.testing.assert_allclose(ref_small, util.oracle_predict(model, xs), rtol=1e-11, atol=1e-9)
    ref = torch_ref.realnvp_log_prob_torch(model, x)
    model.kernel_path = _lib.HB_PATH_TCGEN05
    got = model.predict(x).double()
    rel = ((got - ref).abs() / ref.abs().clamp(min=1.0))
    assert float(rel.max()) <= 1e-5, float(rel.max())
    ev = hm.Evidence(nch, model)
    ev.add_chains(hm.Chains.from_arrays(x, lnp, start))
    # float64 reference statistics from the reference ln phi
    lnarg = ref - lnp.double()
    want, _ = torch_ref.evidence_from_lnarg(lnarg, start)
    assert ev.ln_evidence_inv == pytest.approx(want, abs=1e-4)
    from oracle import evidence as oe
    o = oe.EvidenceOracle(nch)
    o.add(ref.cpu().numpy(), lnp.double().cpu().numpy(), start)
    util.assert_evidence_close(ev, o)
    # CUDA-core path on the same data
    m2 = wl.make_model("cfg4")
    m2.kernel_path = _lib.HB_PATH_SIMT
    sub = hm.Evidence(10, m2)
    sub.add_chains(hm.Chains.from_arrays(x[: 10 * per], lnp[: 10 * per], start[:11]))
    sub_tc = hm.Evidence(10, model)
    sub_tc.add_chains(hm.Chains.from_arrays(x[: 10 * per], lnp[: 10 * per], start[:11]))
    assert sub.ln_evidence_inv == pytest.approx(sub_tc.ln_evidence_inv, abs=1e-5)


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_cfg2_cfg3_full_size_streaming_and_subset_oracle(name):
    """cfg2 (RQSpline, 1e8 samples) / cfg3 (RealNVP 6+2, 1e7 samples) at full size: two streamed
    calls equal one call; a random subset of chains agrees with the float64 numpy oracle."""
    import torch
    model = wl.make_model(name)
    nch, per = wl.default_shape(name)
    x, lnp = wl.device_data(name, nch, per, torch.device("cuda", 0), seed=2)
    start = [i * per for i in range(nch + 1)]
    ev = hm.Evidence(nch, model)
    ev.add_chains(hm.Chains.from_arrays(x, lnp, start))
    h = per // 2
    xv, lv = x.view(nch, per, -1), lnp.view(nch, per)
    ev2 = hm.Evidence(nch, model)
    for sl in (slice(0, h), slice(h, per)):
        xs = xv[:, sl].contiguous().view(-1, x.shape[1])
        ls = lv[:, sl].contiguous().view(-1)
        k = xs.shape[0] // nch
        ev2.add_chains(hm.Chains.from_arrays(xs, ls, [i * k for i in range(nch + 1)]))
    assert ev2.ln_evidence_inv == pytest.approx(ev.ln_evidence_inv, abs=1e-6)
    assert ev2.ln_evidence_inv_var == pytest.approx(ev.ln_evidence_inv_var, abs=1e-4)
    part = ev.partials()
    rng = np.random.default_rng(1)
    for c in rng.integers(0, nch, 3):
        k = min(per, 20000)
        xs = x[c * per:c * per + k].cpu().numpy()
        got = model.predict(x[c * per:c * per + k]).cpu().numpy()
        util.check_predict(model, xs, got)
    assert part[:, 2].sum() == nch * per


def test_host_staging_multi_chunk_matches_device_resident():
    """Host (float64, pageable) chains larger than one 256 MB staging chunk: the double-buffered
    H2D pipeline must give the same statistics as the device-resident path, and host predict must
    chunk correctly above 2^20 rows."""
    import torch
    model = wl.make_model("cfg4")
    nch, per = 60, 20000  # 1.2e6 rows x 64 float64 = 614 MB -> 3 staging chunks
    x, lnp = wl.device_data("cfg4", nch, per, torch.device("cuda", 0), seed=3)
    start = [i * per for i in range(nch + 1)]
    xh = x.double().cpu().numpy()
    lh = lnp.double().cpu().numpy()
    ch = hm.Chains.from_arrays(xh, lh, start)
    ev_h = hm.Evidence(nch, model)
    ev_h.add_chains(ch)
    ev_d = hm.Evidence(nch, model)
    ev_d.add_chains(hm.Chains.from_arrays(x, lnp, start))
    assert ev_h.ln_evidence_inv == pytest.approx(ev_d.ln_evidence_inv, abs=1e-6)
    assert ev_h.ln_evidence_inv_var == pytest.approx(ev_d.ln_evidence_inv_var, abs=1e-4)
    np.testing.assert_array_equal(ev_h.nsamples_per_chain, ev_d.nsamples_per_chain)
    p_h = model.predict(xh)  # > 2^20 rows -> two predict chunks
    p_d = model.predict(x).cpu().numpy()
    assert p_h.shape == (nch * per,)
    np.testing.assert_allclose(p_h, p_d, rtol=0, atol=0)
