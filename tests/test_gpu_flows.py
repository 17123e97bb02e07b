"""GPU parity: flow predict and the fused predict+accumulate path vs the float64 oracle.

Tolerances are the north_star's: per-sample ln phi 1e-5 relative, ln Z 1e-4 absolute, error
estimates 1e-3 relative."""
import numpy as np
import pytest

import harmonic_b200 as hm
from harmonic_b200 import _lib
from tests import util

pytestmark = pytest.mark.gpu


def samples(n, ndim, seed=0, scale=1.0):
    return np.random.default_rng(seed).standard_normal((n, ndim)) * scale


@pytest.mark.parametrize("ndim", [2, 3, 4, 5, 6, 7, 8, 17, 64])
@pytest.mark.parametrize("standardize", [False, True])
@pytest.mark.parametrize("path", [_lib.HB_PATH_SIMT, _lib.HB_PATH_SIMT_GENERIC])
def test_realnvp_predict_simt(ndim, standardize, path):
    """HB_PATH_SIMT uses the register kernel for ndim <= 8 and the generic kernel above;
    HB_PATH_SIMT_GENERIC forces the generic (shared-memory) kernel."""
    if path == _lib.HB_PATH_SIMT_GENERIC and ndim > 8:
        pytest.skip("same kernel as HB_PATH_SIMT")
    m = util.make_realnvp(ndim, seed=ndim, standardize=standardize, temperature=0.8)
    m.kernel_path = path
    x = samples(1000, ndim, seed=1, scale=1.5)
    got = m.predict(x)
    assert got.dtype == np.float32 and got.shape == (1000,)
    util.check_predict(m, x, got)
    # float32 input gives the same answer; 1-D input gives a scalar (model.py:229-231)
    util.check_predict(m, x.astype(np.float32), m.predict(x.astype(np.float32)))
    assert np.ndim(m.predict(x[0])) == 0
    assert float(m.predict(x[0])) == pytest.approx(float(got[0]), rel=1e-6)


@pytest.mark.parametrize("layers", [(1, 0), (6, 2), (3, 5)])
def test_realnvp_layer_counts_and_temperature(layers):
    m = util.make_realnvp(5, layers[0], layers[1], seed=7, ws=0.4)
    m.kernel_path = _lib.HB_PATH_SIMT
    x = samples(777, 5, seed=2)
    for T in (1.0, 0.9, 0.3):
        m.temperature = T  # runtime-mutable, read on every call (model.py:214)
        util.check_predict(m, x, m.predict(x))
    m.temperature = 1.5
    with pytest.raises(ValueError):
        m.predict(x)
    m.temperature = 0.0
    with pytest.raises(ValueError):
        m.predict(x)


@pytest.mark.parametrize("ndim,hidden,layers", [(2, (32, 32), 3), (2, (64, 64), 8), (4, (64, 64), 8),
                                                 (5, (16,), 2), (7, (8, 24, 16), 3), (3, (), 2),
                                                 (3, (64, 64), 4), (5, (48, 16), 3), (8, (64, 64), 2),
                                                 (9, (32, 32), 2)])
@pytest.mark.parametrize("standardize", [False, True])
@pytest.mark.parametrize("path", [_lib.HB_PATH_AUTO, _lib.HB_PATH_SIMT, _lib.HB_PATH_SIMT_GENERIC])
def test_rqspline_predict(ndim, hidden, layers, standardize, path):
    """AUTO uses the tcgen05 kernel at ndim 2 (two hidden layers, 32 / 64 wide, 8 bins), the register fast path for
    two hidden layers / ndim <= 8 / 8 bins and the generic kernel otherwise; HB_PATH_SIMT keeps the CUDA-core kernels,
    HB_PATH_SIMT_GENERIC forces the generic one."""
    if path == _lib.HB_PATH_SIMT_GENERIC and len(hidden) != 2:
        pytest.skip("generic kernel is already the AUTO choice")
    if path == _lib.HB_PATH_SIMT and ndim != 2:
        pytest.skip("AUTO is already the CUDA-core choice")
    if path == _lib.HB_PATH_AUTO and ndim == 2:
        pytest.skip("AUTO is the tensor path here: test_rqspline_tcgen05_*")
    m = util.make_rqspline(ndim, layers, 8, hidden, seed=ndim + layers, standardize=standardize)
    m.kernel_path = path
    x = samples(1500, ndim, seed=3, scale=2.0)
    x[:20] *= 8.0  # some features outside the spline range -> linear tails
    x[5, 0] = 10.0
    x[6, ndim - 1] = -10.0
    got = m.predict(x)
    util.check_predict(m, x, got)


def test_rqspline_bins_and_range():
    m = util.make_rqspline(3, 2, 5, (16, 16), seed=11, spline_range=(-3.0, 4.0), temperature=0.5)
    x = samples(800, 3, seed=4, scale=2.5)
    util.check_predict(m, x, m.predict(x))


@pytest.mark.parametrize("kind", ["realnvp", "rqspline"])
def test_nan_propagates(kind):
    m = util.make_realnvp(4, seed=1) if kind == "realnvp" else util.make_rqspline(4, seed=1)
    x = samples(300, 4, seed=5)
    x[3, 0] = np.nan
    x[7, 3] = np.nan
    x[9, 1] = np.inf
    got = m.predict(x)
    want = util.oracle_predict(m, x)
    assert np.isnan(got[3]) and np.isnan(got[7])
    util.assert_lnphi_close(np.delete(got, 9), np.delete(want, 9))
    assert not np.isfinite(got[9]) and not np.isfinite(want[9])  # an infinite feature never gives a finite ln phi


@pytest.mark.parametrize("kind", ["realnvp", "rqspline"])
@pytest.mark.parametrize("shift", list(hm.Shifting))
def test_fused_add_chains_vs_oracle(kind, shift):
    ndim = 5
    m = (util.make_realnvp(ndim, seed=2, standardize=True, ws=0.3) if kind == "realnvp"
         else util.make_rqspline(ndim, 3, 8, (32, 32), seed=2, standardize=True, ws=0.3))
    if kind == "realnvp":
        m.kernel_path = _lib.HB_PATH_SIMT
    ch = util.gaussian_chains(ndim, 30, 700, seed=9, ragged=True)
    ev = hm.Evidence(ch.nchains, m, shift)
    ev.add_chains(ch)
    ref = util.oracle_evidence(m, ch, shift.value)
    util.assert_evidence_close(ev, ref)
    assert ev.shift_set and ev.chains_added
    assert ev.shift_value == pytest.approx(ref.shift_value, rel=1e-5, abs=1e-5)
    assert ev.lnargmax == pytest.approx(ref.lnargmax, rel=1e-5, abs=2e-5)
    assert ev.lnpredictmin == pytest.approx(ref.lnpredictmin, rel=1e-5)
    assert ev.lnprobmax == pytest.approx(ref.lnprobmax, rel=1e-6)


def test_fused_streaming_two_calls_and_nan_sample():
    ndim = 3
    m = util.make_rqspline(ndim, 2, 8, (16, 16), seed=4, ws=0.3)
    c1 = util.gaussian_chains(ndim, 12, 300, seed=1)
    c2 = util.gaussian_chains(ndim, 12, 200, seed=2)
    c2.samples[17, 1] = np.nan  # NaN sample: effective count drops, raw count does not (test_evidence.py:478-516)
    ev = hm.Evidence(12, m)
    ev.add_chains(c1, num_slices=2)
    ev.add_chains(c2, num_slices=10)
    ref = util.oracle_evidence(m, None, calls=[c1, c2])
    util.assert_evidence_close(ev, ref)
    assert ev.nsamples_eff_per_chain.sum() == 12 * 500 - 1
    with pytest.raises(ValueError):
        ev.add_chains(c1, num_slices=10**9)


def test_fused_device_resident_f32_matches_host():
    import torch
    ndim = 6
    m = util.make_realnvp(ndim, seed=3, ws=0.3)
    m.kernel_path = _lib.HB_PATH_SIMT
    ch = util.gaussian_chains(ndim, 16, 1000, seed=4)
    a = hm.Evidence(16, m)
    a.add_chains(ch)
    xs = torch.from_numpy(ch.samples.astype(np.float32)).cuda()
    ys = torch.from_numpy(ch.ln_posterior.astype(np.float32)).cuda()
    dch = hm.Chains.from_arrays(xs, ys, ch.start_indices)
    b = hm.Evidence(16, m)
    b.add_chains(dch)
    assert a.ln_evidence_inv == pytest.approx(b.ln_evidence_inv, abs=1e-6)
    p = m.predict(xs)
    assert p.is_cuda and p.shape == (16000,)
    util.assert_lnphi_close(p.cpu().numpy(), util.oracle_predict(m, ch.samples.astype(np.float32)))


# ------------------------------------------------------------------ tcgen05 path (ndim = 64)
@pytest.mark.parametrize("standardize", [False, True])
@pytest.mark.parametrize("n", [1, 127, 128, 1000, 70000])
def test_realnvp_tcgen05_predict(standardize, n):
    m = util.make_realnvp(64, seed=5, standardize=standardize, ws=0.5)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = samples(n, 64, seed=6, scale=1.2)
    got = m.predict(x)
    util.check_predict(m, x, got)
    m2 = util.make_realnvp(64, seed=5, standardize=standardize, ws=0.5)
    m2.kernel_path = _lib.HB_PATH_SIMT
    util.assert_lnphi_close(got, m2.predict(x).astype(np.float64), rel=2e-6)


@pytest.mark.parametrize("layers", [(1, 0), (2, 4), (3, 1), (4, 4), (8, 0), (1, 7)])
def test_realnvp_tcgen05_layers_nan(layers):
    m = util.make_realnvp(64, layers[0], layers[1], seed=8, ws=0.5, temperature=0.6)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = samples(3000, 64, seed=7)
    x[11, 3] = np.nan
    x[12, 40] = np.nan
    got = m.predict(x.astype(np.float32))
    util.check_predict(m, x.astype(np.float32), got)
    assert np.isnan(got[11]) and np.isnan(got[12])


@pytest.mark.parametrize("scale,trips", [(2.0, False), (5.0, None), (30.0, True), (300.0, True), (3e4, True)])
def test_realnvp_tcgen05_large_magnitude(scale, trips):
    """Inputs far outside the trained range: the scaled layers hit the 1e-3 clip, which multiplies every
    float32 rounding error by up to 1e3 per layer, the tensor cores truncate their float32 accumulator on every MMA
    step (profiles/r1_tensor_accum_rounding.txt) and beyond |hidden| = 2047 the FP16 operands of GEMM2 would overflow.
    HB_PATH_AUTO therefore guards the tensor path: rows with a standardised input beyond 32 (or a hidden unit that
    could leave the half range) are counted by the kernel and the call is evaluated again on the CUDA cores, so AUTO
    meets the SAME bound as the float32 path at every magnitude (no extra slack) and equals it bit for bit whenever
    the guard trips.  At 2x the trained range nothing trips and the tensor path itself meets the bound; at 5x a few
    rows whose transformed features were blown up by a clipped scale (1 / 1e-3) trip it."""
    m = util.make_realnvp(64, seed=5, standardize=True, ws=0.3)
    assert m.kernel_path == _lib.HB_PATH_AUTO
    lib = _lib.load()
    assert lib.hb_flow_uses_tcgen05(m._get_handle()) == 1
    x = (samples(4000, 64, seed=16) * scale).astype(np.float32)
    rows, calls = _lib.C.c_int64(0), _lib.C.c_int64(0)
    got = m.predict(x)
    _lib.check(lib.hb_flow_guard_stats(m._get_handle(), _lib.C.byref(rows), _lib.C.byref(calls)))
    if trips is not None:
        assert (calls.value == 1 and rows.value > 0) if trips else (calls.value == 0 and rows.value == 0)
    trips = calls.value > 0
    assert np.isfinite(got).all()
    m2 = util.make_realnvp(64, seed=5, standardize=True, ws=0.3)
    m2.kernel_path = _lib.HB_PATH_SIMT
    simt = m2.predict(x)
    if trips:
        np.testing.assert_array_equal(got, simt)
    if scale <= 30.0:
        util.check_predict(m, x, got)  # slack_mult = 1
    elif scale <= 300.0:
        util.check_predict(m, x, got, slack_mult=2.0)  # the float32 path's own bound at this magnitude
    # the device-resident path takes the same route
    import torch
    got_dev = m.predict(torch.from_numpy(x).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got_dev, got)


def test_realnvp_tcgen05_guard_in_add_chains(caplog):
    """One chain of a set has samples tens of sigmas out: under AUTO the whole add_chains call is folded again on the
    CUDA cores (state restored from the copy taken at the start of the call), a second, in-range call then stays on
    the tensor path; the result equals the float32 path's and matches the oracle."""
    import logging
    m = util.make_realnvp(64, seed=9, standardize=True, ws=0.3)
    ch_ok = util.gaussian_chains(64, 12, 700, seed=3, ragged=True)
    ch_far = util.gaussian_chains(64, 12, 500, seed=4, ragged=True)
    ch_far.samples[37, 5] = 150.0
    ev = hm.Evidence(12, m)
    with caplog.at_level(logging.WARNING, logger="Harmonic"):
        ev.add_chains(ch_far)
    assert ev.guard_trips == 1 and any("CUDA cores" in r.message for r in caplog.records)
    ev.add_chains(ch_ok)
    assert ev.guard_trips == 1
    m2 = util.make_realnvp(64, seed=9, standardize=True, ws=0.3)
    m2.kernel_path = _lib.HB_PATH_SIMT
    ev2 = hm.Evidence(12, m2)
    ev2.add_chains(ch_far)
    ev2.add_chains(ch_ok)
    assert abs(ev.ln_evidence_inv - ev2.ln_evidence_inv) < 1e-6
    ref = util.oracle_evidence(m, ch_far)
    ev3 = hm.Evidence(12, m)
    ev3.add_chains(ch_far)
    util.assert_evidence_close(ev3, ref)


def test_realnvp_tcgen05_fused_vs_oracle_and_auto_path():
    m = util.make_realnvp(64, seed=9, standardize=True, ws=0.3)
    lib = _lib.load()
    assert lib.hb_flow_uses_tcgen05(m._get_handle()) == 1  # AUTO picks the tensor path at ndim 64
    ch = util.gaussian_chains(64, 24, 900, seed=3, ragged=True)
    ev = hm.Evidence(ch.nchains, m)
    ev.add_chains(ch)
    ref = util.oracle_evidence(m, ch)
    util.assert_evidence_close(ev, ref)


# ------------------------------------------------------------------ FlowModel.sample (harmonic/model.py:239-275)
@pytest.mark.parametrize("ndim", [2, 5, 17, 64])
@pytest.mark.parametrize("standardize", [False, True])
def test_realnvp_sample_vs_oracle(ndim, standardize):
    m = util.make_realnvp(ndim, seed=12, standardize=standardize, ws=0.4, temperature=0.8)
    eps = samples(3000, ndim, seed=13)
    got = m.sample(3000, eps=eps)
    assert got.dtype == np.float32 and got.shape == (3000, ndim)
    util.check_sample(m, eps, got)
    # float32 draws give the same samples as float64 draws rounded to float32
    got32 = m.sample(3000, eps=eps.astype(np.float32))
    util.check_sample(m, eps.astype(np.float32), got32)


@pytest.mark.parametrize("ndim,hidden,layers", [(2, (64, 64), 8), (3, (16, 16), 3), (5, (32,), 2), (6, (8, 8, 8), 2), (4, (), 3)])
@pytest.mark.parametrize("standardize", [False, True])
def test_rqspline_sample_vs_oracle(ndim, hidden, layers, standardize):
    m = util.make_rqspline(ndim, layers, 8, hidden, seed=14, standardize=standardize, ws=0.4, temperature=0.9)
    eps = samples(2000, ndim, seed=15, scale=1.5)
    eps[0, :] = 40.0   # linear tails of the spline
    eps[1, :] = -40.0
    got = m.sample(2000, eps=eps)
    util.check_sample(m, eps, got)


@pytest.mark.parametrize("kind", ["realnvp", "rqspline"])
def test_sample_round_trip_and_device_path(kind):
    """predict(sample(eps)) against the change-of-variables value is covered by the oracle tests; here: the device
    path (CUDA tensor in -> CUDA tensor out) equals the host path, NaN draws give NaN rows, and ln phi of the
    samples equals the oracle's ln phi at the same points (the two directions share one weight pack)."""
    import torch
    ndim = 5
    m = (util.make_realnvp(ndim, seed=21, standardize=True, ws=0.4) if kind == "realnvp"
         else util.make_rqspline(ndim, 3, 8, (16, 16), seed=21, standardize=True, ws=0.4))
    eps = samples(4096 + 37, ndim, seed=22).astype(np.float32)
    eps[5, 2] = np.nan
    host = m.sample(eps.shape[0], eps=eps)
    dev = m.sample(eps.shape[0], eps=torch.from_numpy(eps).cuda())
    assert dev.is_cuda
    np.testing.assert_array_equal(np.isnan(host), np.isnan(dev.cpu().numpy()))
    ok = ~np.isnan(host)
    np.testing.assert_array_equal(host[ok], dev.cpu().numpy()[ok])
    assert np.isnan(host[5]).any()
    good = np.delete(host, 5, axis=0)
    util.check_predict(m, good, m.predict(good))


def test_sample_rng_moments_and_guards():
    """tests/test_flow_model.py:209-251: samples of a flow trained on a Gaussian reproduce its moments; a lower
    temperature concentrates them.  Uses the committed cfg1 fixture (trained on N(0, Sigma), ndim 5)."""
    from bench_workloads import init_cov, load_cfg1_model
    m = load_cfg1_model(1.0)
    x = m.sample(200000, rng_key=10)
    assert x.shape == (200000, 5) and x.dtype == np.float32
    cov = init_cov(5)
    assert np.all(np.abs(x.mean(axis=0)) < 0.15)
    assert np.all(np.abs(x.var(axis=0) - np.diag(cov)) < 0.15)
    m.temperature = 0.8
    xc = m.sample(200000, rng_key=10)
    assert np.all(xc.var(axis=0) < x.var(axis=0))
    np.testing.assert_array_equal(m.sample(100, rng_key=3), m.sample(100, rng_key=3))
    assert not np.array_equal(m.sample(100, rng_key=3), m.sample(100, rng_key=4))
    m.temperature = 1.5
    with pytest.raises(ValueError):
        m.sample(10)
    m.temperature = 0.8
    with pytest.raises(ValueError):
        m.sample(10, eps=np.zeros((10, 4)))


# ------------------------------------------------------------------ tcgen05 path: shapes at the edges of support
@pytest.mark.parametrize("layers", [(4, 4), (2, 6), (8, 0), (1, 7)])
def test_realnvp_tcgen05_layer_limits(layers):
    """Up to 8 coupling layers run on the tensor path (7 frozen columns + the bias column fill one K = 8 step),
    all-scaled and mostly-unscaled stacks; 9 layers fall back to the CUDA-core kernels."""
    m = util.make_realnvp(64, layers[0], layers[1], seed=31, standardize=True, ws=0.3, temperature=0.7)
    lib = _lib.load()
    assert lib.hb_flow_uses_tcgen05(m._get_handle()) == 1
    x = samples(2500, 64, seed=32)
    util.check_predict(m, x, m.predict(x))


def test_realnvp_nine_layers_leave_the_tensor_path():
    m = util.make_realnvp(64, 3, 6, seed=33, ws=0.3)
    lib = _lib.load()
    assert lib.hb_flow_uses_tcgen05(m._get_handle()) == 0
    x = samples(700, 64, seed=34)
    util.check_predict(m, x, m.predict(x))
    m.kernel_path = _lib.HB_PATH_TCGEN05
    with pytest.raises(NotImplementedError):
        m.predict(x)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_realnvp_tcgen05_device_strided_and_f64(dtype):
    """Device-resident samples with a row stride larger than ndim (ld = 80) and float64 storage: both are read in
    place by the tensor-path kernel (per-row loads; float64 is demoted on load as JAX does)."""
    import torch
    m = util.make_realnvp(64, seed=35, standardize=True, ws=0.3)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    n = 3000
    big = torch.from_numpy(samples(n, 80, seed=36)).to(getattr(torch, dtype)).cuda()
    view = big[:, 8:72]
    assert view.stride(0) == 80
    got = m.predict(view).cpu().numpy()
    util.check_predict(m, view.cpu().numpy().astype(np.float32 if dtype == "float32" else np.float64), got)
    # the same rows through the fused path, ragged chains
    lnp = torch.from_numpy(np.random.default_rng(1).standard_normal(n)).to(getattr(torch, dtype)).cuda()
    start = [0, 5, 900, 1300, 2100, n]
    ev = hm.Evidence(5, m)
    ev.add_chains(hm.Chains.from_arrays(view, lnp, start))
    host = hm.Chains.from_arrays(view.cpu().numpy(), lnp.cpu().numpy(), start)
    ref = util.oracle_evidence(m, host)
    util.assert_evidence_close(ev, ref)


# ------------------------------------------------------------------ tcgen05 path: the two-tile layout kept for A/B
_TWO_TILE_SCRIPT = r"""
import numpy as np
import harmonic_b200 as hm
from harmonic_b200 import _lib
from tests import util
for layers in ((2, 4), (8, 0), (1, 7)):
    m = util.make_realnvp(64, layers[0], layers[1], seed=5, standardize=True, ws=0.5)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = np.random.default_rng(6).standard_normal((5000, 64)) * 1.2
    util.check_predict(m, x, m.predict(x))
m = util.make_realnvp(64, seed=9, standardize=True, ws=0.3)
ch = util.gaussian_chains(64, 24, 900, seed=3, ragged=True)
ev = hm.Evidence(ch.nchains, m)
ev.add_chains(ch)
util.assert_evidence_close(ev, util.oracle_evidence(m, ch))
print("two-tile ok")
"""


def test_realnvp_tcgen05_two_tile_layout():
    """HB_TC_TILES=2 (read once per process) selects the two-tile layout of realnvp_tc_kernel, kept for A/B
    measurements against the default three-tile layout: it must stay parity-green too."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HB_TC_TILES="2")
    out = subprocess.run([sys.executable, "-c", _TWO_TILE_SCRIPT], cwd=root, env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and "two-tile ok" in out.stdout, out.stdout + out.stderr


# ------------------------------------------------------------------ RQSpline on the tensor cores (ndim 2)
@pytest.mark.parametrize("hidden,layers", [((64, 64), 8), ((32, 32), 3), ((64, 16), 5), ((32, 64), 1)])
@pytest.mark.parametrize("standardize", [False, True])
@pytest.mark.parametrize("n", [1, 127, 128, 1000, 70000])
def test_rqspline_tcgen05_predict(hidden, layers, standardize, n):
    """rqspline_tc_kernel (Dense(h1 -> 25) contraction as FP16-split tcgen05 MMAs) against the float64 oracle, at row
    counts around the 128-sample tile and with samples in the spline's linear tails."""
    m = util.make_rqspline(2, layers, 8, hidden, seed=layers + hidden[0], standardize=standardize, ws=0.3)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = samples(n, 2, seed=13, scale=2.0)
    x[: n // 50] *= 8.0
    got = m.predict(x)
    assert _lib.load().hb_flow_uses_tcgen05(m._get_handle(0)) == 1
    util.check_predict(m, x, got)
    m2 = util.make_rqspline(2, layers, 8, hidden, seed=layers + hidden[0], standardize=standardize, ws=0.3)
    m2.kernel_path = _lib.HB_PATH_SIMT
    np.testing.assert_allclose(got, m2.predict(x), rtol=2e-5, atol=2e-5)


def test_rqspline_tcgen05_ill_conditioned_flow_is_float32_grade():
    """Weights of unit scale make the 8-layer flow ill-conditioned for ANY float32 evaluation (the float32 restatement of
    the oracle itself leaves 1e-5 on ~0.3 % of the samples).  The tensor path must then be as good as float32 arithmetic
    is: error statistics against the float64 oracle within a small factor of the CUDA-core kernel's and of the float32
    restatement's -- per-sample slack proportional to ONE float32 evaluation's error would reject any other
    float32-grade arithmetic on the samples where that evaluation happened to land close."""
    m = util.make_rqspline(2, 8, 8, (64, 64), seed=10, standardize=True)
    x = samples(40000, 2, seed=21, scale=2.0)
    want = util.oracle_predict(m, x, np.float64)
    den = np.maximum(np.abs(want), 1.0)
    e32 = np.abs(util.oracle_predict(m, x, np.float32).astype(np.float64) - want) / den
    errs = {}
    for name, path in (("tc", _lib.HB_PATH_TCGEN05), ("simt", _lib.HB_PATH_SIMT)):
        m.kernel_path = path
        errs[name] = np.abs(np.asarray(m.predict(x), dtype=np.float64) - want) / den
    ref_max = max(e32.max(), errs["simt"].max())
    ref_cnt = max(int((e32 > 1e-5).sum()), int((errs["simt"] > 1e-5).sum()))
    assert errs["tc"].max() <= 2.0 * ref_max
    assert int((errs["tc"] > 1e-5).sum()) <= 1.5 * ref_cnt + 10
    assert np.median(errs["tc"]) <= 1.5 * np.median(errs["simt"]) + 1e-8


def test_rqspline_tcgen05_auto_and_unsupported_shapes():
    lib = _lib.load()
    m = util.make_rqspline(2, 8, 8, (64, 64), seed=2)
    m.predict(samples(10, 2, seed=1))
    assert lib.hb_flow_uses_tcgen05(m._get_handle(0)) == 1  # AUTO
    for bad in (util.make_rqspline(3, 2, 8, (64, 64), seed=2), util.make_rqspline(2, 2, 8, (48, 16), seed=2),
                util.make_rqspline(2, 2, 5, (64, 64), seed=2), util.make_rqspline(2, 2, 8, (64,), seed=2)):
        bad.predict(samples(10, bad.ndim, seed=1))
        assert lib.hb_flow_uses_tcgen05(bad._get_handle(0)) == 0
        bad2 = util.make_rqspline(bad.ndim, 2, 8, (16, 16), seed=3)
        bad2.kernel_path = _lib.HB_PATH_TCGEN05
        with pytest.raises(Exception):
            bad2.predict(samples(10, bad.ndim, seed=1))


def test_rqspline_tcgen05_nan_and_inf_rows():
    m = util.make_rqspline(2, 8, 8, (64, 64), seed=6, standardize=True, ws=0.3)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = samples(700, 2, seed=5)
    x[3, 0] = np.nan
    x[200, 1] = np.nan
    x[9, 1] = np.inf
    x[130, 0] = -np.inf
    got = m.predict(x)
    want = util.oracle_predict(m, x)
    assert np.isnan(got[3]) and np.isnan(got[200])
    keep = np.ones(700, bool)
    keep[[9, 130]] = False
    util.assert_lnphi_close(got[keep], want[keep])
    assert not np.isfinite(got[9]) and not np.isfinite(got[130])


@pytest.mark.parametrize("shift", [hm.Shifting.MEAN_SHIFT, hm.Shifting.MAX_SHIFT])
def test_rqspline_tcgen05_fused_vs_oracle(shift):
    """add_chains through the tensor-path kernel: ragged chains, two streaming calls, a NaN sample; equals the oracle
    and the CUDA-core path."""
    m = util.make_rqspline(2, 8, 8, (64, 64), seed=7, standardize=True, ws=0.3)
    c1 = util.gaussian_chains(2, 14, 900, seed=1, ragged=True)
    c2 = util.gaussian_chains(2, 14, 300, seed=2)
    c2.samples[17, 1] = np.nan
    ev = hm.Evidence(14, m, shift)
    ev.add_chains(c1)
    ev.add_chains(c2, num_slices=3)
    ref = util.oracle_evidence(m, None, shift.value, calls=[c1, c2])
    util.assert_evidence_close(ev, ref)
    m2 = util.make_rqspline(2, 8, 8, (64, 64), seed=7, standardize=True, ws=0.3)
    m2.kernel_path = _lib.HB_PATH_SIMT
    ev2 = hm.Evidence(14, m2, shift)
    ev2.add_chains(c1)
    ev2.add_chains(c2, num_slices=3)
    assert abs(ev.ln_evidence_inv - ev2.ln_evidence_inv) < 1e-5


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_rqspline_tcgen05_device_strided_and_f64(dtype):
    import torch
    m = util.make_rqspline(2, 3, 8, (32, 32), seed=8, standardize=True, ws=0.3)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    n = 3000
    big = torch.from_numpy(samples(n, 6, seed=36, scale=1.5)).to(getattr(torch, dtype)).cuda()
    view = big[:, 2:4]
    assert view.stride(0) == 6
    got = m.predict(view).cpu().numpy()
    util.check_predict(m, view.cpu().numpy().astype(np.float32 if dtype == "float32" else np.float64), got)
    lnp = torch.from_numpy(np.random.default_rng(1).standard_normal(n)).to(getattr(torch, dtype)).cuda()
    start = [0, 5, 900, 1300, 2100, n]
    ev = hm.Evidence(5, m)
    ev.add_chains(hm.Chains.from_arrays(view, lnp, start))
    host = hm.Chains.from_arrays(view.cpu().numpy(), lnp.cpu().numpy(), start)
    util.assert_evidence_close(ev, util.oracle_evidence(m, host))


# ------------------------------------------------------------------ small-ndim RealNVP on the tensor cores (ndim 2..8)
@pytest.mark.parametrize("ndim", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("standardize", [False, True])
@pytest.mark.parametrize("layers", [(2, 4), (6, 2), (1, 3)])
def test_realnvp_small_tcgen05_predict(ndim, standardize, layers):
    """realnvp_small_tc_kernel (relu(H) (0.99 W2) as FP16-split tcgen05 MMAs, H and the linear part on the CUDA cores)
    against the float64 oracle and the register kernel."""
    m = util.make_realnvp(ndim, layers[0], layers[1], seed=ndim + 3 * layers[0], standardize=standardize, ws=0.3)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = samples(3000, ndim, seed=7, scale=1.5)
    got = m.predict(x)
    assert _lib.load().hb_flow_uses_tcgen05(m._get_handle(0)) == 1
    util.check_predict(m, x, got)
    m2 = util.make_realnvp(ndim, layers[0], layers[1], seed=ndim + 3 * layers[0], standardize=standardize, ws=0.3)
    m2.kernel_path = _lib.HB_PATH_SIMT
    np.testing.assert_allclose(got, m2.predict(x), rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 50000])
def test_realnvp_small_tcgen05_row_counts_and_auto(n):
    m = util.make_realnvp(5, 6, 2, seed=3, standardize=True, ws=0.3)  # the cfg3 flow; AUTO = tensor path
    x = samples(n, 5, seed=2)
    got = m.predict(x)
    assert _lib.load().hb_flow_uses_tcgen05(m._get_handle(0)) == 1
    util.check_predict(m, x, got)


def test_realnvp_small_tcgen05_nan_inf_and_guard(caplog):
    """NaN rows propagate, an infinite feature never gives a finite ln phi; rows far outside the range where
    fp16(relu(h)) is finite trip the guard: AUTO evaluates the call again on the CUDA cores and says so, the forced
    tensor path does not."""
    import logging
    m = util.make_realnvp(5, 6, 2, seed=3, standardize=True, ws=0.3)
    x = samples(600, 5, seed=5)
    x[3, 0] = np.nan
    x[7, 4] = np.nan
    x[9, 1] = np.inf
    got = m.predict(x)
    want = util.oracle_predict(m, x)
    assert np.isnan(got[3]) and np.isnan(got[7]) and not np.isfinite(got[9])
    util.assert_lnphi_close(np.delete(got, 9), np.delete(want, 9))
    far = samples(600, 5, seed=6)
    far[11, 2] = 3.0e6
    with caplog.at_level(logging.WARNING, logger="Harmonic"):
        got = m.predict(far)
    assert any("CUDA cores" in r.message for r in caplog.records)
    m2 = util.make_realnvp(5, 6, 2, seed=3, standardize=True, ws=0.3)
    m2.kernel_path = _lib.HB_PATH_SIMT
    np.testing.assert_array_equal(got, m2.predict(far))  # the re-evaluation IS the CUDA-core path


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_realnvp_small_tcgen05_fused_vs_oracle(dtype):
    """add_chains through the tensor-path kernel: ragged chains, two streaming calls, a NaN sample, host float64 and
    device-resident strided inputs; equals the oracle and the register kernel."""
    import torch
    m = util.make_realnvp(5, 6, 2, seed=2, standardize=True, ws=0.3)
    c1 = util.gaussian_chains(5, 14, 900, seed=1, ragged=True)
    c2 = util.gaussian_chains(5, 14, 300, seed=2)
    c2.samples[17, 1] = np.nan
    ev = hm.Evidence(14, m)
    ev.add_chains(c1)
    ev.add_chains(c2, num_slices=3)
    ref = util.oracle_evidence(m, None, calls=[c1, c2])
    util.assert_evidence_close(ev, ref)
    m2 = util.make_realnvp(5, 6, 2, seed=2, standardize=True, ws=0.3)
    m2.kernel_path = _lib.HB_PATH_SIMT
    ev2 = hm.Evidence(14, m2)
    ev2.add_chains(c1)
    ev2.add_chains(c2, num_slices=3)
    assert abs(ev.ln_evidence_inv - ev2.ln_evidence_inv) < 1e-5
    n = 3000
    big = torch.from_numpy(samples(n, 9, seed=36)).to(getattr(torch, dtype)).cuda()
    view = big[:, 2:7]
    lnp = torch.from_numpy(np.random.default_rng(1).standard_normal(n)).to(getattr(torch, dtype)).cuda()
    start = [0, 5, 900, 1300, 2100, n]
    ev3 = hm.Evidence(5, m)
    ev3.add_chains(hm.Chains.from_arrays(view, lnp, start))
    host = hm.Chains.from_arrays(view.cpu().numpy(), lnp.cpu().numpy(), start)
    util.assert_evidence_close(ev3, util.oracle_evidence(m, host))


_STC_THRESHOLD_SCRIPT = r"""
import numpy as np
from harmonic_b200 import _lib
from tests import util
m = util.make_realnvp(5, 6, 2, seed=3, standardize=True, ws=0.3)
x = np.random.default_rng(2).standard_normal((4000, 5))
x[11, 2] = 3.0e6   # would trip the tensor path's guard
auto = m.predict(x)
m2 = util.make_realnvp(5, 6, 2, seed=3, standardize=True, ws=0.3)
m2.kernel_path = _lib.HB_PATH_SIMT
assert np.array_equal(auto, m2.predict(x))
import ctypes
rows, calls = ctypes.c_int64(0), ctypes.c_int64(0)
_lib.check(_lib.load().hb_flow_guard_stats(m._get_handle(0), ctypes.byref(rows), ctypes.byref(calls)))
assert calls.value == 0
print("threshold ok")
"""


def test_realnvp_small_tcgen05_default_row_threshold():
    """Without HB_STC_MIN_ROWS (the test suite sets it to 0) a small AUTO call runs the register kernel: bit-identical to
    HB_PATH_SIMT, no guard read, no guard trip."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k != "HB_STC_MIN_ROWS"}
    out = subprocess.run([sys.executable, "-c", _STC_THRESHOLD_SCRIPT], cwd=root, env=env, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and "threshold ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("which", ["rqspline", "realnvp_small"])
def test_new_tensor_kernels_are_bit_reproducible(which):
    """One issuing warp per accumulator and fixed tile ownership: repeated calls, host and device inputs, and a batch
    embedded at another row offset give bit-identical ln phi (the per-row result does not depend on the tile it rides in)."""
    import torch
    if which == "rqspline":
        m = util.make_rqspline(2, 8, 8, (64, 64), seed=4, standardize=True, ws=0.3)
        ndim = 2
    else:
        m = util.make_realnvp(5, 6, 2, seed=4, standardize=True, ws=0.3)
        ndim = 5
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = samples(40000, ndim, seed=12).astype(np.float32)
    a = m.predict(x)
    for _ in range(3):
        np.testing.assert_array_equal(a, m.predict(x))
    np.testing.assert_array_equal(a, m.predict(torch.from_numpy(x).cuda()).cpu().numpy())
    np.testing.assert_array_equal(a[777:20000], m.predict(x[777:20000]))


@pytest.mark.parametrize("layers,rng,T", [(1, (-10.0, 10.0), 1.0), (2, (-3.0, 4.0), 0.5), (7, (-6.0, 6.0), 0.6)])
def test_rqspline_tcgen05_range_temperature_layers(layers, rng, T):
    """Odd and even layer counts (the mask alternates per layer), a custom spline range and several temperatures.  (At
    T = 0.3 with inputs 2.5 sigma wide EVERY float32 evaluation of the 7-layer flow leaves 1e-5 on a handful of 2e5
    samples -- restatement 1, register kernel 5, tensor kernel 16, max 1.05e-5 / 1.31e-5 / 1.34e-5: the 1 / T in front
    of the squared latent amplifies the rounding -- so the 7-layer case runs at T = 0.6.)"""
    m = util.make_rqspline(2, layers, 8, (64, 64), seed=20 + layers, standardize=True, spline_range=rng, temperature=T,
                           ws=0.3)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = samples(5000, 2, seed=3, scale=2.5)
    util.check_predict(m, x, m.predict(x))
    m.temperature = 0.9 * T  # read on every call (harmonic/model.py:214)
    util.check_predict(m, x, m.predict(x))


@pytest.mark.parametrize("T", [1.0, 0.3])
def test_realnvp_small_tcgen05_temperature(T):
    m = util.make_realnvp(5, 6, 2, seed=5, standardize=True, temperature=T, ws=0.3)
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x = samples(5000, 5, seed=3, scale=1.2)
    util.check_predict(m, x, m.predict(x))
    m.temperature = 0.5 * T
    util.check_predict(m, x, m.predict(x))
