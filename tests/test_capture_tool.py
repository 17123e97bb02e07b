"""tools/capture_reference_golden.py end to end against stand-ins for `harmonic` / `jax` (tests/stub_reference):
proves, in this JAX-less image, that the tool runs, that its .npz layout is what tests/test_reference_golden.py
reads, that the flattened flax parameter tree round-trips through load_params / the flat weight buffer, and that both
consumers (numpy oracle on CPU, CUDA kernels on a B200) reproduce the captured ln phi.  The REAL capture (one command
on a machine with the reference installed, see README.md) is what pins parity; this only keeps that path working."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from oracle import flows as of
from tests import util
from tests.test_reference_golden import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "stub_reference")


def run_capture(outdir, monkeypatch):
    for name in [k for k in sys.modules if k in ("harmonic", "jax") or k.startswith(("harmonic.", "jax."))]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(STUB)
    spec = importlib.util.spec_from_file_location("capture_reference_golden",
                                                  os.path.join(ROOT, "tools", "capture_reference_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["capture_reference_golden.py", str(outdir)])
    try:
        mod.main()
    finally:
        # the stand-ins must not outlive the call: scipy's array-API helpers probe sys.modules["jax"]
        for name, m in list(sys.modules.items()):
            if (getattr(m, "__file__", None) or "").startswith(STUB):
                del sys.modules[name]
    files = sorted(str(p) for p in outdir.glob("reference_*.npz"))
    assert len(files) == 5
    return files


def test_capture_tool_feeds_the_oracle_consumer(tmp_path, monkeypatch):
    for path in run_capture(tmp_path, monkeypatch):
        z = np.load(path)
        m = build(z)
        assert m.fitted and m.ndim == int(z["ndim"])
        # the flat weight buffer of the C ABI is built from the round-tripped tree without a shape complaint
        assert m._flat_weights().dtype == np.float32
        got = of.predict(m.oracle_spec(), z["x"], np.float32)
        util.assert_lnphi_close(got, z["lnphi"], rel=2e-5)
        assert float(of.predict(m.oracle_spec(), z["x"][0], np.float32)) == pytest.approx(float(z["lnphi_one"]), rel=2e-5)


@pytest.mark.gpu
def test_capture_tool_feeds_the_gpu_consumer(tmp_path, monkeypatch):
    for path in run_capture(tmp_path, monkeypatch):
        z = np.load(path)
        m = build(z)
        util.assert_lnphi_close(m.predict(z["x"]), z["lnphi"], rel=2e-5,
                                slack=3.0 * np.abs(of.predict(m.oracle_spec(), z["x"], np.float64) - z["lnphi"])
                                / np.maximum(np.abs(z["lnphi"]), 1.0))
        one = m.predict(z["x"][0])
        assert np.ndim(one) == 0
