"""Accuracy of the RQSpline kernels against the float64 oracle (development aid; numbers quoted in DESIGN.md 4.3):
python tools/rq_accuracy.py [nsamples].  Prints max / median |d ln phi| / max(|ln phi|, 1) and the number of samples
beyond 1e-5 for the tensor path and the CUDA-core register path, for the cfg2 flow and the test-suite flow."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench_workloads as wl  # noqa: E402
from harmonic_b200 import _lib  # noqa: E402
from oracle import flows as of  # noqa: E402  (the checker, not the product path)
from tests import util  # noqa: E402


def report(tag, m, x):
    want = of.predict(m.oracle_spec(), x.astype(np.float64), np.float64)
    w32 = of.predict(m.oracle_spec(), x.astype(np.float64), np.float32).astype(np.float64)
    den = np.maximum(np.abs(want), 1.0)
    e32 = np.abs(w32 - want) / den
    print("%-28s float32 restatement: max %.2e median %.2e  >1e-5: %d" % (tag, np.nanmax(e32), np.nanmedian(e32), int((e32 > 1e-5).sum())))
    for name, path in (("tcgen05", _lib.HB_PATH_TCGEN05), ("simt f32", _lib.HB_PATH_SIMT)):
        m.kernel_path = path
        got = np.asarray(m.predict(x), dtype=np.float64)
        err = np.abs(got - want) / den
        i = int(np.nanargmax(err))
        print("%-28s %-9s max %.2e (row %d, f32 err there %.1e) median %.2e  >1e-5: %d  mean signed %.2e" % (
            tag, name, err[i], i, e32[i], np.nanmedian(err), int((err > 1e-5).sum()), np.nanmean(got - want)))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    rng = np.random.default_rng(3)
    m = wl.make_model("cfg2")
    x = (rng.standard_normal((n, 2)) * np.array([1.0, 2.0]) + np.array([1.0, 1.0])).astype(np.float32)
    report("cfg2 flow, in range", m, x)
    report("cfg2 flow, 3x wider", m, (x * 3).astype(np.float32))
    m = util.make_rqspline(2, 8, 8, (64, 64), seed=10, standardize=True)
    x = (rng.standard_normal((n, 2)) * 2.0).astype(np.float32)
    report("test flow (ws default)", m, x)


if __name__ == "__main__":
    main()
