"""Accuracy of the small-ndim RealNVP kernels against the float64 oracle (development aid; numbers quoted in DESIGN.md
4.3): python tools/stc_accuracy.py [nsamples].  Prints max / median |d ln phi| / max(|ln phi|, 1) and the number of
samples beyond 1e-5 for the tensor path and the register path, for the cfg3 flow in and out of range."""
import os
import sys

import numpy as np

os.environ.setdefault("HB_STC_MIN_ROWS", "0")  # AUTO = the tensor path at every size

sys.path.insert(0, ".")
import bench_workloads as wl  # noqa: E402
from harmonic_b200 import _lib  # noqa: E402
from oracle import flows as of  # noqa: E402  (the checker, not the product path)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    m = wl.make_model("cfg3")
    x, _ = wl.gaussian_host(5, n, 1)
    mults = [float(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1.0, 2.0, 3.0, 5.0, 30.0)
    for mult in mults:
        xs = (x * mult).astype(np.float32)
        want = of.predict(m.oracle_spec(), xs.astype(np.float64), np.float64)
        w32 = of.predict(m.oracle_spec(), xs.astype(np.float64), np.float32).astype(np.float64)
        den = np.maximum(np.abs(want), 1.0)
        e32 = np.abs(w32 - want) / den
        print("inputs x%-4g float32 restatement max %.2e median %.2e  >1e-5: %d" % (mult, e32.max(), np.median(e32), int((e32 > 1e-5).sum())))
        for name, path in (("tcgen05", _lib.HB_PATH_TCGEN05), ("auto", _lib.HB_PATH_AUTO), ("simt f32", _lib.HB_PATH_SIMT)):
            m.kernel_path = path
            got = np.asarray(m.predict(xs), dtype=np.float64)
            err = np.abs(got - want) / den
            print("inputs x%-4g %-9s max %.2e median %.2e  >1e-5: %d  NaN: %d (oracle %d)  mean signed %.2e" % (
                mult, name, np.nanmax(err), np.nanmedian(err), int((err > 1e-5).sum()), int(np.isnan(got).sum()),
                int(np.isnan(want).sum()), np.nanmean(got - want)))


if __name__ == "__main__":
    main()
