"""Accuracy of the tcgen05 RealNVP path against the float64 oracle (development aid; numbers quoted in DESIGN.md 4.1).

usage (B200): python tools/tc_accuracy.py [nsamples]
Prints max / median |d ln phi| / |ln phi| for the tensor path and the float32 CUDA-core path at cfg4, for inputs
drawn from the target and 5x / 30x outside it."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench_workloads as wl  # noqa: E402
from harmonic_b200 import _lib  # noqa: E402
from oracle import flows as of  # noqa: E402  (the checker, not the product path)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    m = wl.make_model("cfg4")
    x, _ = wl.gaussian_host(64, n, 1)
    for mult in (1.0, 5.0, 30.0):
        xs = (x * mult).astype(np.float32)
        want = of.predict(m.oracle_spec(), xs.astype(np.float64), np.float64)
        for name, path in (("tcgen05", _lib.HB_PATH_TCGEN05), ("simt f32", _lib.HB_PATH_SIMT)):
            m.kernel_path = path
            got = np.asarray(m.predict(xs), dtype=np.float64)
            rel = np.abs(got - want) / np.abs(want)
            print("inputs x%-4g %-9s max rel %.2e  median %.2e  max abs %.2e  mean signed %.2e" % (
                mult, name, rel.max(), np.median(rel), np.abs(got - want).max(), (got - want).mean()))


if __name__ == "__main__":
    main()
