"""Where does the small-ndim tensor path stop being float32-grade?  (development aid for the guard limit in
hb_realnvp_small_tc.cu).  Rows whose forced-tensor-path result is off by more than 1e-5 (or not finite) are evaluated
again, alone, under HB_PATH_AUTO with candidate guard limits: the guard count says how many of them a limit catches."""
import ctypes
import os
import sys

import numpy as np

os.environ["HB_STC_MIN_ROWS"] = "0"
sys.path.insert(0, ".")
import bench_workloads as wl  # noqa: E402
from harmonic_b200 import _lib  # noqa: E402
from oracle import flows as of  # noqa: E402


def main():
    n = 400000
    x, _ = wl.gaussian_host(5, n, 1)
    for mult in (4.0, 5.0, 8.0):
        xs = (x * mult).astype(np.float32)
        m = wl.make_model("cfg3")
        want = of.predict(m.oracle_spec(), xs.astype(np.float64), np.float64)
        w32 = of.predict(m.oracle_spec(), xs.astype(np.float64), np.float32).astype(np.float64)
        den = np.maximum(np.abs(want), 1.0)
        m.kernel_path = _lib.HB_PATH_TCGEN05
        got = np.asarray(m.predict(xs), dtype=np.float64)
        err = np.abs(got - want) / den
        e32 = np.abs(w32 - want) / den
        bad = ~(err <= np.maximum(1e-5, 3 * e32))  # NaN counts as bad
        good = ~bad
        print("x%g: %d of %d rows off by more than max(1e-5, 3x float32's own error) (%d not finite)" % (
            mult, bad.sum(), n, (~np.isfinite(got)).sum()))
        for L in (32, 64, 128, 256, 512, 1024, 2048):
            os.environ["HB_STC_GUARD_LIMIT"] = str(L)
            cnt = []
            for rows in (xs[bad], xs[good][:100000]):
                mm = wl.make_model("cfg3")
                mm.predict(rows)
                r, c = ctypes.c_int64(0), ctypes.c_int64(0)
                _lib.check(_lib.load().hb_flow_guard_stats(mm._get_handle(0), ctypes.byref(r), ctypes.byref(c)))
                cnt.append(r.value)
            print("   limit %5d: catches %d of the %d bad rows; trips on %d of %d good rows" % (
                L, cnt[0], bad.sum(), cnt[1], min(100000, good.sum())))
        os.environ.pop("HB_STC_GUARD_LIMIT")


if __name__ == "__main__":
    main()
