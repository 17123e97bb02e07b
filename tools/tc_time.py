"""Kernel time of the forced tcgen05 path on device-resident cfg4 data (development aid for variant builds whose
results may be wrong on purpose): python tools/tc_time.py [nsamples]"""
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench_workloads as wl  # noqa: E402
from harmonic_b200 import _lib  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
    m = wl.make_model("cfg4")
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x, _ = wl.device_data("cfg4", 1000, n // 1000, torch.device("cuda", 0), seed=1)
    for _ in range(2):
        m.predict(x)
    torch.cuda.synchronize()
    ts = []
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4  # ~16 back-to-back repetitions reach the power-capped steady state
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.predict(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("predict ms for %d samples:" % n, " ".join("%.2f" % t for t in ts),
          "-> per 1e8: first %.2f, last %.2f" % (ts[0] * 1e8 / n, ts[-1] * 1e8 / n))


if __name__ == "__main__":
    main()
