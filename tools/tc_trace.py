"""Phase timeline of realnvp_tc_kernel on one SM (development aid).

Build with  HB_NVCC_EXTRA=-DHB_TC_TRACE python harmonic_b200/build.py --force , run this on a B200, then
rebuild without the flag.  With  HB_NVCC_EXTRA="-DHB_TC_TRACE -DHB_TC_TRACE_SKEW"  and  --skew  the four streams
are the four warps of epilogue warpgroup 0 and the output is their spread at every stamp (round 1: ~80-120
cycles at the barrier arrivals; the warp that shares its scheduler with no MMA / producer warp leads).  CTA 0 records clock64() stamps from lane 0 of: epilogue warpgroup 0 / 1 (streams 0 / 1)
and the two MMA warps (streams 2 / 3).  Prints mean cycles per phase over steady-state tiles.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench_workloads as wl  # noqa: E402
from harmonic_b200 import _lib  # noqa: E402

NAMES = {1: "tile start", 2: "A written (A1FULL arrive)", 3: "HFULL seen", 4: "E1 q0 done", 5: "E1 q1 done",
         6: "E1 q2 done", 7: "E1 q3 done", 9: "OFULL seen", 10: "E2 done", 11: "scale math done", 12: "OFULL2 seen", 20: "mma: A1FULL seen",
         21: "mma: corr issued", 22: "mma: A2 q0 seen", 23: "mma: A2 q1 seen", 24: "mma: A2 q2 seen",
         25: "mma: A2 q3 seen", 26: "mma: q0 issued", 27: "mma: q1 issued", 28: "mma: q2 issued",
         29: "mma: q3 issued", 30: "mma: OFULL seen", 31: "mma: main issued",
         32: "mma: OCONS seen", 33: "mma: pass 2 issued"}


def main():
    lib = _lib.load()
    m = wl.make_model("cfg4")
    m.kernel_path = _lib.HB_PATH_TCGEN05
    x, _ = wl.device_data("cfg4", 148 * 2, 128 * int(os.environ.get("HB_TRACE_TILES", "24")), torch.device("cuda", 0), seed=1)
    for _ in range(2):
        m.predict(x)
    torch.cuda.synchronize()
    buf = np.zeros((4, 4096), dtype=np.uint64)
    n = np.zeros(4, dtype=np.int32)
    fn = lib.hb_debug_trace_read
    fn.restype = C.c_int
    assert fn(buf.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p)) == 0
    if "--raw" in sys.argv:  # dump (stream, clock, code) for offline analysis
        np.save("gpurun_out/tc_trace_raw.npy", buf)
        np.save("gpurun_out/tc_trace_n.npy", n)
    if "--skew" in sys.argv:
        clk = [(buf[s, :n[s]] >> np.uint64(8)).astype(np.int64) for s in range(4)]
        code = [(buf[s, :n[s]] & np.uint64(255)).astype(np.int64) for s in range(4)]
        L = int(min(n))
        assert all(np.array_equal(code[0][:L], code[s][:L]) for s in range(4)), "streams differ"
        c = np.stack([clk[s][:L] for s in range(4)])
        for cd in sorted(set(code[0][:L])):
            idx = np.nonzero(code[0][:L] == cd)[0]
            idx = idx[idx >= L // 4]
            if len(idx):
                sp = c[:, idx].max(axis=0) - c[:, idx].min(axis=0)
                rel = c[:, idx] - c[:, idx].min(axis=0, keepdims=True)
                print("%-28s n=%4d  spread mean %6.0f max %6d | mean lateness per warp %s" % (
                    NAMES.get(cd, cd), len(idx), sp.mean(), sp.max(), np.round(rel.mean(axis=1)).astype(int)))
        return
    for s in range(4):
        ev = buf[s, :n[s]]
        clk = (ev >> np.uint64(8)).astype(np.int64)
        code = (ev & np.uint64(255)).astype(np.int64)
        print("== stream %d (%s): %d events" % (s, ["epilogue wg0", "epilogue wg1", "mma tile0", "mma tile1"][s], n[s]))
        # per-transition statistics (prev code -> code), skipping the first quarter of events (cold start)
        lo = len(ev) // 4
        stats = {}
        for i in range(max(lo, 1), len(ev)):
            stats.setdefault((code[i - 1], code[i]), []).append(clk[i] - clk[i - 1])
        for (a, b), v in sorted(stats.items(), key=lambda kv: -np.sum(kv[1])):
            print("  %-28s -> %-28s  n=%4d  mean %7.0f  min %6d  max %6d  total %8d" % (
                NAMES.get(a, a), NAMES.get(b, b), len(v), np.mean(v), np.min(v), np.max(v), np.sum(v)))
        if len(ev) > lo + 1:
            print("  span %d cycles" % (clk[-1] - clk[lo]))
        if s == 0:
            # one full tile, raw, for a picture
            starts = np.nonzero(code == 1)[0]
            if len(starts) > 4:
                a, b = starts[3], starts[4]
                t0 = clk[a]
                print("  raw tile: " + " ".join("%d@%d" % (code[i], clk[i] - t0) for i in range(a, b + 1)))


if __name__ == "__main__":
    main()
