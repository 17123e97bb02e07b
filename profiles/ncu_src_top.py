"""Summarise an ncu --page source --csv dump: top stall sites (SASS) by sample count.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python profiles/ncu_src_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ci["# Samples"]]) for r in data)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ci[s]]) for r in data) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][ci["# Samples"]]))[:n]:
    top = sorted(((int(r[ci[s]]), s) for s in stalls), reverse=True)[:2]
    print("%5d %6.2f%% exec=%-9s %-70s %s" % (idx, 100.0 * int(r[ci["# Samples"]]) / tot,
          r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:70], top))
