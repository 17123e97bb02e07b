#!/bin/bash
# A/B of two builds on one box: tools/ab.sh <tag_new> ; baseline = harmonic_b200/build_base/libhb_base.so (development aid)
set -u
T=${1:-new}
timeout 150 python -m pytest tests/test_gpu_flows.py -x -q -m gpu -k tcgen05 2>&1 | tail -3
for i in 1 2; do
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_${T}_$i.json 2> gpurun_out/ab_${T}_$i.err
HARMONIC_B200_LIB=/root/repo/harmonic_b200/build_base/libhb_base.so timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_base_$i.json 2> gpurun_out/ab_base_$i.err
done
python - <<PY
import json
for f in ["ab_${T}_1","ab_base_1","ab_${T}_2","ab_base_2"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],2), round(d["roofline"]["kernel_ms"],2), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], d["result"]["ln_evidence_inv"])
    except Exception as e: print(f, "ERR", e)
PY
