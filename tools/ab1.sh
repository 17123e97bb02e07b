#!/bin/bash
# one-shot A/B of two builds (short version of ab.sh): tools/ab1.sh <tag>
set -u
T=${1:-new}
timeout 100 python -m pytest tests/test_gpu_flows.py -x -q -m gpu -k tcgen05 2>&1 | tail -3
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_${T}_1.json 2> gpurun_out/ab_${T}_1.err
HARMONIC_B200_LIB=/root/repo/harmonic_b200/build_base/libhb_base.so timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_base_1.json 2> gpurun_out/ab_base_1.err
python - <<PY
import json
for f in ["ab_${T}_1","ab_base_1"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],2), round(d["roofline"]["kernel_ms"],2), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["result"]["ln_evidence_inv"])
    except Exception as e: print(f, "ERR", e)
PY
