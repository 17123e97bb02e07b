#!/bin/bash
# A/B of variant builds on one box: tools/ab_libs.sh <lib path>...  ("-" = the in-tree library) (development aid)
set -u
for rep in 1 2; do
for L in "$@"; do
  if [ "$L" = "-" ]; then unset HARMONIC_B200_LIB; else export HARMONIC_B200_LIB=$L; fi
  timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-secondary ${BENCH_ARGS:-} > gpurun_out/ab_tmp.json 2>/dev/null
  python - "$L" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/ab_tmp.json")); print(sys.argv[1][-28:], round(d["ms_per_step"],2), round(d["roofline"]["kernel_ms"],2), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["result"]["ln_evidence_inv"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done; done
