#!/bin/bash
# variants of the CUDA-core flow kernels on one box (development aid): tools/r2g_ab.sh "<workloads>" <lib>...
set -u
WL=$1; shift
for rep in 1 2; do
for L in "$@"; do
  if [ "$L" = "-" ]; then unset HARMONIC_B200_LIB; else export HARMONIC_B200_LIB=$L; fi
  for W in $WL; do
  timeout 100 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/ab_tmp.json 2>/dev/null
  python - "$L" $W <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/ab_tmp.json")); print(sys.argv[2], sys.argv[1][-28:], round(d["ms_per_step"],3), round(d["roofline"]["kernel_ms"],3), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["result"]["ln_evidence_inv"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
  done
done; done
