#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 tools/issue_probe.bin > gpurun_out/r2_issue_probe2.txt 2>&1; echo "probe rc=$?"
grep "shared acc" gpurun_out/r2_issue_probe2.txt
timeout 120 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_cfg5_pre.json 2>gpurun_out/r2_cfg5_pre.err; echo "cfg5 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg5.csv \
  python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_ncu_cfg5.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/r2_launches_cfg5.csv | awk -F'","' '{print $5, $NF}' | tail -30
