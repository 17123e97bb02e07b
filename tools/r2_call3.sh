#!/bin/bash
# round-2 call 3: E1 sequence probe, GPU evidence tests after the finalize rewrite, cfg5 step (development aid)
set -u
mkdir -p gpurun_out
timeout 120 tools/alu_probe.bin > gpurun_out/r2_alu_probe.txt 2>&1; echo "probe rc=$?"
cat gpurun_out/r2_alu_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest3.log
timeout 120 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_cfg5_fin.json 2>gpurun_out/r2_cfg5_fin.err; echo "cfg5 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_cfg5_fin.json")); print("cfg5 ms/step", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms"], "value", d["value"])
PY
