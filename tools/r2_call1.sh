#!/bin/bash
# round-2 call 1: probes, GPU tests, default bench (development aid)
set -u
mkdir -p gpurun_out
timeout 120 tools/issue_probe.bin > gpurun_out/r2_issue_probe.txt 2>&1; echo "probe rc=$?"
cat gpurun_out/r2_issue_probe.txt | head -60
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest1.log
timeout 60 python __graft_entry__.py --smoke > gpurun_out/r2_smoke1.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench1.err
cut -c1-3000 gpurun_out/r2_bench1.json
