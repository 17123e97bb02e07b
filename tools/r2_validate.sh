#!/bin/bash
# round-2 validation of a build (tag = $1): GPU tests, smoke, default bench (development aid)
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_${TAG}.log
tail -8 gpurun_out/pytest_${TAG}.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/bench_default_${TAG}.json 2> gpurun_out/bench_default_${TAG}.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_default_${TAG}.err
cut -c1-600 gpurun_out/bench_default_${TAG}.json
