#!/bin/bash
# validation of a build (tag = $1): GPU tests, default bench, ncu full capture, launch list (development aid)
set -u
TAG=${1:-r1k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_${TAG}.log
tail -3 gpurun_out/pytest_${TAG}.log
timeout 60 python __graft_entry__.py --smoke > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"
HB_TC_TILES=2 timeout 200 python -m pytest tests/test_gpu_flows.py -x -q -m gpu -k tcgen05 2>&1 | tail -2
HB_TC_TILES=2 timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_nt2_${TAG}.json 2> gpurun_out/bench_nt2_${TAG}.err
timeout 300 python bench.py > gpurun_out/bench_default_${TAG}.json 2> gpurun_out/bench_default_${TAG}.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_default_${TAG}.json
timeout 120 python bench.py --steps 1 --warmup 1 --scale 0.02 --no-e2e --no-cpu > gpurun_out/b_pre_${TAG}.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:realnvp_tc -c 1 -f -o gpurun_out/prof_tc_${TAG} \
  python bench.py --steps 1 --warmup 1 --scale 0.02 --no-e2e --no-cpu > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1; echo "ncu launches rc=$?"
