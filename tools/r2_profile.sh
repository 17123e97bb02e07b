#!/bin/bash
# round-2 ncu captures (tag = $1): full capture of the tcgen05 kernel and of the accumulation-only kernel, launch list
# of the default bench command (development aid; each capture only after the same command exited 0 without ncu)
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 120 python bench.py --steps 1 --warmup 1 --scale 0.02 --no-e2e --no-cpu --no-secondary > gpurun_out/b_pre_${TAG}.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:realnvp_tc -c 1 -f -o gpurun_out/prof_tc_${TAG} \
  python bench.py --steps 1 --warmup 1 --scale 0.02 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu tc rc=$?"
timeout 120 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_pre5_${TAG}.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:accumulate_precomputed -c 1 -f -o gpurun_out/prof_acc_${TAG} \
  python bench.py --workload cfg5 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu5_${TAG}.log 2>&1; echo "ncu acc rc=$?"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/b_pre_l_${TAG}.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/ncu_launches_${TAG}.log 2>&1; echo "ncu launches rc=$?"
