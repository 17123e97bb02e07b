#!/bin/bash
# final-tree validation + captures (tag r2s): GPU tests, smoke, default bench, ncu of the flow kernels, launch list
set -u
bash tools/r2_validate.sh r2s
bash tools/r2_profile.sh r2s
for W in cfg3 cfg2; do
K=realnvp_small_tc; [ $W = cfg2 ] && K=rqspline_tc
timeout 120 python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/b_pre_${W}_r2s.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o gpurun_out/prof_${W}_r2s \
  python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_${W}_r2s.log 2>&1; echo "ncu $W rc=$?"
done
