#!/bin/bash
# N-GPU bench line (development aid): tools/bench_n.sh N tag
set -u
N=${1:-8}; TAG=${2:-r1k}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err; echo "bench n$N rc=$?"
cut -c1-260 gpurun_out/bench_n${N}_${TAG}.json
