#!/bin/bash
# two-GPU check of the three-tile build: multi-GPU tests + the 2-GPU bench line (development aid)
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -rs 2>&1 | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/bench_n2_r1j.json 2> gpurun_out/bench_n2_r1j.err; echo "bench n2 rc=$?"
cut -c1-300 gpurun_out/bench_n2_r1j.json
