#!/bin/bash
# relu-E1 variant against HEAD on one box (development aid)
set -u
timeout 200 python -m pytest tests/test_gpu_flows.py -x -q -m gpu -k "tcgen05 or large_magnitude" 2>&1 | tail -3
timeout 120 python tools/tc_accuracy.py 200000 2>&1 | tail -6
echo "--- base accuracy"
HARMONIC_B200_LIB=/root/repo/harmonic_b200/build_base/libhb_base.so timeout 120 python tools/tc_accuracy.py 200000 2>&1 | grep tcgen05
bash tools/ab_libs.sh - /root/repo/harmonic_b200/build_base/libhb_base.so
echo "--- tc_time new"; timeout 100 python tools/tc_time.py 5e7 16
echo "--- tc_time base"; HARMONIC_B200_LIB=/root/repo/harmonic_b200/build_base/libhb_base.so timeout 100 python tools/tc_time.py 5e7 16
