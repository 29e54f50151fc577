#!/bin/bash
mkdir -p gpurun_out
FSIM_TEST_WORLD=2 timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q --durations=6 > gpurun_out/r2_t_dist2_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_dist2_b.log
tail -4 gpurun_out/r2_t_dist2_b.log
bash tools/gpu_n.sh 2 c5 --steps 20 --warmup 5
