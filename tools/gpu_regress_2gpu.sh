#!/bin/bash
# 2-GPU regression after the last kernel changes: world-2 tests (NCCL), C5 slab bench line
mkdir -p gpurun_out
FSIM_TEST_WORLD=2 timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q --durations=6 > gpurun_out/r2b_t_dist2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_t_dist2.log
tail -4 gpurun_out/r2b_t_dist2.log
bash tools/gpu_bench_n.sh 2 c5 --steps 20 --warmup 5
