#!/bin/bash
# round-2 GPU call I: index scatter from the ranks the histogram atomics return -- parity (product + debug-bounds build), A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_periodic.py tests/test_parity_sizes_gpu.py -m gpu -x -q -k "not thousand" > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log
tail -3 gpurun_out/r2_t6.log
FSIM_LIB_PATH=tools/scratch/ab/dbg/fusion_sim_b200/csrc/libfusionsim.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_periodic.py -m gpu -x -q > gpurun_out/r2_t6_debug_bounds.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6_debug_bounds.log
tail -3 gpurun_out/r2_t6_debug_bounds.log
for f in 0 256 0 256; do
  timeout 200 python tools/tune_sort_interval.py c5 f64 8 $f 2>> gpurun_out/r2_tune_i.err | cut -c1-560 >> gpurun_out/r2_tune_i.jsonl
done
timeout 200 python tools/tune_sort_interval.py c5 f32 8 0 2>> gpurun_out/r2_tune_i.err | cut -c1-560 >> gpurun_out/r2_tune_i.jsonl
timeout 200 python tools/tune_sort_interval.py c5 f32 8 256 2>> gpurun_out/r2_tune_i.err | cut -c1-560 >> gpurun_out/r2_tune_i.jsonl
cat gpurun_out/r2_tune_i.jsonl; tail -2 gpurun_out/r2_tune_i.err
