#!/bin/bash
# round-2 GPU call H: does the power-of-two stride between the four channel planes of the per-cell sums cost anything?
mkdir -p gpurun_out
export FSIM_LIB_PATH=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so
for pad in 0 4096 528 1052672 0; do
  echo "pad $pad" >> gpurun_out/r2_tune_h.txt
  FSIM_PLANE_PAD=$pad timeout 200 python tools/tune_sort_interval.py c5 f64 8 0 2>> gpurun_out/r2_tune_h.err | cut -c1-700 >> gpurun_out/r2_tune_h.txt
done
cat gpurun_out/r2_tune_h.txt; tail -2 gpurun_out/r2_tune_h.err
