#!/bin/bash
# ncu pair for the fp32 sweep: the 128-bit-load variant (4 particles per thread, tuning build variant 7) beside the
# product's capture (profiles/r2b_ncu_full_summary_c5_f32.csv)
mkdir -p gpurun_out
export FSIM_LIB_PATH=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so
CMD="python tools/tune.py c5 f32 7"
timeout 60 $CMD > gpurun_out/r2g_tune_v7.txt 2>&1 &&
timeout 100 ncu --set full --clock-control none -k regex:push_kernel -s 6 -c 1 -o gpurun_out/r2g_prof_c5_f32_v4 $CMD > gpurun_out/r2g_ncu_v4.log 2>&1
tail -1 gpurun_out/r2g_tune_v7.txt | cut -c1-200; ls -la gpurun_out/r2g_prof_c5_f32_v4.ncu-rep 2>&1 | tail -1
