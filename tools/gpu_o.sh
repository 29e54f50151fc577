#!/bin/bash
mkdir -p gpurun_out
export FSIM_LIB_PATH=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so
timeout 400 python tools/tune.py c5 f32 4,31,36,37,38,39,35,34,0,4 > gpurun_out/r2_tune_occ_f32_b.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2_tune_occ_f32_b.txt | tail -12
timeout 400 python tools/tune.py c3 f32 4,31,32,0 > gpurun_out/r2_tune_occ_f32_c3.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2_tune_occ_f32_c3.txt | tail -5
