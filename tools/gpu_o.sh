#!/bin/bash
mkdir -p gpurun_out
export FSIM_LIB_PATH=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so
timeout 400 python tools/tune.py c5 f32 0,40,41,42,0,40 > gpurun_out/r2_tune_rec12_f32.txt 2>&1; echo "rc=$?"; tail -7 gpurun_out/r2_tune_rec12_f32.txt
timeout 400 python tools/tune.py c5 f64 0,40,41,0 > gpurun_out/r2_tune_rec12_f64.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_tune_rec12_f64.txt
