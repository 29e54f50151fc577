#!/bin/bash
# round-2 GPU call F: sweep variants (entropy gather without L1 allocation / issued early), short intervals
mkdir -p gpurun_out
T=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so
FSIM_LIB_PATH=$T timeout 300 python tools/tune.py c5 f64 4,20,21,22,23,4 0 > gpurun_out/r2_tune_f_f64.txt 2> gpurun_out/r2_tune_f.err
FSIM_LIB_PATH=$T timeout 300 python tools/tune.py c5 f32 4,20,21,22,2,7 0 > gpurun_out/r2_tune_f_f32.txt 2>> gpurun_out/r2_tune_f.err
timeout 300 python tools/tune_sort_interval.py c5 f64 4,5,6,7 0 > gpurun_out/r2_tune_f_interval.jsonl 2>> gpurun_out/r2_tune_f.err
cat gpurun_out/r2_tune_f_f64.txt gpurun_out/r2_tune_f_f32.txt; cut -c1-330 gpurun_out/r2_tune_f_interval.jsonl; tail -3 gpurun_out/r2_tune_f.err
