#!/bin/bash
# round-2 GPU call E: sweep variants (entropy L1 no-allocate, early second entropy gather), interval sweep, fp32 variants
mkdir -p gpurun_out
T=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so
FSIM_LIB_PATH=$T timeout 900 python tools/tune.py c5 f64 4,20,21,22,23,24,10 0 > gpurun_out/r2_tune_e_f64.txt 2> gpurun_out/r2_tune_e.err
FSIM_LIB_PATH=$T timeout 900 python tools/tune.py c5 f32 4,20,21,22,2,7,1 0 > gpurun_out/r2_tune_e_f32.txt 2>> gpurun_out/r2_tune_e.err
timeout 900 python tools/tune_sort_interval.py c5 f64 6,8,12,16 0 > gpurun_out/r2_tune_e_interval.jsonl 2>> gpurun_out/r2_tune_e.err
cat gpurun_out/r2_tune_e_f64.txt gpurun_out/r2_tune_e_f32.txt gpurun_out/r2_tune_e_interval.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_spindle.py -m gpu -x -q > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
tail -3 gpurun_out/r2_t4.log; tail -3 gpurun_out/r2_tune_e.err
