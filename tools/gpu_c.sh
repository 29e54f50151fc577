#!/bin/bash
# round-2 GPU call C: TMA-staged sweep: parity, then A/B against per-thread loads and over the second-stream options
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_sizes_gpu.py -m gpu -x -q -k "not c5 and not c1_thousand" > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -4 gpurun_out/r2_t3.log
for f in 8 40 0 64 32; do
  timeout 600 python tools/tune_sort_interval.py c5 f64 8 $f >> gpurun_out/r2_tune_c.jsonl 2>> gpurun_out/r2_tune_c.err
done
FSIM_LIB_PATH=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so timeout 900 python tools/tune.py c5 f64 10,11,12,13,14,15,4 8 > gpurun_out/r2_tune_c_variants.txt 2>> gpurun_out/r2_tune_c.err
FSIM_LIB_PATH=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so timeout 900 python tools/tune.py c5 f32 10,11,12,13,15,4,2,7 8 > gpurun_out/r2_tune_c_variants_f32.txt 2>> gpurun_out/r2_tune_c.err
cat gpurun_out/r2_tune_c.jsonl gpurun_out/r2_tune_c_variants.txt gpurun_out/r2_tune_c_variants_f32.txt
tail -3 gpurun_out/r2_tune_c.err
