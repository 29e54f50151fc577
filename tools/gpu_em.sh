#!/bin/bash
# EM slice: parity + physics on the GPU, then the same tests on the debug-bounds build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_em.py -m gpu -x -q --durations=6 > gpurun_out/r2_t_em.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_em.log
tail -12 gpurun_out/r2_t_em.log
FSIM_LIB_PATH=tools/scratch/ab/dbg/fusion_sim_b200/csrc/libfusionsim.so timeout 600 python -m pytest tests/test_em.py -m gpu -x -q > gpurun_out/r2_t_em_dbg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_em_dbg.log
tail -4 gpurun_out/r2_t_em_dbg.log
timeout 300 python tools/em_bench.py > gpurun_out/r2_em_bench.jsonl 2> gpurun_out/r2_em_bench.err; echo "em_bench rc=$?"; cat gpurun_out/r2_em_bench.jsonl; tail -3 gpurun_out/r2_em_bench.err
