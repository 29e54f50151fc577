#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_em.py tests/test_fieldsolve.py tests/test_reference_glsl.py -m gpu -x -q > gpurun_out/r2_t_precalc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_precalc.log
tail -4 gpurun_out/r2_t_precalc.log
timeout 300 python tools/em_bench.py > gpurun_out/r2_em_bench_b.jsonl 2> gpurun_out/r2_em_bench_b.err; echo "em_bench rc=$?"; cat gpurun_out/r2_em_bench_b.jsonl; tail -3 gpurun_out/r2_em_bench_b.err
