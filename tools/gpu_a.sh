#!/bin/bash
# round-2 GPU call A: environment facts, GPU test suite, default bench line, access-pattern microbenchmark
mkdir -p gpurun_out
{ free -g; nproc; nvidia-smi -L; } > gpurun_out/r2_env.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
tail -5 gpurun_out/r2_t1.log
timeout 900 python bench.py > gpurun_out/r2_bench_c5_a.json 2> gpurun_out/r2_bench_c5_a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_c5_a.err
tools/scratch/stream_layout_bench > gpurun_out/r2_stream_layout.jsonl 2>&1
cat gpurun_out/r2_stream_layout.jsonl
