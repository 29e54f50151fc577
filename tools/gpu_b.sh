#!/bin/bash
# round-2 GPU call B: why is the sweep slow?  A/B over the second stream, the fused re-sort and the interval
mkdir -p gpurun_out
for f in 0 8 16 24; do
  timeout 600 python tools/tune_sort_interval.py c5 f64 8 $f >> gpurun_out/r2_tune_b.jsonl 2>> gpurun_out/r2_tune_b.err
done
timeout 600 python tools/tune_sort_interval.py c5 f64 2,4 0 >> gpurun_out/r2_tune_b.jsonl 2>> gpurun_out/r2_tune_b.err
cat gpurun_out/r2_tune_b.jsonl
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
tail -5 gpurun_out/r2_t2.log
