#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extension-probe --precision f32"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 4 -c 2 -o gpurun_out/r2b_prof_c5_f32 $CMD > gpurun_out/r2b_ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3; tail -2 gpurun_out/r2_plain.log | cut -c1-300
