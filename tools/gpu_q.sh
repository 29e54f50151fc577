#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_glsl.py tests/test_parity_sizes_gpu.py -m gpu -x -q -k "canvas or render or c2 or C2 or demo or density" > gpurun_out/r2_t_render.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_render.log
tail -4 gpurun_out/r2_t_render.log
timeout 400 python tools/tune_sort_interval.py c5 f64 8 > gpurun_out/r2_frame_after_render.jsonl 2>&1; cat gpurun_out/r2_frame_after_render.jsonl
timeout 400 python tools/tune_sort_interval.py c5 f32 8 >> gpurun_out/r2_frame_after_render.jsonl 2>&1; tail -1 gpurun_out/r2_frame_after_render.jsonl
