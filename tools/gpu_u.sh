#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "run_frames" > gpurun_out/r2b_t_graph.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_t_graph.log
tail -15 gpurun_out/r2b_t_graph.log
FSIM_LIB_PATH=tools/scratch/ab/dbg/fusion_sim_b200/csrc/libfusionsim.so timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "run_frames" > gpurun_out/r2b_t_graph_dbg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_t_graph_dbg.log
tail -3 gpurun_out/r2b_t_graph_dbg.log
