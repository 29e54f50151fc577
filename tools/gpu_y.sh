#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_glsl.py -m gpu -x -q > gpurun_out/r2e_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_t.log
tail -3 gpurun_out/r2e_t.log
FSIM_LIB_PATH=tools/scratch/ab/dbg/fusion_sim_b200/csrc/libfusionsim.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "crowded or c1 or demo or density or edge" > gpurun_out/r2e_t_dbg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_t_dbg.log
tail -2 gpurun_out/r2e_t_dbg.log
for w in c1 c2; do timeout 300 python bench.py --workload $w --no-cpu-baseline --no-extension-probe > gpurun_out/r2e_bench_$w.json 2>/dev/null; python -c "import json,sys; d=json.load(open('gpurun_out/r2e_bench_$w.json')); k=d['roofline']['kernels_ms_per_step']; print('$w %.3e'%d['value'], d['ms_per_step'], d['check']['ok'], round(k['cellsum_warp']['ms_per_launch']*1000,1), round(k['cellsum_heavy']['ms_per_launch']*1000,1), d['clocks']['samples'])"; done
