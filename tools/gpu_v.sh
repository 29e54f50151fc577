#!/bin/bash
# last 1-GPU validation of the round: whole GPU suite, debug-bounds subset, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_t.log
tail -4 gpurun_out/r2c_t.log
FSIM_LIB_PATH=tools/scratch/ab/dbg/fusion_sim_b200/csrc/libfusionsim.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_periodic.py tests/test_em.py tests/test_parity_sizes_gpu.py -m gpu -x -q -k "not c5 and not thousand" > gpurun_out/r2c_t_debug_bounds.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_t_debug_bounds.log
tail -3 gpurun_out/r2c_t_debug_bounds.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c_smoke.log 2>&1; tail -1 gpurun_out/r2c_smoke.log
timeout 600 python bench.py > gpurun_out/r2c_bench_c5_n1.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2c_bench_ref.json 2>> gpurun_out/r2c_bench.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench_c5_n1.json"))
print("%.3e" % d["value"], round(d["ms_per_step"], 4), "e2e %.3e" % d["e2e"]["value"], "frac", round(d["roofline"]["frac"], 3), d["clocks"]["samples"], d["check"]["ok"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
r = json.load(open("gpurun_out/r2c_bench_ref.json")); print("ref", r.get("value"), r.get("impl"), r.get("config") == d["config"])
PY
