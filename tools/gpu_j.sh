#!/bin/bash
# round-2 GPU call J (1 GPU): sort on ingest -- whole test suite, debug-bounds build, smoke, bench line, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log
tail -3 gpurun_out/r2_t7.log
FSIM_LIB_PATH=tools/scratch/ab/dbg/fusion_sim_b200/csrc/libfusionsim.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_periodic.py tests/test_spindle.py -m gpu -x -q > gpurun_out/r2_t7_debug_bounds.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7_debug_bounds.log
tail -2 gpurun_out/r2_t7_debug_bounds.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
/usr/bin/time -v timeout 600 python bench.py > gpurun_out/r2_bench_c5_n1_b.json 2> gpurun_out/r2_bench_c5_n1_b.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/r2_bench_c5_n1_b.err
/usr/bin/time -v timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_c5.json 2> gpurun_out/r2_bench_ref_c5.err; echo "ref rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/r2_bench_ref_c5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c5_n1_b.json"))
print("%.4e" % d["value"], round(d["ms_per_step"], 4), "e2e %.4e" % d["e2e"]["value"], round(d["e2e"]["ms_total"], 1), d["clocks"], d["check"]["ok"], d["gpu_launches"])
print({k: round(v["ms_per_step"], 4) for k, v in d["roofline"]["kernels_ms_per_step"].items()})
r = json.load(open("gpurun_out/r2_bench_ref_c5.json"))
print("ref %.4e" % r["value"], r["steps"], r["cpu_baseline"]["sample"], r["config"] == d["config"])
PY
