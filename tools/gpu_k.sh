#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s); timeout 600 python bench.py > gpurun_out/r2_bench_c5_n1_b.json 2> gpurun_out/r2_bench_c5_n1_b.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"
t0=$(date +%s); timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_c5.json 2> gpurun_out/r2_bench_ref_c5.err; echo "ref rc=$? $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c5_n1_b.json"))
print("%.4e" % d["value"], round(d["ms_per_step"], 4), "e2e %.4e" % d["e2e"]["value"], round(d["e2e"]["ms_total"], 1), d["clocks"], d["check"]["ok"], d["gpu_launches"])
print({k: round(v["ms_per_step"], 4) for k, v in d["roofline"]["kernels_ms_per_step"].items()})
r = json.load(open("gpurun_out/r2_bench_ref_c5.json"))
print("ref %.4e" % r["value"], r["steps"], r["cpu_baseline"]["sample"], r["config"] == d["config"])
PY
