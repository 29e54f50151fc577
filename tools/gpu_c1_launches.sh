#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "run_frames" > gpurun_out/r2d_t_graph.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_t_graph.log
tail -3 gpurun_out/r2d_t_graph.log
CMD="python bench.py --workload c1 --graph off --steps 32 --warmup 16 --no-cpu-baseline --no-extension-probe"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2d_launches_c1.csv $CMD > gpurun_out/r2d_ncu_c1.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2d_launches_c1.csv")) if len(r) > 5]
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik].split("(")[0][:60]].append(float(r[iv].replace(",", "")))
    except Exception: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(k.ljust(60), len(v), "mean ns", round(sum(v) / len(v)))
PY
