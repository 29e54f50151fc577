#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_exp_cellsum_*.jsonl
for rep in 1 2; do
timeout 300 python tools/tune_sort_interval.py c5 f64 8 >> gpurun_out/r2_exp_cellsum_base.jsonl 2>&1
for mb in 10 12; do
FSIM_LIB_PATH=tools/scratch/ab/mb$mb/fusion_sim_b200/csrc/libfusionsim.so timeout 300 python tools/tune_sort_interval.py c5 f64 8 >> gpurun_out/r2_exp_cellsum_mb$mb.jsonl 2>&1
done
done
python - <<'PY'
import json
for f in ("base","mb10","mb12"):
    for l in open(f"gpurun_out/r2_exp_cellsum_{f}.jsonl"):
        try:
            d=json.loads(l); print(f, d["frame_ms"], d["cellsum"]["per_launch"], d["push2"]["per_launch"])
        except Exception as e: print(f, l[:100])
PY
