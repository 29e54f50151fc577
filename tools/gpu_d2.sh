#!/bin/bash
# round-2 GPU call D2 (2 GPUs): multi-rank tests at world 2, slab bench at N=2 (fixed-region and exact exchange)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q --durations=8 > gpurun_out/r2_t_dist2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_dist2.log
tail -6 gpurun_out/r2_t_dist2.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); run 2 --steps 20 --warmup 5" > gpurun_out/r2_bench_c5_n2_fixed.json 2> gpurun_out/r2_bench_c5_n2_fixed.err; echo "bench fixed rc=$?"
timeout 600 bash -c "$(declare -f run); run 2 --steps 20 --warmup 5 --exchange exact --no-reduced-check" > gpurun_out/r2_bench_c5_n2_exact.json 2> gpurun_out/r2_bench_c5_n2_exact.err; echo "bench exact rc=$?"
tail -c 400 gpurun_out/r2_bench_c5_n2_fixed.err
python - <<'PY'
import json
for nm in ("fixed", "exact"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_c5_n2_{nm}.json"))
        print(nm, d["value"], d["ms_per_step"], d["e2e"]["value"], d["check"]["ok"], d["comm_ms_per_step"])
    except Exception as e:
        print(nm, "no line", e)
PY
