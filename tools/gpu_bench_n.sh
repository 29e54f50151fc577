#!/bin/bash
# tools/gpu_bench_n.sh N workload [extra bench args]: one slab bench line on N GPUs
N=$1; W=$2; shift 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W "$@" > gpurun_out/r2_bench_${W}_n${N}_b.json 2> gpurun_out/r2_bench_${W}_n${N}_b.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_${W}_n${N}_b.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_${W}_n${N}_b.json"))
print("%.4e" % d["value"], round(d["ms_per_step"], 3), "e2e %.3e" % d["e2e"]["value"], d["check"]["ok"])
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in (d["comm_ms_per_step"] or {}).items() if k != "what"})
print(d["per_rank"])
PY
