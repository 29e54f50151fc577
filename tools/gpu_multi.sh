#!/bin/bash
# round-2 multi-GPU call: tools/gpu_multi.sh N   (N = 2, 4 or 8 GPUs of one box)
N=$1
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 "${@:2}"; }
export -f run
if [ "$N" != "2" ]; then
  FSIM_TEST_WORLD=$N timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q --durations=6 > gpurun_out/r2_t_dist$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_dist$N.log
  tail -4 gpurun_out/r2_t_dist$N.log
  FSIM_LIB_PATH=$PWD/tools/scratch/ab/dbg/fusion_sim_b200/csrc/libfusionsim.so FSIM_TEST_WORLD=$N timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q -k "match_single" > gpurun_out/r2_t_dist${N}_debug_bounds.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t_dist${N}_debug_bounds.log
  tail -2 gpurun_out/r2_t_dist${N}_debug_bounds.log
  timeout 600 bash -c "run $N --steps 20 --warmup 5" > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err; echo "c5 fixed rc=$?"
fi
if [ "$N" == "8" ]; then
  timeout 600 bash -c "run $N --steps 20 --warmup 5 --exchange exact --no-reduced-check" > gpurun_out/r2_bench_c5_n${N}_exact.json 2> gpurun_out/r2_bench_c5_n${N}_exact.err; echo "c5 exact rc=$?"
fi
timeout 900 bash -c "run $N --workload c4 --steps 10 --warmup 3 --no-reduced-check" > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err; echo "c4 rc=$?"
tail -c 300 gpurun_out/r2_bench_c4_n$N.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_c*_n$N*.json")):
    try:
        d = json.load(open(f))
        print(f, "%.4e" % d["value"], round(d["ms_per_step"], 3), "e2e %.3e" % d["e2e"]["value"], d["check"]["ok"], {k: (round(v, 4) if isinstance(v, float) else v) for k, v in (d["comm_ms_per_step"] or {}).items() if k != "what"})
    except Exception as e:
        print(f, "no line", e)
PY
