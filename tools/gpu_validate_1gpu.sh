#!/bin/bash
# round-2 final 1-GPU call: whole GPU test suite, debug-bounds build, smoke, bench lines, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r2b_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_t.log
tail -4 gpurun_out/r2b_t.log
FSIM_LIB_PATH=tools/scratch/ab/dbg/fusion_sim_b200/csrc/libfusionsim.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_periodic.py tests/test_em.py tests/test_parity_sizes_gpu.py -m gpu -x -q -k "not c5 and not thousand" > gpurun_out/r2b_t_debug_bounds.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_t_debug_bounds.log
tail -3 gpurun_out/r2b_t_debug_bounds.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2b_smoke.log 2>&1; tail -1 gpurun_out/r2b_smoke.log
timeout 600 python bench.py > gpurun_out/r2b_bench_c5_n1.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --precision f32 --no-cpu-baseline > gpurun_out/r2b_bench_c5_f32.json 2>> gpurun_out/r2b_bench.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2b_bench_c3.json 2>> gpurun_out/r2b_bench.err
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2b_bench_c2.json 2>> gpurun_out/r2b_bench.err
timeout 300 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r2b_bench_c1.json 2>> gpurun_out/r2b_bench.err
timeout 300 python bench.py --field-sweeps 8 --no-cpu-baseline > gpurun_out/r2b_bench_c5_fieldsolve8.json 2>> gpurun_out/r2b_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extension-probe"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_c5_f64.csv $CMD > gpurun_out/r2b_ncu_a.log 2>&1
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'push_kernel|index_scatter|cellsum_kernel|conv_kernel|render_kernel' -s 12 -c 10 -o gpurun_out/r2b_prof_c5_f64 $CMD > gpurun_out/r2b_ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
python - <<'PY'
import json
for nm in ("c5_n1", "c5_f32", "c3", "c2", "c1", "c5_fieldsolve8"):
    try:
        d = json.load(open(f"gpurun_out/r2b_bench_{nm}.json"))
        print(nm, "%.3e" % d["value"], round(d["ms_per_step"], 4), "e2e %.3e" % d["e2e"]["value"], "frac", round(d["roofline"]["frac"], 3), "frame frac", round(d["roofline"]["frame"]["frac"], 3), d["clocks"]["samples"], d["check"]["ok"])
    except Exception as e:
        print(nm, "no line", e)
PY
