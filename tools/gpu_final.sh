#!/bin/bash
# last call of the round: whole GPU suite + smoke + default bench line on the final library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2f_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_t.log
tail -4 gpurun_out/r2f_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2f_bench_c5_n1.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2f_bench_c5_n1.json')); print('%.3e'%d['value'], d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d['check']['ok'], round(d['roofline']['frac'],3), d['clocks']['samples'])"
