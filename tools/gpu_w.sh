#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_parity_sizes_gpu.py tests/test_reference_glsl.py -m gpu -x -q > gpurun_out/r2d_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_t.log
tail -3 gpurun_out/r2d_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu-baseline --no-extension-probe 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e'%d['value'], d['ms_per_step'], d['check']['ok'], d['roofline']['frac'])"
