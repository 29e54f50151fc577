"""Kernel tuning sweep on one B200 (not part of the product): builds the bench scene once and
times frames under each push variant.  Needs the tuning build of the library:
    tools/ab_build.sh tune WORK -DFSIM_TUNE
    FSIM_LIB_PATH=tools/scratch/ab/tune/fusion_sim_b200/csrc/libfusionsim.so python tools/tune.py [workload] [precision] [variants]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fusion_sim_b200 import makeCylindricalParticlePusher  # noqa: E402
from fusion_sim_b200.scenes import apply_scene  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c5"
precision = sys.argv[2] if len(sys.argv) > 2 else "f64"
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(7))
sc = bench.build_scene(workload, 0, 1)
spec = dict(sc["spec"], precision=precision, flags=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
sim = makeCylindricalParticlePusher(spec)
import ctypes as _C  # noqa: E402
_L = _C.CDLL(os.environ["FSIM_LIB_PATH"])  # the tuning build exports fsim_tune_set
apply_scene(sim, sc)
names = ("push", "push2", "push2_resort", "scan", "permute", "index_scatter", "cellsum", "cellsum_warp", "cellsum_heavy", "conv", "prepass")
out = {}
for v in variants:
    _L.fsim_tune_set(int(v), 0)
    for _ in range(3):
        sim.step(); sim.density()
    sim.sync()
    sim.timing(True); sim.timing_reset()
    sim.mark(0)
    K = 16
    for _ in range(K):
        sim.step(); sim.density()
    sim.mark(1)
    ms = sim.elapsed_ms(0, 1) / K
    row = {"frame_ms": round(ms, 3)}
    for nm in names:
        t, c = sim.timing_get(nm)
        if c:
            row[nm] = round(t / c, 3)
    sim.timing(False)
    out[v] = row
    print(v, json.dumps(row), flush=True)
