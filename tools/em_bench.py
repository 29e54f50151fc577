"""Measurement tool: frames of the electromagnetic loop (EXTENSION, BASELINE.json configs[2]: "Boris + Yee FDTD, 16M
particles, 2048x2048 grid, 1 B200") on one GPU -- half_step() + density() + emStep() + the canvas draws -- with the
device time of every kernel and the HBM figure of the three field kernels.
    python tools/em_bench.py [particles] [nr] [nz] [precision] [frames]
One JSON line.  dt is set from the Courant limit of the mesh (c dt sqrt(1/dr^2 + 1/dz^2) = 0.5)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion_sim_b200 import makeCylindricalParticlePusher  # noqa: E402
from fusion_sim_b200.scenes import apply_scene, c1_sink_source, plasma_particles, scaled_loops, scaled_spec  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
nz = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
precision = sys.argv[4] if len(sys.argv) > 4 else "f64"
K = int(sys.argv[5]) if len(sys.argv) > 5 else 100
CL = 2.998e8
spec = scaled_spec(nr, nz, n, precision=precision)
dr, dz = spec["radius"] / nr, spec["height"] / nz
spec["dt"] = 0.5 / (CL * np.sqrt(1 / dr ** 2 + 1 / dz ** 2))
spec["keep_moments"] = True
pos, vel = plasma_particles(spec, n, 2026, z_lo=0.02, z_hi=0.98)
sink, source = c1_sink_source(nr, nz)
sc = dict(spec=spec, position=pos, velocity=vel, sink_mask=sink, source_pdf=source, loops=scaled_loops(spec))
sim = makeCylindricalParticlePusher(spec)
apply_scene(sim, sc)
sim.emInit()
weight = 1e12 * np.pi * spec["radius"] ** 2 * spec["height"] / n   # 1e12 m^-3


def frame():
    sim.half_step(); sim.density(); sim.emStep(weight, True); sim.draw_canvas()


for _ in range(10):
    frame()
sim.sync()
sim.mark(0)
for _ in range(K):
    frame()
sim.mark(1)
ms = sim.elapsed_ms(0, 1) / K
sim.timing(True); sim.timing_reset()
for _ in range(K):
    frame()
sim.sync()
rs = 8 if precision == "f64" else 4
cells = nr * nz
alg = {"em_b": 9 * rs * cells, "em_e": 12 * rs * cells, "em_cells": 15 * rs * cells, "precalc": (6 + 8) * rs * cells}
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
peak = float(peaks.get("hbm_gbs", 6524.9) or 6524.9)
row = {"workload": f"EM frame: {n} particles, {nr}x{nz} grid, {precision}", "frame": "half_step + density + emStep + canvas draws",
       "ms_per_frame": round(ms, 4), "pushes_per_s": round(n / (ms * 1e-3), 0), "kernels": {}, "hbm_peak_gbps": peak}
for nm in ("push", "prepass", "permute", "scan", "index_scatter", "cellsum", "cellsum_warp", "cellsum_heavy", "conv",
           "em_b", "em_e", "em_cells", "precalc", "render"):
    t, c = sim.timing_get(nm)
    if c:
        k = {"per_launch_ms": round(t / c, 4), "per_frame_ms": round(t / K, 4)}
        if nm in alg:
            k["algorithmic_bytes"] = alg[nm]
            k["gbps"] = round(alg[nm] / (t / c * 1e-3) / 1e9, 1)
            k["frac_of_hbm_peak"] = round(k["gbps"] / peak, 3)
        row["kernels"][nm] = k
E = sim.getField("E")
row["finite"] = bool(np.isfinite(E).all())
print(json.dumps(row), flush=True)
