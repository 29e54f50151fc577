"""Device time of the field-solve kernels (csrc/fieldsolve.cu, EXTENSION row N4) on one B200:
python tools/bench_fieldsolve.py > gpurun_out/fieldsolve.json
relax4 = 4 weighted-Jacobi sweeps per launch on TMA-staged tiles; algorithmic bytes per launch =
3 reals per cell (phi in, src in, phi out)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusion_sim_b200 import makeCylindricalParticlePusher  # noqa: E402

peak = 6524.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
out = []
for precision in ("f64", "f32"):
    for nr, nz in ((8192, 2048), (2048, 2048)):
        spec = dict(radius=nr * 0.0025, height=nz * 0.0025, nr=nr, nz=nz, dt=2e-9, nparticles=512,
                    particle_mass=1.67e-27, particle_charge=1.602e-19, precision=precision)
        g = makeCylindricalParticlePusher(spec)
        rng = np.random.default_rng(0)
        g.set({"position": np.stack([rng.random(512 * 512) * 0.5 * spec["radius"], np.zeros(512 * 512),
                                      rng.random(512 * 512) * spec["height"]], 1),
               "velocity": np.zeros((512 * 512, 3))})
        g.addBZ(0.1)
        g.precalc()
        g.density()
        v = {"macro_weight": 1e9, "sweeps": 8}
        g.solveFields(v)  # warm-up, allocation
        g.sync()
        g.timing(True); g.timing_reset()
        g.solveFields({"macro_weight": 1e9, "sweeps": 87})  # 21 x relax4 + relax2 + relax1
        g.sync()
        rs = 8 if precision == "f64" else 4
        cells = nr * nz
        row = {"precision": precision, "grid": [nr, nz]}
        for name, nbytes in (("relax4", 3 * rs), ("relax2", 3 * rs), ("relax1", 3 * rs), ("charge_source", 2 * rs),
                             ("efield", 4 * rs), ("precalc", 14 * rs)):
            ms, n = g.timing_get(name)
            if n:
                gbs = nbytes * cells / (ms / n * 1e-3) / 1e9
                row[name] = {"ms_per_launch": ms / n, "launches": n, "algorithmic_bytes_per_cell": nbytes,
                             "achieved_GBps": gbs, "frac_of_measured_peak": gbs / peak}
        ms4, n4 = g.timing_get("relax4")
        row["ms_per_sweep_at_T4"] = ms4 / n4 / 4
        row["cell_sweeps_per_s"] = cells * 4 / (ms4 / n4 * 1e-3)
        g.timing(False)
        out.append(row)
        del g
print(json.dumps({"peak_GBps": peak, "results": out}, indent=1))
