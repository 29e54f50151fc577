"""Device time of the dense weighted-Jacobi mat-vec (csrc/jacobi.cu) on one B200:
python tools/bench_jacobi.py > gpurun_out/jacobi.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusion_sim_b200.matrix import makeSORIterative  # noqa: E402

peak = 6524.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
out = []
for precision in ("f64", "f32"):
    for n_power in (4, 5, 6):
        L = 4 * (2 ** n_power) ** 2
        rng = np.random.default_rng(0)
        A = rng.random((L, L)) - 0.5
        A[np.arange(L), np.arange(L)] = 2 * np.abs(A).sum(1)
        b = rng.random(L)
        g = makeSORIterative({"n_power": n_power, "precision": precision}).set_matrix(A).set_b(b)
        g.solve({"tolerance": -1.0, "substep": 3, "max_iterations": 1})  # warm-up
        ms0, n0 = g.timing()
        r = g.solve({"tolerance": -1.0, "substep": 10, "max_iterations": 3})
        ms1, n1 = g.timing()
        ms = (ms1 - ms0) / (n1 - n0)
        rs = 8 if precision == "f64" else 4
        gbs = L * L * rs / (ms * 1e-3) / 1e9
        out.append({"kernel": "jacobi_mv_kernel", "precision": precision, "n_power": n_power, "unknowns": L,
                    "matrix_bytes": L * L * rs, "ms_per_matvec": ms, "achieved_GBps": gbs, "peak_GBps": peak,
                    "frac": gbs / peak, "residual_inf": float(np.abs(A @ r["result"] - b).max())})
        g.destroy()
print(json.dumps(out, indent=1))
