"""TEST INFRASTRUCTURE ONLY -- CPU oracle of matrix_webgl.makeSORIterative
(public/javascripts/matrix_webgl.js:35-711), front end of oracle/fsim_oracle_jacobi_impl.h.
The literal mode is held bit for bit to the outputs of the reference's own shader text
(tests/test_reference_glsl.py); the reference has no tests and its only live caller does not run."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from .oracle import _p, lib, tofixed20


class OracleSOR:
    def __init__(self, spec: dict, nthreads: int = 1):
        self.n_power = int(spec["n_power"])
        self.vec_height = vh = 2 ** self.n_power
        self.vec_length = L = 4 * vh * vh
        self.omega = spec.get("relaxation") or 1.0
        self.literal = bool(spec.get("literal"))
        self.precision = spec.get("precision", "f64")
        self.dt = np.float64 if self.precision == "f64" else np.float32
        self.sfx = self.precision
        self.nthreads = nthreads
        self.A = np.zeros((L, L), self.dt)
        self.b = np.zeros(L, self.dt)
        self.R = np.zeros((L, L), self.dt)
        self.Cv = np.zeros(L, self.dt)
        self.x_guess = np.zeros(L, self.dt)
        self.x_result = np.zeros(L, self.dt)

    def _f(self, name):
        return getattr(lib(), f"{name}_{self.sfx}")

    def set_matrix(self, m):
        self.A[:] = np.asarray(m, np.float64).reshape(self.A.shape)
        return self

    def set_b(self, b):
        self.b[:] = np.asarray(b, np.float64)
        return self

    def init_vector(self, v):
        self.x_result[:] = np.asarray(v, np.float64)
        return self

    def mv_product(self):
        out = np.empty_like(self.x_result)
        self._f("orcj_mv_product")(C.c_int64(self.vec_height), _p(self.R), _p(self.Cv), _p(self.x_guess), _p(out),
                                   C.c_double(tofixed20(1.0 - self.omega)), C.c_int(1 if self.omega == 1.0 else 0),
                                   C.c_int(1 if self.literal else 0), C.c_int(self.nthreads))
        return out

    def solve(self, params: dict):
        L = self.vec_length
        self._f("orcj_setup")(C.c_int64(L), _p(self.A), _p(self.b), C.c_double(tofixed20(self.omega)),
                              C.c_int(1 if self.omega == 1.0 else 0), _p(self.R), _p(self.Cv), C.c_int(self.nthreads))
        tol = params["tolerance"]
        max_it = params.get("max_iterations") or 0
        sub = params.get("substep") or 1
        stats = np.zeros(L, self.dt)
        corr = 0.0
        x1 = x2 = x1x2 = x1x1 = x2x2 = 0.0
        diff = tol + 1
        it = 0
        while it < max_it and diff > tol:  # matrix_webgl.js:643
            for _ in range(sub):
                self.x_guess = self.x_result.copy()
                self.x_result = self.mv_product()
            self._f("orcj_stats")(C.c_int64(L // 4), _p(self.x_guess), _p(self.x_result), _p(stats))
            if not self.literal:
                x1 = x2 = x1x2 = x1x1 = x2x2 = 0.0
            max_diff = 0.0
            a1, a2, st = self.x_guess.astype(np.float64), self.x_result.astype(np.float64), stats.astype(np.float64)
            for i in range(L // 4):  # :675-683, doubles like the JS loop
                x1 += a1[4 * i] + a1[4 * i + 1] + a1[4 * i + 2] + a1[4 * i + 3]
                x2 += a2[4 * i] + a2[4 * i + 1] + a2[4 * i + 2] + a2[4 * i + 3]
                x1x2 += st[4 * i]
                x1x1 += st[4 * i + 1]
                x2x2 += st[4 * i + 2]
                max_diff = max(max_diff, st[4 * i + 3])
            with np.errstate(all="ignore"):
                den = np.float64((L * x1x1 - x1 * x1) * (L * x2x2 - x2 * x2))
                corr = float(np.float64(L * x1x2 - x1 * x2) / np.sqrt(den))
                diff = float(np.float64(2 * L * max_diff) / np.float64(abs(x1) + abs(x2)))
            it += 1
        return {"correlation": corr, "diff": diff, "iterations": it, "result": self.x_result.astype(np.float64)}
