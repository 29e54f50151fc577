"""Host mirror of ``matrix_webgl.makeSORIterative(spec)`` (public/javascripts/matrix_webgl.js:35),
backed by the CUDA dense weighted-Jacobi routine of libfusionsim.so (csrc/jacobi.cu).

Same members as the reference object: ``vec_length, vec_height, set_matrix, set_b, init_vector,
solve, x_result_tex`` (the last returns the current solution vector instead of a GL texture).
``spec.literal = True`` reproduces two defects of the reference (see include/fusionsim.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Error, check, lib, ptr
from .pusher import validate_object


class SORIterative:
    def __init__(self, spec: dict):
        validate_object(spec, {"n_power": "number"})  # matrix_webgl.js:36-40
        self.n_power = int(spec["n_power"])
        self.vec_height = 2 ** self.n_power
        self.vec_length = 4 * self.vec_height * self.vec_height
        prec = {"f64": 0, "f32": 1}[spec.get("precision", "f64")]
        flags = 1 if spec.get("literal") else 0
        self._h = C.c_void_p()
        check(lib().fsim_jacobi_create(self.n_power, float(spec.get("relaxation") or 0.0), prec,
                                       int(spec.get("device", 0)), flags, C.byref(self._h)))

    def destroy(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().fsim_jacobi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _vec(self, a, n):
        a = np.ascontiguousarray(np.asarray(a, np.float64))
        if a.size != n:
            raise Error(f"array has {a.size} elements, expected {n}", _lib.ERR_INVALID)
        return a

    def set_matrix(self, matrix):
        check(lib().fsim_jacobi_set_matrix(self._h, ptr(self._vec(matrix, self.vec_length ** 2))))
        return self

    def set_b(self, b):
        check(lib().fsim_jacobi_set_b(self._h, ptr(self._vec(b, self.vec_length))))
        return self

    def init_vector(self, vector):
        check(lib().fsim_jacobi_init_vector(self._h, ptr(self._vec(vector, self.vec_length))))
        return self

    def solve(self, params: dict):
        validate_object(params, {"tolerance": "number"})  # matrix_webgl.js:577-581
        corr, diff, it = C.c_double(), C.c_double(), C.c_int32()
        result = np.empty(self.vec_length, np.float64)
        # `iteration < params.max_iterations` with max_iterations undefined is false: no iterations
        check(lib().fsim_jacobi_solve(self._h, float(params["tolerance"]), int(params.get("substep") or 0),
                                      int(params.get("max_iterations") or 0), C.byref(corr), C.byref(diff),
                                      C.byref(it), ptr(result)))
        return {"correlation": corr.value, "diff": diff.value, "iterations": it.value, "result": result}

    def x_result_tex(self):
        out = np.empty(self.vec_length, np.float64)
        check(lib().fsim_jacobi_get_result(self._h, ptr(out)))
        return out

    def timing(self):
        ms, n = C.c_double(), C.c_int64()
        check(lib().fsim_jacobi_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def makeSORIterative(spec: dict) -> SORIterative:
    """matrix_webgl.makeSORIterative(spec), matrix_webgl.js:35."""
    return SORIterative(spec)
