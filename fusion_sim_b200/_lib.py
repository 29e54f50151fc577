"""ctypes binding of libfusionsim.so (include/fusionsim.h).  No CPU fallback: if the CUDA
library is missing this raises, and fsim_create fails without a CUDA device."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FSIM_LIB_PATH: A/B measurement of another build of the same library (tools/ab_build.sh); never a fallback
LIB_PATH = os.environ.get("FSIM_LIB_PATH") or os.path.join(_HERE, "csrc", "libfusionsim.so")

FSIM_F64, FSIM_F32 = 0, 1
FLAG_CORRECTED_PREA, FLAG_KEEP_MOMENTS, FLAG_ATOMIC_DEPOSIT = 1, 2, 4
FLAG_POST_STREAM, FLAG_UNFUSED_SORT, FLAG_POST_NO_PRIORITY, FLAG_PERIODIC_Z = 8, 16, 64, 128
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE, ERR_RANGE = 1, 2, 3, 4, 5


class Error(Exception):
    """The reference reports every failure with `throw new Error(msg)`; so does this driver."""

    def __init__(self, msg, code=0):
        super().__init__(msg)
        self.code = code


class FsimSpec(C.Structure):
    _fields_ = [
        ("radius", C.c_double), ("height", C.c_double), ("nr", C.c_int64), ("nz", C.c_int64),
        ("dt", C.c_double), ("nparticles", C.c_int64), ("particle_mass", C.c_double),
        ("particle_charge", C.c_double),
        ("precision", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32),
        ("sort_interval", C.c_int32), ("nparticles_total", C.c_int64), ("capacity", C.c_int64),
        ("slab_row0", C.c_int64), ("slab_rows", C.c_int64), ("halo_rows", C.c_int64),
        ("id_base", C.c_uint64),
    ]


_P = C.c_void_p
_SIGS = {
    "fsim_last_error": (C.c_char_p, []),
    "fsim_abi_version": (C.c_int, []),
    "fsim_create": (C.c_int, [C.POINTER(FsimSpec), C.POINTER(_P)]),
    "fsim_destroy": (C.c_int, [_P]),
    "fsim_set_E": (C.c_int, [_P, _P]),
    "fsim_set_B": (C.c_int, [_P, _P]),
    "fsim_set_position": (C.c_int, [_P, _P]),
    "fsim_set_velocity": (C.c_int, [_P, _P]),
    "fsim_set_sink_mask": (C.c_int, [_P, _P]),
    "fsim_set_source_pdf": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "fsim_set_rand": (C.c_int, [_P, _P]),
    "fsim_set_entropy": (C.c_int, [_P, _P]),
    "fsim_set_inv_cdf": (C.c_int, [_P, _P]),
    "fsim_set_particle_count": (C.c_int, [_P, C.c_int64]),
    "fsim_set_ids": (C.c_int, [_P, _P]),
    "fsim_add_current_loop": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "fsim_add_current_z": (C.c_int, [_P, C.c_double]),
    "fsim_add_bz": (C.c_int, [_P, C.c_double]),
    "fsim_add_btheta": (C.c_int, [_P, C.c_double]),
    "fsim_add_spindle_cusp_plasma_field": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "fsim_get_spindle": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "fsim_precalc": (C.c_int, [_P]),
    "fsim_step": (C.c_int, [_P]),
    "fsim_half_step": (C.c_int, [_P]),
    "fsim_density": (C.c_int, [_P]),
    "fsim_render_rgba8": (C.c_int, [_P, _P]),
    "fsim_render_rgba8_async": (C.c_int, [_P, _P]),
    "fsim_render_rows_async": (C.c_int, [_P, _P]),
    "fsim_draw_canvas": (C.c_int, [_P]),
    "fsim_sort": (C.c_int, [_P]),
    "fsim_sync": (C.c_int, [_P]),
    "fsim_particle_count": (C.c_int64, [_P]),
    "fsim_local_cells": (C.c_int64, [_P]),
    "fsim_get_position": (C.c_int, [_P, _P]),
    "fsim_get_velocity": (C.c_int, [_P, _P]),
    "fsim_get_rand": (C.c_int, [_P, _P]),
    "fsim_get_ids": (C.c_int, [_P, _P]),
    "fsim_get_cells": (C.c_int, [_P, _P]),
    "fsim_get_field": (C.c_int, [_P, C.c_char_p, _P]),
    "fsim_get_cell_count": (C.c_int, [_P, _P]),
    "fsim_get_sink_mask": (C.c_int, [_P, _P]),
    "fsim_timing_enable": (C.c_int, [_P, C.c_int]),
    "fsim_timing_reset": (C.c_int, [_P]),
    "fsim_timing_get": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "fsim_launch_count": (C.c_int64, [_P]),
    "fsim_mark": (C.c_int, [_P, C.c_int]),
    "fsim_elapsed_ms": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "fsim_set_stream": (C.c_int, [_P, _P]),
    "fsim_migrate_record_bytes": (C.c_int64, [_P]),
    "fsim_migrate_pack": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.POINTER(_P)]),
    "fsim_migrate_unpack": (C.c_int, [_P, _P, C.c_int64]),
    "fsim_halo_ptrs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                 C.POINTER(C.c_int64)]),
    "fsim_jacobi_create": (C.c_int, [C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(_P)]),
    "fsim_jacobi_destroy": (C.c_int, [_P]),
    "fsim_jacobi_vec_length": (C.c_int64, [_P]),
    "fsim_jacobi_set_matrix": (C.c_int, [_P, _P]),
    "fsim_jacobi_set_b": (C.c_int, [_P, _P]),
    "fsim_jacobi_init_vector": (C.c_int, [_P, _P]),
    "fsim_jacobi_solve": (C.c_int, [_P, C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(C.c_int32), _P]),
    "fsim_jacobi_get_result": (C.c_int, [_P, _P]),
    "fsim_jacobi_launch_count": (C.c_int64, [_P]),
    "fsim_jacobi_timing": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "fsim_set_state": (C.c_int, [_P, _P, _P, _P]),
    "fsim_set_field": (C.c_int, [_P, C.c_char_p, _P]),
    "fsim_solve_fields": (C.c_int, [_P, C.c_double, C.c_int32, C.c_double, C.c_int32]),
    "fsim_solve_fields_stage": (C.c_int, [_P, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_int32]),
    "fsim_run_frames": (C.c_int, [_P, C.c_int64]),
    "fsim_frame_graph_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fsim_em_init": (C.c_int, [_P]),
    "fsim_em_set": (C.c_int, [_P, C.c_char_p, _P]),
    "fsim_em_get": (C.c_int, [_P, C.c_char_p, _P]),
    "fsim_em_step": (C.c_int, [_P, C.c_double, C.c_int32]),
    "fsim_field_rows": (C.c_int, [_P, C.c_char_p, C.c_int64, C.c_int64, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "fsim_cellsum_ptrs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(_P), C.POINTER(C.c_int64)]),
    "fsim_migrate_setup": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, C.POINTER(_P), C.POINTER(_P), _P, _P]),
    "fsim_migrate_begin": (C.c_int, [_P]),
    "fsim_migrate_end": (C.c_int, [_P]),
    "fsim_migrate_stats": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "fsim_density_interior": (C.c_int, [_P]),
    "fsim_check_digest": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "fsim_density_begin": (C.c_int, [_P]),
    "fsim_density_end": (C.c_int, [_P]),
}

_lib = None


def lib():
    """The loaded library.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Error(f"{LIB_PATH} is missing: build it with `python -m fusion_sim_b200.build` "
                        "(nvcc, sm_100a). There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise Error(lib().fsim_last_error().decode("utf-8", "replace"), rc)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)
