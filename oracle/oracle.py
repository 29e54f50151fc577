"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/fsim_oracle.c).

The reference (kcdodd/fusion-sim, public/javascripts/empic.js) has no tests or golden vectors and
cannot run headless as a whole; this oracle restates its shaders and is held bit for bit to the
outputs of executing the reference's own GLSL text (tests/test_reference_glsl.py).

``OraclePusher`` mirrors the reference's simulation object (empic.js:1157-1526: set,
addCurrentLoop, addCurrentZ, addBZ, addBTheta, precalc, step, density, canvas) so that a
parity test can drive it and the CUDA product with the same script.  Only tests/,
__graft_entry__.smoke() and bench.py (cpu_baseline, --impl reference) may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

C_LIGHT = 2.998e8  # empic.js:27
N_ENTROPY = 1024
FSIM_EPS0 = 8.8541878128e-12  # include/fsim_constants.h
FSIM_PI = 3.14159265358979323846
N_INVCDF = 512


def build(force: bool = False) -> str:
    """Build libfsim_oracle.so if it is missing or older than its sources.  Serialised with a file
    lock: the multi-rank tests import this module from several processes at once."""
    import fcntl
    so = os.path.join(_HERE, "libfsim_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("fsim_oracle.c", "fsim_oracle_impl.h", "fsim_oracle_jacobi_impl.h",
                                             "fsim_oracle_fields_impl.h", "fsim_oracle_spindle_impl.h", "fsim_oracle_em_impl.h", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "fsim_constants.h"))

    def stale():
        return (not os.path.exists(so)) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)

    if force or stale():
        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            if force or stale():  # another process may have built it while we waited
                subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_tofixed20.restype = C.c_double
        _LIB.orc_tofixed20.argtypes = [C.c_double]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def tofixed20(x: float) -> float:
    return float(lib().orc_tofixed20(float(x)))


def cos_table(as_f32: bool = False) -> np.ndarray:
    out = np.empty(1000, np.float64)
    lib().orc_cos_table(_p(out), C.c_int(1 if as_f32 else 0))
    return out


def shape_table(as_f32: bool) -> np.ndarray:
    out = np.empty(121, np.float64)
    lib().orc_shape_table(_p(out), C.c_int(1 if as_f32 else 0))
    return out


def inv_cdf(pdf: np.ndarray) -> np.ndarray:
    pdf = np.ascontiguousarray(pdf, np.float64)
    out = np.empty((N_INVCDF * N_INVCDF, 2), np.float64)
    rc = lib().orc_inv_cdf(_p(pdf), C.c_int64(pdf.shape[0]), C.c_int64(pdf.shape[1]), _p(out))
    if rc != 0:
        raise MemoryError
    return out


class OraclePusher:
    """CPU mirror of ``empic.makeCylindricalParticlePusher(spec)`` (empic.js:30)."""

    def __init__(self, spec: dict, nthreads: int = 1):
        for k in ("radius", "height", "nr", "nz", "dt", "nparticles", "particle_mass",
                  "particle_charge"):
            if k not in spec:
                raise ValueError("." + k + " <- Non-optional property is undefined!")
        self.spec = dict(spec)
        self.precision = spec.get("precision", "f64")
        self.dt = np.float64 if self.precision == "f64" else np.float32
        self.sfx = "f64" if self.precision == "f64" else "f32"
        self.nthreads = int(nthreads)
        self.nr, self.nz = int(spec["nr"]), int(spec["nz"])
        self.ncell = self.nr * self.nz
        side = int(spec["nparticles"])
        self.n = int(spec.get("nparticles_total", 0)) or side * side
        # physical quantities, empic.js:44-46
        self.h = spec["particle_charge"] * spec["dt"] / (2 * spec["particle_mass"])
        self.factor_r = 1 / spec["radius"]
        self.factor_z = 1 / spec["height"]
        self.step_factor = spec["dt"] * C_LIGHT  # empic.js:852
        self.corrected = bool(spec.get("corrected_preA", False))
        self.periodic = 1 if spec.get("periodic_z") else 0  # EXTENSION, no reference counterpart
        z4 = lambda n: np.zeros((n, 4), self.dt)
        self.position, self.velocity, self.rand = z4(self.n), z4(self.n), z4(self.n)
        self.entropy = z4(N_ENTROPY * N_ENTROPY)
        self.E, self.B = z4(self.ncell), z4(self.ncell)
        self.sink_mask = z4(self.ncell)
        self.inv_cdf = z4(N_INVCDF * N_INVCDF)
        self.R1, self.R2, self.R3, self.A = (z4(self.ncell) for _ in range(4))
        self.cell_sums = z4(self.ncell)
        self.cell_count = np.zeros(self.ncell, np.uint32)
        self.moments01, self.moments01_norm, self.moments01_avg = (z4(self.ncell) for _ in range(3))
        self.shape = shape_table(self.precision == "f32").astype(self.dt)
        self.costab = cos_table(self.precision == "f32").astype(self.dt)
        self._loop_tables = None
        self.last_cell = np.zeros(self.n, np.int64)
        self.deposit_cell = np.zeros(self.n, np.int64)
        seed = spec.get("seed")
        if seed is not None:
            from fusion_sim_b200.scenes import seeded_rand_entropy
            r, e = seeded_rand_entropy(int(seed), self.n)
            self.set({"rand": r, "entropy": e})

    def _f(self, name):
        return getattr(lib(), f"{name}_{self.sfx}")

    # -- set(), empic.js:1157-1350 -------------------------------------------------
    def _grid3(self, v):
        a = np.asarray(v, np.float64).reshape(self.nr, self.nz, 3)
        out = np.empty((self.ncell, 4), self.dt)
        out[:, :3] = a.transpose(1, 0, 2).reshape(self.ncell, 3)  # texel i + j*nr
        out[:, 3] = 1.0
        return out

    def set(self, value: dict):
        if value.get("E") is not None:
            self.E = self._grid3(value["E"])
        if value.get("B") is not None:
            self.B = self._grid3(value["B"])
        fac = np.array([self.factor_r, self.factor_r, self.factor_z])
        if value.get("position") is not None:
            p = np.asarray(value["position"], np.float64).reshape(self.n, 3) * fac
            self.position[:, :3] = p
            self.position[:, 3] = 1.0
        if value.get("velocity") is not None:
            v = np.asarray(value["velocity"], np.float64).reshape(self.n, 3) * fac
            self.velocity[:, :3] = v
            self.velocity[:, 3] = 1.0
        if value.get("sink_mask") is not None:
            s = np.asarray(value["sink_mask"], np.float64).reshape(self.nr, self.nz)
            self.sink_mask[:, 0] = s.T.reshape(-1)
        if value.get("source_pdf") is not None:
            t = inv_cdf(np.asarray(value["source_pdf"], np.float64))
            self.inv_cdf[:, :2] = t
        if value.get("inv_cdf") is not None:  # extension: raw table
            self.inv_cdf[:, :2] = np.asarray(value["inv_cdf"], np.float64).reshape(-1, 2)
        if value.get("rand") is not None:  # extension
            self.rand[:] = np.asarray(value["rand"], np.float64).reshape(self.n, 4)
        if value.get("entropy") is not None:
            self.entropy[:] = np.asarray(value["entropy"], np.float64).reshape(-1, 4)

    # -- static field builders, empic.js:1352-1411 -----------------------------------
    def _tables(self):
        if self._loop_tables is None:
            half = np.empty((self.ncell, 4), self.dt)
            tenth = np.empty((self.ncell, 4), self.dt)
            f = self._f("orc_loop_shape")
            f(C.c_int64(self.nr), C.c_int64(self.nz), C.c_double(0.5), _p(self.costab), _p(half),
              C.c_int(self.nthreads))
            f(C.c_int64(self.nr), C.c_int64(self.nz), C.c_double(0.1), _p(self.costab), _p(tenth),
              C.c_int(self.nthreads))
            self._loop_tables = (half, tenth)
        return self._loop_tables

    def addCurrentLoop(self, r, z, I):
        half, tenth = self._tables()
        self._f("orc_add_current_loop")(
            C.c_int64(self.nr), C.c_int64(self.nz), C.c_double(r * self.factor_r),
            C.c_double(z * self.factor_z), C.c_double(I), _p(half), _p(tenth), _p(self.B),
            C.c_int(self.nthreads))

    def _uniform(self, kind, val):
        self._f("orc_add_uniform")(C.c_int64(self.nr), C.c_int64(self.nz), C.c_int(kind),
                                   C.c_double(val), _p(self.B))

    def addCurrentZ(self, I):
        self._uniform(0, I)

    def addBZ(self, Bz):
        self._uniform(1, Bz)

    def addBTheta(self, Bt):
        self._uniform(2, Bt)

    def addSpindleCuspPlasmaField(self, r, B_c, beta_c=1.0):
        """Spindle cusp with a field-excluding plasma (specification: include/fusionsim.h,
        oracle/fsim_oracle_spindle_impl.h; written from the intent of spindle.js, which does not run)."""
        from .jacobi import OracleSOR
        L = 256
        sp = self.spec
        radius, height = float(sp["radius"]), float(sp["height"])
        nodes, points, normals = np.empty((L + 1, 2)), np.empty((L, 2)), np.empty((L, 2))
        lib().orcs_geometry(C.c_double(radius), C.c_double(height), _p(nodes), _p(points), _p(normals))
        cos64 = cos_table(False)
        coil_I = 2.0 * float(r) * float(B_c) / 1.25663706e-6
        A, rhs = np.empty((L, L)), np.empty(L)
        lib().orcs_assemble(C.c_double(radius), C.c_double(height), C.c_double(float(r)), C.c_double(coil_I), _p(cos64),
                            _p(nodes), _p(points), _p(normals), _p(A), _p(rhs), C.c_int(self.nthreads))
        sor = OracleSOR({"n_power": 3, "relaxation": 1.0}, nthreads=self.nthreads)
        res = sor.set_matrix(A).set_b(rhs).solve({"tolerance": 1e-9, "substep": 64, "max_iterations": 4000})
        cur = np.empty(L + 1)
        lib().orcs_node_currents(_p(np.ascontiguousarray(res["result"])), C.c_double(float(beta_c)), _p(cur))
        loops = [(float(r), 0.0, coil_I), (float(r), height, -coil_I)]
        for l in range(L + 1):
            loops.append((nodes[l, 0], nodes[l, 1], cur[l]))
            loops.append((nodes[l, 0], height - nodes[l, 1], -cur[l]))
        loops = np.ascontiguousarray(np.asarray(loops, np.float64))
        self._f("orcs_add_loops")(C.c_int64(self.nr), C.c_int64(self.nz), C.c_double(radius), C.c_double(height),
                                  C.c_int64(len(loops)), _p(loops), _p(self.costab), _p(self.B), C.c_int(self.nthreads))
        self.spindle = dict(nodes=nodes, points=points, normals=normals, A=A, rhs=rhs, x=res["result"], currents=cur,
                            iterations=res["iterations"], diff=res["diff"], coil_current=coil_I, loops=loops)
        return self.spindle

    # -- precalc(), empic.js:1413-1434 -------------------------------------------------
    def precalc(self):
        self._f("orc_precalc")(
            C.c_int64(self.ncell), C.c_double(self.h),
            C.c_double(tofixed20(self.factor_r / self.factor_z)),
            C.c_double(tofixed20(self.factor_z / self.factor_r)),
            C.c_double(tofixed20(self.factor_r)), C.c_double(tofixed20(self.factor_z)),
            _p(self.E), _p(self.B), _p(self.R1), _p(self.R2), _p(self.R3), _p(self.A),
            C.c_int(1 if self.corrected else 0), C.c_int(self.nthreads))

    # -- step(), empic.js:1436-1469 ------------------------------------------------------
    def half_step(self):
        lib().orc_set_periodic_z(C.c_int(self.periodic))
        self._f("orc_half_step")(
            C.c_int64(self.n), _p(self.position), _p(self.velocity), _p(self.rand),
            _p(self.entropy), _p(self.R1), _p(self.R2), _p(self.R3), _p(self.A),
            _p(self.sink_mask), _p(self.inv_cdf), C.c_int64(self.nr), C.c_int64(self.nz),
            C.c_double(self.step_factor), _p(self.last_cell), C.c_int(self.nthreads))

    def step(self):
        self.half_step()
        self.half_step()

    # -- density(), empic.js:1471-1505 ---------------------------------------------------
    def density(self, literal_sprites: bool = False, timing_mt: bool = False):
        lib().orc_set_periodic_z(C.c_int(self.periodic))
        if literal_sprites:
            self._f("orc_deposit_sprites")(
                C.c_int64(self.n), _p(self.position), _p(self.velocity), _p(self.shape),
                C.c_int64(self.nr), C.c_int64(self.nz), _p(self.moments01))
        else:
            if timing_mt:
                self._f("orc_cell_sums_mt")(
                    C.c_int64(self.n), _p(self.position), _p(self.velocity), C.c_int64(self.nr),
                    C.c_int64(self.nz), _p(self.cell_sums), _p(self.cell_count),
                    C.c_int(self.nthreads))
            else:
                self._f("orc_cell_sums")(
                    C.c_int64(self.n), _p(self.position), _p(self.velocity), C.c_int64(self.nr),
                    C.c_int64(self.nz), _p(self.cell_sums), _p(self.cell_count),
                    _p(self.deposit_cell))
            self._f("orc_convolve")(C.c_int64(self.nr), C.c_int64(self.nz), _p(self.cell_sums),
                                    _p(self.shape), _p(self.moments01), C.c_int(self.nthreads))
        self._f("orc_normalize_ema")(C.c_int64(self.nr), C.c_int64(self.nz), _p(self.moments01),
                                     _p(self.moments01_norm), _p(self.moments01_avg),
                                     C.c_int(self.nthreads))

    # -- EXTENSION: self-consistent electrostatic field solve ------------------
    def solveFields(self, value: dict):
        """{macro_weight, sweeps, omega=1, source="avg"|"instant"}: charge density from the
        deposited moments -> `sweeps` weighted-Jacobi sweeps on the potential (warm start) ->
        E = -grad(phi) -> precalc().  Specification: oracle/fsim_oracle_fields_impl.h."""
        sp = self.spec
        dr, dz = sp["radius"] / self.nr, sp["height"] / self.nz
        dens = self.moments01_avg if value.get("source", "avg") == "avg" else self.moments01_norm
        rho_scale = (sp["particle_charge"] * float(value["macro_weight"])
                     / (FSIM_PI * sp["radius"] * dr * dz * FSIM_EPS0))
        if not hasattr(self, "phi"):
            self.phi = np.zeros(self.ncell, self.dt)
            self.rho_src = np.zeros(self.ncell, self.dt)
        lib().orc_set_periodic_z(C.c_int(self.periodic))
        if not hasattr(self, "background"):
            self.background = np.zeros(self.ncell, self.dt)
        self._f("orc_charge_source_bg")(C.c_int64(self.ncell), _p(dens), C.c_double(rho_scale), _p(self.background),
                                        _p(self.rho_src))
        coef = np.empty((self.nr, 4), np.float64)
        lib().orc_relax_coeffs(C.c_int64(self.nr), C.c_double(dr), C.c_double(dz), _p(coef))
        tmp = np.empty_like(self.phi)
        self._f("orc_relax")(C.c_int64(self.nr), C.c_int64(self.nz), _p(self.phi), _p(tmp), _p(self.rho_src),
                             _p(coef), C.c_double(float(value.get("omega", 1.0))), C.c_int(int(value["sweeps"])),
                             C.c_int(self.nthreads))
        self._f("orc_efield")(C.c_int64(self.nr), C.c_int64(self.nz), _p(self.phi), C.c_double(1 / (2 * dr)),
                              C.c_double(1 / (2 * dz)), _p(self.E))
        self.precalc()

    # -- EXTENSION (BASELINE configs[2]): axisymmetric Yee update driven by the deposited current ----
    EM_SHAPES = {"Er": (1, 0), "Ez": (0, 1), "Bt": (0, 0), "Et": (1, 1), "Br": (0, 1), "Bz": (1, 0)}  # (extra rows, extra columns)

    def emInit(self):
        """Zero Yee fields; the static field present now stays underneath (specification: fsim_oracle_em_impl.h)."""
        self.em = {k: np.zeros((self.nz + dj) * (self.nr + di), self.dt) for k, (dj, di) in self.EM_SHAPES.items()}
        self.B0 = self.B.copy()
        sp = self.spec
        dr, dz = sp["radius"] / self.nr, sp["height"] / self.nz
        if C_LIGHT * sp["dt"] * np.sqrt(1 / dr ** 2 + 1 / dz ** 2) >= 1.0:
            raise RuntimeError(".dt <- the Yee update needs c dt sqrt(1/dr^2 + 1/dz^2) < 1")

    def emSet(self, name, data):
        self.em[name][:] = np.asarray(data, np.float64).reshape(-1)

    def emGet(self, name):
        return self.em[name].astype(np.float64)

    def emStep(self, macro_weight=0.0, with_current=True):
        sp = self.spec
        coef, scal = np.empty((self.nr + 1, 6)), np.empty(6)
        lib().orce_coeffs(C.c_int64(self.nr), C.c_int64(self.nz), C.c_double(sp["radius"]), C.c_double(sp["height"]),
                          C.c_double(sp["dt"]), C.c_double(sp["particle_charge"]), C.c_double(float(macro_weight)), _p(coef), _p(scal))
        e = self.em
        self._f("orce_step")(C.c_int64(self.nr), C.c_int64(self.nz), _p(e["Er"]), _p(e["Ez"]), _p(e["Bt"]), _p(e["Et"]), _p(e["Br"]),
                             _p(e["Bz"]), _p(coef), _p(scal), _p(self.moments01) if with_current else None, C.c_int(self.nthreads))
        self._f("orce_cells")(C.c_int64(self.nr), C.c_int64(self.nz), _p(e["Er"]), _p(e["Ez"]), _p(e["Bt"]), _p(e["Et"]), _p(e["Br"]),
                              _p(e["Bz"]), _p(self.B0), _p(self.E), _p(self.B))
        self.precalc()

    @property
    def canvas(self) -> np.ndarray:
        out = np.empty((self.nz, self.nr, 4), np.uint8)
        self._f("orc_render_mt")(C.c_int64(self.nr), C.c_int64(self.nz), _p(self.B),
                                 _p(self.moments01_avg), _p(out), C.c_int(self.nthreads))
        return out

    # -- accessors (extension; same names as the product) --------------------------------
    def getPosition(self):
        return self.position.astype(np.float64)

    def getVelocity(self):
        return self.velocity[:, :3].astype(np.float64)

    def getRand(self):
        return self.rand.astype(np.float64)

    def getField(self, name):
        a = getattr(self, name)
        if name == "sink_mask":
            return (a[:, 0] > 0.5).astype(np.uint8)
        if name == "inv_cdf":
            return a[:, :2].astype(np.float64)
        if name == "cell_count":
            return a.copy()
        if name in ("phi", "rho_src"):
            return a.astype(np.float64)
        if name in ("E", "B", "R1", "R2", "R3", "A"):
            return a[:, :3].astype(np.float64)
        return a.astype(np.float64)
