"""Headless host driver: the simulation object of the reference, backed by libfusionsim.so.

``makeCylindricalParticlePusher(spec)`` mirrors ``empic.makeCylindricalParticlePusher``
(public/javascripts/empic.js:30) and returns an object with the reference's ten members
(empic.js:60, :1157-:1526): ``canvas, set, addCurrentLoop, addSpindleCuspPlasmaField,
addCurrentZ, addBZ, addBTheta, precalc, step, density`` -- same names, argument meaning, units
and error behaviour (a synchronous ``Error``).  Beside them sit the accessors the reference
lacks and the seeding inputs (row 4).

All arithmetic happens in the CUDA library; this file only re-shapes arrays.  There is no
CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import numbers

import numpy as np

from . import _lib
from ._lib import Error, FsimSpec, check, lib, ptr

_REQUIRED = ("radius", "height", "nr", "nz", "dt", "nparticles", "particle_mass", "particle_charge")
_EXT = ("precision", "device", "flags", "sort_interval", "nparticles_total", "capacity", "slab_row0",
        "slab_rows", "halo_rows", "id_base", "seed", "corrected_preA", "keep_moments", "periodic_z")


def validate_object(test: dict, control: dict):
    """utilities.validate_object / validate_property (utilities.js:11-127) for 'number' controls."""
    for prop, kind in control.items():
        v = test.get(prop) if isinstance(test, dict) else getattr(test, prop, None)
        if v is None:
            raise Error("." + prop + " <- Non-optional property is undefined!", _lib.ERR_INVALID)
        if kind == "number" and (isinstance(v, bool) or not isinstance(v, numbers.Real)):
            raise Error("." + prop + " <- Property does not match any given possible types!",
                        _lib.ERR_INVALID)


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        if a.size != int(np.prod(shape)):
            raise Error(f"array has {a.size} elements, expected shape {tuple(shape)}", _lib.ERR_INVALID)
        a = a.reshape(shape)
    return a


class CylindricalParticlePusher:
    """Object returned by :func:`makeCylindricalParticlePusher`."""

    def __init__(self, spec: dict):
        validate_object(spec, {k: "number" for k in _REQUIRED})
        self.spec = dict(spec)
        cs = FsimSpec()
        for k in _REQUIRED:
            setattr(cs, k, spec[k])
        prec = spec.get("precision", "f64")
        cs.precision = {"f64": 0, "f32": 1, 0: 0, 1: 1}[prec]
        flags = int(spec.get("flags", 0))
        if spec.get("corrected_preA"):
            flags |= _lib.FLAG_CORRECTED_PREA
        if spec.get("keep_moments"):
            flags |= _lib.FLAG_KEEP_MOMENTS
        if spec.get("periodic_z"):  # EXTENSION: z periodic; the reference has no such boundary
            flags |= _lib.FLAG_PERIODIC_Z
        cs.flags = flags
        # periodic z: the engine's cell tables carry ghost rows either side of the nz owned rows; the accessors of
        # this mirror return / take the owned rows only
        self.periodic = bool(flags & _lib.FLAG_PERIODIC_Z)
        self.ghost_rows = max(int(spec.get("halo_rows", 0)), 8) if self.periodic else 0
        for k in ("device", "sort_interval", "nparticles_total", "capacity", "slab_row0", "slab_rows",
                  "halo_rows", "id_base"):
            setattr(cs, k, int(spec.get(k, 0)))
        self.precision = "f64" if cs.precision == 0 else "f32"
        self.nr, self.nz = int(spec["nr"]), int(spec["nz"])
        self._h = C.c_void_p()
        check(lib().fsim_create(C.byref(cs), C.byref(self._h)))
        self.ncell_local = int(lib().fsim_local_cells(self._h))
        seed = spec.get("seed")
        if seed is not None:
            from .scenes import seeded_rand_entropy
            r, e = seeded_rand_entropy(int(seed), self.n)
            self.set({"rand": r, "entropy": e})

    # -- lifetime ---------------------------------------------------------------------------
    def destroy(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().fsim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def n(self) -> int:
        return int(lib().fsim_particle_count(self._h))

    # -- out.set(value), empic.js:1157-1350 -------------------------------------------------
    def set(self, value: dict):
        L, h = lib(), self._h
        if value.get("E") is not None:
            check(L.fsim_set_E(h, ptr(_f64(value["E"], (self.nr, self.nz, 3)))))
        if value.get("B") is not None:
            check(L.fsim_set_B(h, ptr(_f64(value["B"], (self.nr, self.nz, 3)))))
        if value.get("position") is not None:
            check(L.fsim_set_position(h, ptr(_f64(value["position"], (self.n, 3)))))
        if value.get("velocity") is not None:
            check(L.fsim_set_velocity(h, ptr(_f64(value["velocity"], (self.n, 3)))))
        if value.get("sink_mask") is not None:
            check(L.fsim_set_sink_mask(h, ptr(_f64(value["sink_mask"], (self.nr, self.nz)))))
        if value.get("source_pdf") is not None:
            pdf = _f64(value["source_pdf"])
            if pdf.ndim != 2:
                raise Error(".source_pdf <- must be a 2-D array", _lib.ERR_INVALID)
            check(L.fsim_set_source_pdf(h, ptr(pdf), pdf.shape[0], pdf.shape[1]))
        # extensions (seeding; raw inverse-cdf table)
        if value.get("rand") is not None:
            check(L.fsim_set_rand(h, ptr(_f64(value["rand"], (self.n, 4)))))
        if value.get("entropy") is not None:
            check(L.fsim_set_entropy(h, ptr(_f64(value["entropy"], (1024 * 1024, 4)))))
        if value.get("inv_cdf") is not None:
            check(L.fsim_set_inv_cdf(h, ptr(_f64(value["inv_cdf"], (512 * 512, 2)))))

    # -- static field builders, empic.js:1352-1411 -----------------------------------------
    def addCurrentLoop(self, r, z, I):
        check(lib().fsim_add_current_loop(self._h, r, z, I))

    def addSpindleCuspPlasmaField(self, r, B_c, beta_c=1.0):
        """empic.js:1369.  The reference's version does not run; this one implements its intent (a spindle cusp
        of two opposing coils of radius r, B_c tesla at a coil's centre, and the surface currents that exclude
        the field from the plasma, scaled by 1 - sqrt(1 - beta_c)): specification in include/fusionsim.h.
        Returns what the boundary solve found."""
        check(lib().fsim_add_spindle_cusp_plasma_field(self._h, float(r), float(B_c), float(beta_c)))
        x, cur, A, rhs = np.empty(256), np.empty(257), np.empty((256, 256)), np.empty(256)
        it, diff = C.c_int32(), C.c_double()
        check(lib().fsim_get_spindle(self._h, ptr(x), ptr(cur), ptr(A), ptr(rhs), C.byref(it), C.byref(diff)))
        return dict(x=x, currents=cur, A=A, rhs=rhs, iterations=it.value, diff=diff.value)

    def addCurrentZ(self, I):
        check(lib().fsim_add_current_z(self._h, I))

    def addBZ(self, Bz):
        check(lib().fsim_add_bz(self._h, Bz))

    def addBTheta(self, Btheta):
        check(lib().fsim_add_btheta(self._h, Btheta))

    # -- precalc / step / density / canvas, empic.js:1413-1505 --------------------------------
    def precalc(self):
        check(lib().fsim_precalc(self._h))

    def step(self):
        check(lib().fsim_step(self._h))

    def half_step(self):
        check(lib().fsim_half_step(self._h))

    def density(self):
        check(lib().fsim_density(self._h))

    def run_frames(self, nframes: int):
        """`nframes` iterations of the page loop (fusionsim.js:170-178): step(), density() with its two canvas draws --
        bit-identical to calling them one by one; launch-bound scenes replay a captured CUDA graph of 2 x sort_interval
        frames (include/fusionsim.h, fsim_run_frames)."""
        check(lib().fsim_run_frames(self._h, int(nframes)))

    def frame_graph_info(self) -> dict:
        f, l, r = C.c_int32(), C.c_int64(), C.c_int64()
        check(lib().fsim_frame_graph_info(self._h, C.byref(f), C.byref(l), C.byref(r)))
        return {"frames_per_cycle": f.value, "launches_per_cycle": l.value, "replays": r.value}

    def solveFields(self, value: dict):
        """EXTENSION (no reference counterpart): close the PIC loop.
        value = {macro_weight, sweeps, omega=1, source="avg"|"instant"}: charge density from the
        deposited moments -> `sweeps` weighted-Jacobi sweeps on the potential (warm start) ->
        E = -grad(phi) -> precalc().  Specification: include/fusionsim.h (fsim_solve_fields)."""
        validate_object(value, {"macro_weight": "number", "sweeps": "number"})
        source = value.get("source", "avg")
        if source not in ("avg", "instant"):
            raise Error(".source <- 'avg' or 'instant'", _lib.ERR_INVALID)
        check(lib().fsim_solve_fields(self._h, float(value["macro_weight"]), int(value["sweeps"]),
                                      float(value.get("omega", 1.0)), 0 if source == "avg" else 1))

    # -- EXTENSION (no reference counterpart, SURVEY.md section 8f N4 / BASELINE configs[2]): Yee field update --------
    EM_SHAPES = {"Er": (1, 0), "Ez": (0, 1), "Bt": (0, 0), "Et": (1, 1), "Br": (0, 1), "Bz": (1, 0)}  # (extra rows, extra columns)

    def _em_shape(self, name):
        if name not in self.EM_SHAPES:
            raise Error(f".name <- unknown field {name!r} (Er Ez Bt Et Br Bz)", _lib.ERR_INVALID)
        dj, di = self.EM_SHAPES[name]
        return (self.nz + dj, self.nr + di)

    def emInit(self):
        """Zero electromagnetic fields on an axisymmetric Yee mesh; the B present now stays underneath as the static
        field.  Specification: include/fusionsim.h (fsim_em_step)."""
        check(lib().fsim_em_init(self._h))

    def emSet(self, name: str, data):
        shape = self._em_shape(name)
        a = np.ascontiguousarray(data, np.float64)
        if a.size != shape[0] * shape[1]:
            raise Error(f".{name} <- expected {shape[0]} x {shape[1]} values", _lib.ERR_INVALID)
        check(lib().fsim_em_set(self._h, name.encode(), ptr(a)))

    def emGet(self, name: str) -> np.ndarray:
        out = np.empty(self._em_shape(name), np.float64)
        check(lib().fsim_em_get(self._h, name.encode(), ptr(out)))
        return out.reshape(-1)

    def emStep(self, macro_weight: float = 0.0, with_current: bool = True):
        """One leap-frog step of dt: B from curl E, E from curl B and the deposited current (moments01 of the last
        density(); needs keep_moments), cell-centred E and B for the push, precalc().  Pair with half_step()."""
        check(lib().fsim_em_step(self._h, float(macro_weight), 1 if with_current else 0))

    def render(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.nz, self.nr, 4), np.uint8)
        check(lib().fsim_render_rgba8(self._h, ptr(out)))
        return out

    def render_async(self, out: np.ndarray) -> np.ndarray:
        """Same image, copied to `out` (pinned memory) on a second stream; complete after sync()."""
        check(lib().fsim_render_rgba8_async(self._h, ptr(out)))
        return out

    def render_rows_async(self, out: np.ndarray) -> np.ndarray:
        """A slab rank's own rows of the image, [slab_rows][nr][4], into `out` (pinned) on the copy stream."""
        check(lib().fsim_render_rows_async(self._h, ptr(out)))
        return out

    def draw_canvas(self):
        """The two canvas draws of out.density (empic.js:1497-1504) into the device-resident canvas."""
        check(lib().fsim_draw_canvas(self._h))

    @property
    def canvas(self) -> np.ndarray:
        """RGBA8 image [nz][nr][4], top row first: what `drawImage(simulation.canvas)` shows."""
        return self.render()

    def sort(self):
        check(lib().fsim_sort(self._h))

    def sync(self):
        check(lib().fsim_sync(self._h))

    # -- accessors (extension) -----------------------------------------------------------------
    def _get(self, fn, shape, dtype=np.float64):
        out = np.empty(shape, dtype)
        check(fn(self._h, ptr(out)))
        return out

    def getPosition(self):
        """[N][4] normalised x, y, z and the alive flag (position.w of the reference)."""
        return self._get(lib().fsim_get_position, (self.n, 4))

    def getVelocity(self):
        return self._get(lib().fsim_get_velocity, (self.n, 3))

    def getRand(self):
        return self._get(lib().fsim_get_rand, (self.n, 4))

    def getIds(self):
        return self._get(lib().fsim_get_ids, (self.n,), np.uint64)

    def getCells(self):
        return self._get(lib().fsim_get_cells, (self.n,), np.int64)

    def _owned(self, a):
        """Periodic z: strip the ghost rows of a cell-indexed array [local cells][...]."""
        if not self.periodic:
            return a
        g = self.ghost_rows * self.nr
        return np.ascontiguousarray(a[g:g + self.nr * self.nz])

    def _with_ghosts(self, a):
        """Periodic z: owned rows [nz*nr][...] -> the engine's local table with the wrapped rows as ghosts."""
        if not self.periodic:
            return a
        g = self.ghost_rows * self.nr
        return np.ascontiguousarray(np.concatenate([a[-g:], a, a[:g]]))

    def setBackground(self, background):
        """EXTENSION: the neutralising background solveFields() subtracts from the density, [nz*nr] in the units
        of the density texture (e.g. getField("moments01_norm")[:, 3] at t = 0: immobile ions)."""
        b = self._with_ghosts(_f64(background, (self.nr * self.nz,)))
        check(lib().fsim_set_field(self._h, b"background", ptr(b)))

    def getField(self, name: str):
        if self.periodic and name not in ("sink_mask", "inv_cdf", "entropy"):
            self.periodic = False
            try:
                a = self.getField(name)
            finally:
                self.periodic = True
            return self._owned(a)
        nc = self.ncell_local
        if name == "cell_count":
            return self._get(lib().fsim_get_cell_count, (nc,), np.uint32)
        if name == "sink_mask":
            return self._get(lib().fsim_get_sink_mask, (self.nr * self.nz,), np.uint8)
        shape = {"E": (nc, 3), "B": (nc, 3), "R1": (nc, 3), "R2": (nc, 3), "R3": (nc, 3), "A": (nc, 3),
                 "cell_sums": (nc, 4), "moments01": (nc, 4), "moments01_norm": (nc, 4),
                 "moments01_avg": (nc, 4), "inv_cdf": (512 * 512, 2), "entropy": (1024 * 1024, 4),
                 "phi": (nc,), "rho_src": (nc,)}.get(name)
        if shape is None:
            raise Error("unknown field name: " + name, _lib.ERR_INVALID)
        out = np.empty(shape, np.float64)
        check(lib().fsim_get_field(self._h, name.encode(), ptr(out)))
        return out

    # -- checkpoint / restore (extension: the reference keeps its state in closure-private textures) ----
    def checkpoint(self) -> dict:
        """Everything a run needs to continue bit for bit: particle state (normalised units, alive
        flags, RNG state), fields, running average, potential."""
        ck = {"position": self.getPosition(), "velocity": self.getVelocity(), "rand": self.getRand(),
              "E": self.getField("E"), "B": self.getField("B"), "moments01_avg": self.getField("moments01_avg")}
        try:
            ck["phi"] = self.getField("phi")
        except Error:
            pass  # no solveFields() yet
        return ck

    def restore(self, ck: dict):
        """Inverse of checkpoint() on a simulation created with the same spec and the same static
        tables (sink mask, source pdf, entropy); runs precalc()."""
        L, h = lib(), self._h
        nc = self.ncell_local
        cells = lambda a: np.ascontiguousarray(np.asarray(a, np.float64).reshape(self.nz, self.nr, 3).transpose(1, 0, 2))
        check(L.fsim_set_E(h, ptr(cells(ck["E"]))))
        check(L.fsim_set_B(h, ptr(cells(ck["B"]))))
        check(L.fsim_precalc(h))
        check(L.fsim_set_state(h, ptr(_f64(ck["position"], (self.n, 4))), ptr(_f64(ck["velocity"], (self.n, 3))),
                               ptr(_f64(ck["rand"], (self.n, 4)))))
        nown = self.nr * self.nz if self.periodic else nc
        check(L.fsim_set_field(h, b"moments01_avg", ptr(self._with_ghosts(_f64(ck["moments01_avg"], (nown, 4))))))
        if ck.get("phi") is not None:
            check(L.fsim_set_field(h, b"phi", ptr(self._with_ghosts(_f64(ck["phi"], (nown,))))))

    def setState(self, position4=None, velocity3=None, rand4=None):
        """Raw particle state in normalised units (see fsim_set_state)."""
        p = ptr(_f64(position4, (self.n, 4))) if position4 is not None else None
        v = ptr(_f64(velocity3, (self.n, 3))) if velocity3 is not None else None
        r = ptr(_f64(rand4, (self.n, 4))) if rand4 is not None else None
        check(lib().fsim_set_state(self._h, p, v, r))

    # -- measurement hooks ---------------------------------------------------------------------
    def timing(self, on: bool):
        check(lib().fsim_timing_enable(self._h, 1 if on else 0))

    def timing_reset(self):
        check(lib().fsim_timing_reset(self._h))

    def timing_get(self, name: str):
        ms, n = C.c_double(), C.c_int64()
        check(lib().fsim_timing_get(self._h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def mark(self, slot: int):
        check(lib().fsim_mark(self._h, slot))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        check(lib().fsim_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def check_digest(self) -> dict:
        """Run invariants reduced on the device (fsim_check_digest): live particles, XOR and sum of
        their ids, particles deposited on the owned rows, sum of the weight channel."""
        out = np.zeros(4, np.uint64)
        a = C.c_double()
        check(lib().fsim_check_digest(self._h, ptr(out), C.byref(a)))
        return {"particles": int(out[0]), "id_xor": int(out[1]), "id_sum": int(out[2]),
                "deposited": int(out[3]), "sum_alpha": a.value}

    @property
    def launch_count(self) -> int:
        return int(lib().fsim_launch_count(self._h))


def makeCylindricalParticlePusher(spec: dict) -> CylindricalParticlePusher:
    """empic.makeCylindricalParticlePusher(spec), empic.js:30."""
    return CylindricalParticlePusher(spec)
