"""Scene construction: the reference's default demo (config C1) and its scaled synthetic
relatives (configs C2-C5 of BASELINE.json), from a seeded generator.

C1 restates the scene of public/javascripts/fusionsim.js:72-148 exactly; only the random source
changes (the reference uses unseeded Math.random / window.crypto):
NumPy PCG64(seed) draws, in this order, positions (3N), velocities (3N), rand (4N) and the
entropy table (4*1024^2); positions, velocities and rand uniform in [0,1), the entropy table as
the reference builds it (uint32 / 0xFFFFFFFF stored into a Float32Array, empic.js:143-155).
"""
from __future__ import annotations

import numpy as np

N_ENTROPY = 1024

C1_SPEC = dict(radius=1, height=2, nr=400, nz=800, dt=2e-9, nparticles=400,
               particle_mass=1.67e-27, particle_charge=1.602e-19)  # fusionsim.js:74-83


def entropy_table(rng):
    """[1024*1024][4] entropy texels, empic.js:143-155: `random_bytes[k] / 0xFFFFFFFF` stored into a
    Float32Array -- every entry is exactly a float (the fp64 engine then gathers a half-size copy)."""
    u = rng.integers(0, 1 << 32, size=(N_ENTROPY * N_ENTROPY, 4), dtype=np.uint64)
    return (u.astype(np.float64) / float(0xFFFFFFFF)).astype(np.float32).astype(np.float64)


def seeded_rand_entropy(seed: int, n: int):
    """rand [n][4] in [0,1) and the entropy table [1024*1024][4] for `spec.seed`."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rand = rng.random((n, 4))
    entropy = entropy_table(rng)
    return rand, entropy


def c1_sink_source(nr: int, nz: int):
    """Sink mask and source pdf of fusionsim.js:94-122, at the same proportions for any grid
    (nr=400, nz=800 reproduces the reference indices exactly)."""
    sink = np.ones((nr, nz))
    source = np.zeros((nr, nz))
    sink[nr - 1, :] = 0  # :105-108  (the axis column 0 stays open, :106 is commented out)
    sink[1:nr - 1, 0] = 0  # :109-112  (corners [0][0], [0][nz-1] stay 1)
    sink[1:nr - 1, nz - 1] = 0
    source[0:(50 * nr) // 400, (350 * nz) // 800:(450 * nz) // 800] = 1.0  # :116-122
    return sink, source


def c1_scene(seed: int = 12345, spec: dict | None = None):
    """The default demo scene (fusionsim.js:72-148)."""
    spec = dict(C1_SPEC if spec is None else spec)
    n = int(spec.get("nparticles_total", 0)) or int(spec["nparticles"]) ** 2
    rng = np.random.Generator(np.random.PCG64(seed))
    up = rng.random((n, 3))
    uv = rng.random((n, 3))
    rand = rng.random((n, 4))
    entropy = entropy_table(rng)
    # fusionsim.js:126-127
    position = 0.2 * (up - 0.5)
    position[:, 2] += 1
    velocity = 0.002 * (uv - 0.5)
    sink, source = c1_sink_source(int(spec["nr"]), int(spec["nz"]))
    return dict(spec=spec, position=position, velocity=velocity, sink_mask=sink, source_pdf=source,
                rand=rand, entropy=entropy,
                loops=[(0.8, 2.0, -10000000.0), (0.8, 0.0, 10000000.0)])  # :137-138


def scaled_spec(nr: int, nz: int, n: int, **ext):
    """Spec of a scaled scene: same cell size (2.5 mm) and time step as C1, so a particle moves the
    same fraction of a cell per half-step; N need not be a perfect square (nparticles_total)."""
    side = int(round(n ** 0.5))
    spec = dict(radius=nr / 400.0, height=nz / 400.0, nr=nr, nz=nz, dt=2e-9, nparticles=side,
                particle_mass=1.67e-27, particle_charge=1.602e-19)
    if side * side != n:
        spec["nparticles_total"] = n
    spec.update(ext)
    return spec


def scaled_loops(spec):
    """Two opposing loops (a cusp) at the C1 proportions, current scaled with the radius."""
    R, H = spec["radius"], spec["height"]
    return [(0.8 * R, H, -1.0e7 * R), (0.8 * R, 0.0, 1.0e7 * R)]


def plasma_particles(spec, n: int, seed: int, z_lo: float = 0.02, z_hi: float = 0.98,
                     r_lo: float = 0.02, r_hi: float = 0.98, dtype=np.float64):
    """Synthetic plasma for C2-C5: uniform in (r, z) over the interior of the domain (uniform
    occupancy per grid cell), random azimuth, velocity components uniform in +-0.001 c
    (fusionsim.js:127).  z_lo/z_hi are fractions of the height (a slab for multi-GPU runs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R, H = spec["radius"], spec["height"]
    pos = np.empty((n, 3), dtype)
    vel = np.empty((n, 3), dtype)
    chunk = 1 << 22
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        m = b - a
        r = (r_lo + (r_hi - r_lo) * rng.random(m)) * R
        phi = 2 * np.pi * rng.random(m)
        pos[a:b, 0] = r * np.cos(phi)
        pos[a:b, 1] = r * np.sin(phi)
        pos[a:b, 2] = (z_lo + (z_hi - z_lo) * rng.random(m)) * H
        vel[a:b] = 0.002 * (rng.random((m, 3)) - 0.5)
    return pos, vel


def apply_scene(sim, scene: dict, precalc: bool = True):
    """Drive a simulation object (product or oracle: same method names) through the init
    sequence of fusionsim.js:130-148."""
    value = {k: scene[k] for k in ("position", "velocity", "sink_mask", "source_pdf", "rand", "entropy",
                                   "E", "B", "inv_cdf") if scene.get(k) is not None}
    sim.set(value)
    for (r, z, I) in scene.get("loops", []):
        sim.addCurrentLoop(r, z, I)
    for name, fn in (("current_z", "addCurrentZ"), ("bz", "addBZ"), ("btheta", "addBTheta")):
        if scene.get(name) is not None:
            getattr(sim, fn)(scene[name])
    if precalc:
        sim.precalc()
    return sim
