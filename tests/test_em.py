"""EXTENSION (BASELINE.json configs[2] "Boris + Yee FDTD"; no reference counterpart -- the
reference's E and B are static): electromagnetic field update on an axisymmetric Yee mesh, driven by the deposited
current.  PARITY UNPINNED by construction; what is tested:

  * the CPU oracle (oracle/fsim_oracle_em_impl.h) against a second, separately written NumPy statement of the update,
    and against the PHYSICS: the TM010 and TE011 resonances of the empty conducting cylinder
    (w = 2.405 c / R and w = c sqrt((3.832/R)^2 + (pi/H)^2)), energy that stays bounded, and -- the test of the
    particle coupling, Boris push -> deposit -> J -> Ampere -- the upward shift of the TM010 resonance in a cavity
    filled with cold electrons, w^2 = w_c^2 + w_p^2, within 1 % on a 64 x 32 grid and converging with the cell size
    (the vacuum value would be 20 % off);
  * on the GPU, libfusionsim.so against that oracle bit for bit (fields and particles) in fp64 and fp32, through
    frames of half_step() + density() + emStep(), and the same plasma resonance measured on the product.
"""
import numpy as np
import pytest

from conftest import assert_same

EPS0, QE, ME, CL = 8.8541878128e-12, 1.602e-19, 9.109e-31, 2.998e8
X01, X11 = 2.404825557695773, 3.831705970207512   # first zeros of J0 and J1
NAMES = ("Er", "Ez", "Bt", "Et", "Br", "Bz")


def cavity_spec(nr=32, nz=16, R=0.05, H=0.025, courant=0.5, n=4, precision="f64", **kw):
    dr, dz = R / nr, H / nz
    dt = courant / (CL * np.sqrt(1 / dr ** 2 + 1 / dz ** 2))
    return dict(radius=R, height=H, nr=nr, nz=nz, dt=dt, nparticles=1, nparticles_total=n, particle_mass=ME,
                particle_charge=-QE, precision=precision, **kw)


def shapes(nr, nz):
    return {"Er": (nz + 1, nr), "Ez": (nz, nr + 1), "Bt": (nz, nr), "Et": (nz + 1, nr + 1), "Br": (nz, nr + 1), "Bz": (nz + 1, nr)}


def numpy_step(f, sp, J=None):
    """Second statement of the update (vectorised, its own association): f = dict of 2-D arrays, in place."""
    nr, nz, dt = sp["nr"], sp["nz"], sp["dt"]
    dr, dz = sp["radius"] / nr, sp["height"] / nz
    ri, rh = np.arange(nr + 1) * dr, (np.arange(nr) + 0.5) * dr
    Er, Ez, Bt, Et, Br, Bz = (f[k] for k in NAMES)
    Br += dt * (Et[1:, :] - Et[:-1, :]) / dz
    Bt -= dt * ((Er[1:, :] - Er[:-1, :]) / dz - (Ez[:, 1:] - Ez[:, :-1]) / dr)
    Bz -= dt * (ri[1:] * Et[:, 1:] - ri[:-1] * Et[:, :-1]) / (rh * dr)
    c2 = CL * CL
    dEr = -c2 * (Bt[1:, :] - Bt[:-1, :]) / dz
    dEt = c2 * ((Br[1:, 1:-1] - Br[:-1, 1:-1]) / dz - (Bz[1:-1, 1:] - Bz[1:-1, :-1]) / dr)
    dEz = c2 * (rh[1:] * Bt[:, 1:] - rh[:-1] * Bt[:, :-1]) / (ri[1:-1] * dr)
    ax = c2 * 4 * Bt[:, 0] / dr
    if J is not None:
        Jr, Jt, Jz = J
        dEr = dEr - 0.5 * (Jr[1:, :] + Jr[:-1, :]) / EPS0
        dEt = dEt - 0.25 * (Jt[1:, 1:] + Jt[1:, :-1] + Jt[:-1, 1:] + Jt[:-1, :-1]) / EPS0
        dEz = dEz - 0.5 * (Jz[:, 1:] + Jz[:, :-1]) / EPS0
        ax = ax - Jz[:, 0] / EPS0
    Er[1:-1, :] += dt * dEr
    Et[1:-1, 1:-1] += dt * dEt
    Ez[:, 1:-1] += dt * dEz
    Ez[:, 0] += dt * ax


def angular_frequency(signal, dt):
    s = np.asarray(signal)
    z = np.nonzero(np.diff(np.sign(s)))[0]
    t = np.array([(i + s[i] / (s[i] - s[i + 1])) * dt for i in z])
    return 2 * np.pi / (2 * np.mean(np.diff(t)))


def j0(x):
    from scipy.special import j0 as f
    return f(x)


def j1(x):
    from scipy.special import j1 as f
    return f(x)


def oracle_for(spec):
    from oracle.oracle import OraclePusher
    return OraclePusher(spec, nthreads=4)


def random_fields(nr, nz, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    f = {k: rng.standard_normal(s) for k, s in shapes(nr, nz).items()}
    for k in ("Bt", "Br", "Bz"):
        f[k] /= CL  # comparable energy in E and B
    return f


# ------------------------------------------------------------------------------------------------ CPU: the oracle

def test_oracle_equals_the_numpy_statement():
    sp = cavity_spec(nr=24, nz=20)
    o = oracle_for(sp)
    o.emInit()
    f = random_fields(24, 20, 1)
    # honour the conductor: tangential E on the walls and E_theta on the axis are zero
    f["Ez"][:, -1] = 0; f["Et"][:, -1] = 0; f["Et"][0, :] = 0; f["Et"][-1, :] = 0; f["Er"][0, :] = 0; f["Er"][-1, :] = 0
    f["Et"][:, 0] = 0
    for k in NAMES:
        o.emSet(k, f[k])
    rng = np.random.Generator(np.random.PCG64(2))
    o.moments01[:] = rng.standard_normal(o.moments01.shape) * 1e-6
    mom = o.moments01.reshape(20, 24, 4)
    w = 3e9
    u = (np.arange(24) + 0.5) / 24
    dr, dz = sp["radius"] / 24, sp["height"] / 20
    G = -QE * w * 1000 * CL / (2 * np.pi * u[None, :] * sp["radius"] * dr * dz)
    J = (G * mom[:, :, 0] * sp["radius"], G * mom[:, :, 1] * sp["radius"], G * mom[:, :, 2] * sp["height"])
    for _ in range(5):
        o.emStep(w, True)
        numpy_step(f, sp, J)
    for k in NAMES:
        got, want = o.emGet(k), f[k].reshape(-1)
        assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max(), k
    # the fields the push gathers: edge values averaged to the cell centres
    E = o.getField("E").reshape(20, 24, 3)
    B = o.getField("B").reshape(20, 24, 3)
    np.testing.assert_allclose(E[:, :, 0], 0.5 * (f["Er"][:-1] + f["Er"][1:]), rtol=1e-10, atol=1e-14 * np.abs(f["Er"]).max())
    np.testing.assert_allclose(E[:, :, 2], 0.5 * (f["Ez"][:, :-1] + f["Ez"][:, 1:]), rtol=1e-10, atol=1e-14 * np.abs(f["Ez"]).max())
    np.testing.assert_allclose(B[:, :, 1], f["Bt"], rtol=1e-10, atol=1e-14 * np.abs(f["Bt"]).max())


def test_static_field_stays_underneath():
    sp = cavity_spec(nr=8, nz=8)
    o = oracle_for(sp)
    o.addBZ(0.3)
    o.emInit()
    o.emStep(0.0, False)
    B = o.getField("B").reshape(8, 8, 3)
    assert np.all(B[:, :, 2] == 0.3) and not B[:, :, :2].any() and not o.getField("E").any()


def test_vacuum_resonances_of_the_cavity():
    nr, nz = 32, 16
    sp = cavity_spec(nr, nz)
    R, H, dt = sp["radius"], sp["height"], sp["dt"]
    ri = np.arange(nr + 1) * R / nr
    # TM010: E_z = J0(2.405 r / R), uniform in z
    o = oracle_for(sp)
    o.emInit()
    Ez = np.tile(j0(X01 * ri / R), (nz, 1)); Ez[:, -1] = 0
    o.emSet("Ez", Ez)
    sig, energy = [], []
    vol_c = 2 * np.pi * ((np.arange(nr) + 0.5) * R / nr) * (R / nr) * (H / nz)
    for _ in range(1200):
        o.emStep(0.0, False)
        sig.append(o.emGet("Ez").reshape(nz, nr + 1)[nz // 2, 0])
        E, B = o.getField("E").reshape(nz, nr, 3), o.getField("B").reshape(nz, nr, 3)
        energy.append((0.5 * EPS0 * ((E ** 2).sum(2) + CL * CL * (B ** 2).sum(2)) * vol_c[None, :]).sum())
    w = angular_frequency(sig, dt)
    assert abs(w / (X01 * CL / R) - 1) < 2e-3, w / (X01 * CL / R)
    assert not o.emGet("Et").any() and not o.emGet("Br").any()     # the TE set is never excited
    # E and B are half a step apart, so the cell-centred sum wobbles at 2w -- but it neither grows nor decays
    e = np.array(energy)
    assert e[600:].max() <= 1.05 * e[:600].max() and e[600:].min() >= 0.95 * e[:600].min()
    # TE011: E_theta = J1(3.832 r / R) sin(pi z / H)
    o = oracle_for(sp)
    o.emInit()
    zj = np.arange(nz + 1) * H / nz
    Et = np.sin(np.pi * zj / H)[:, None] * j1(X11 * ri / R)[None, :]
    Et[:, -1] = 0; Et[0, :] = 0; Et[-1, :] = 0
    o.emSet("Et", Et)
    sig = []
    for _ in range(1200):
        o.emStep(0.0, False)
        sig.append(o.emGet("Et").reshape(nz + 1, nr + 1)[nz // 2, nr // 2])
    w = angular_frequency(sig, dt)
    want = CL * np.sqrt((X11 / R) ** 2 + (np.pi / H) ** 2)
    assert abs(w / want - 1) < 5e-3, w / want
    assert not o.emGet("Ez").any() and not o.emGet("Bt").any()


def plasma_scene(n, nr, nz, seed=1, precision="f64", wp_over_wc=0.75):
    sp = cavity_spec(nr, nz, n=n, precision=precision, keep_moments=True)
    R, H = sp["radius"], sp["height"]
    wc = X01 * CL / R
    wp = wp_over_wc * wc
    n0 = wp * wp * EPS0 * ME / QE ** 2
    rng = np.random.Generator(np.random.PCG64(seed))
    r = R * 0.9999 * np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    z = H * (0.001 + 0.998 * rng.random(n))
    source = np.zeros((nr, nz)); source[0:4, :] = 1
    scene = dict(position=np.stack([r * np.cos(th), r * np.sin(th), z], 1), velocity=np.zeros((n, 3)),
                 sink_mask=np.ones((nr, nz)), source_pdf=source, rand=rng.random((n, 4)), entropy=rng.random((1024 * 1024, 4)))
    Ez = 1e3 * np.tile(j0(X01 * np.arange(nr + 1) / nr), (nz, 1)); Ez[:, -1] = 0
    return sp, scene, dict(weight=n0 * np.pi * R * R * H / n, Ez=Ez, want=np.sqrt(wc * wc + wp * wp), vacuum=wc)


def run_plasma(sim, scene, info, frames, nr, nz):
    sim.set(scene)
    sim.emInit()
    sim.emSet("Ez", info["Ez"])
    sim.emStep(0.0, False)   # cell-centred fields of the initial state (one vacuum step)
    sig = []
    for _ in range(frames):
        sim.half_step()
        sim.density()
        sim.emStep(info["weight"], True)
        sig.append(sim.emGet("Ez").reshape(nz, nr + 1)[nz // 2, 0])
    return sig


def test_cold_plasma_shifts_the_resonance():
    """Boris push -> sprite deposit -> J -> Ampere: w^2 = w_c^2 + w_p^2 in a cavity filled with cold electrons."""
    err = []
    for (nr, nz, n, frames) in ((32, 16, 1 << 16, 500), (64, 32, 1 << 17, 900)):
        sp, scene, info = plasma_scene(n, nr, nz)
        o = oracle_for(sp)
        sig = run_plasma(o, scene, info, frames, nr, nz)
        err.append(angular_frequency(sig[60:], sp["dt"]) / info["want"] - 1)
    # the vacuum resonance would read -0.20.  The 11-cell sprite smooths J against the J0 profile and loses charge to
    # the wall cells, which lowers the coupling on a coarse grid: the error shrinks with the cell size
    assert abs(err[0]) < 0.04 and abs(err[1]) < 0.01 and abs(err[1]) < 0.5 * abs(err[0]), err


def test_courant_limit_is_enforced():
    o = oracle_for(cavity_spec(courant=1.01))
    with pytest.raises(RuntimeError, match="c dt"):
        o.emInit()


# ------------------------------------------------------------------------------------------------ GPU: the product

@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_gpu_vacuum_update_equals_the_oracle(precision):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    nr, nz = 40, 24   # not a multiple of anything
    sp = cavity_spec(nr, nz, precision=precision)
    g, o = makeCylindricalParticlePusher(sp), oracle_for(sp)
    for s in (g, o):
        s.addBZ(0.2)
        s.addCurrentLoop(0.03, 0.01, 1e4)
        s.emInit()
    f = random_fields(nr, nz, 5)
    for k in NAMES:
        g.emSet(k, f[k]); o.emSet(k, f[k])
    for _ in range(25):
        g.emStep(0.0, False); o.emStep(0.0, False)
    for k in NAMES:
        assert_same(g.emGet(k), o.emGet(k), k)
    assert_same(g.getField("E"), o.getField("E"), "cell-centred E")
    assert_same(g.getField("B"), o.getField("B"), "cell-centred B0 + B")
    assert_same(g.getField("R1"), o.getField("R1"), "Boris rows")


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_gpu_self_consistent_frames_equal_the_oracle(precision):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    nr, nz, n = 32, 16, 1 << 14
    sp, scene, info = plasma_scene(n, nr, nz, seed=7, precision=precision)
    g, o = makeCylindricalParticlePusher(sp), oracle_for(sp)
    f = random_fields(nr, nz, 9)
    for s in (g, o):
        s.set(scene)
        s.addBZ(0.05)
        s.emInit()
        for k in NAMES:
            s.emSet(k, 1e2 * f[k])
        s.emStep(0.0, False)
    for frame in range(12):
        for s in (g, o):
            s.half_step()
            s.density()
            s.emStep(info["weight"], True)
    assert_same(g.getField("moments01"), o.getField("moments01"), "deposited moments")
    for k in NAMES:
        assert_same(g.emGet(k), o.emGet(k), k)
    assert_same(g.getPosition(), o.getPosition(), "positions")
    assert_same(g.getVelocity(), o.getVelocity(), "velocities")


@pytest.mark.gpu
def test_gpu_cold_plasma_shifts_the_resonance():
    from fusion_sim_b200 import makeCylindricalParticlePusher
    nr, nz = 64, 32
    sp, scene, info = plasma_scene(1 << 20, nr, nz)
    g = makeCylindricalParticlePusher(sp)
    sig = run_plasma(g, scene, info, 900, nr, nz)
    w = angular_frequency(sig[100:], sp["dt"])
    assert abs(w / info["want"] - 1) < 0.015, (w / info["want"], info["vacuum"] / info["want"])


@pytest.mark.gpu
def test_gpu_em_arguments_are_validated():
    from fusion_sim_b200 import Error, makeCylindricalParticlePusher
    g = makeCylindricalParticlePusher(cavity_spec(8, 8))
    with pytest.raises(Error, match="fsim_em_init"):
        g.emStep(0.0, False)
    g.emInit()
    with pytest.raises(Error, match="KEEP_MOMENTS"):
        g.emStep(1.0, True)
    with pytest.raises(Error, match="unknown field"):
        g.emSet("Ex", np.zeros(64))
    with pytest.raises(Error, match="c dt"):
        makeCylindricalParticlePusher(cavity_spec(8, 8, courant=1.2)).emInit()
    with pytest.raises(Error, match="one GPU"):
        makeCylindricalParticlePusher(cavity_spec(8, 8, periodic_z=True)).emInit()
