"""EXTENSION (no reference counterpart -- the reference clamps its textures and absorbs
at the end walls): periodic z.  A pushed position wraps before the sink lookup, the 11x11 deposit footprint wraps,
the Poisson stencil and the gradient of the field solve wrap.  PARITY UNPINNED by construction; what is tested:

  * the oracle's periodic mode against its defining properties (CPU);
  * the TWO-STREAM INSTABILITY (BASELINE.json configs[1] names it): two cold counter-streaming electron beams in a
    periodic box on a neutralising background.  Cold-beam electrostatic dispersion, two equal beams of plasma
    frequency w_b each: w^2 = k^2 v0^2 + w_b^2 - w_b sqrt(4 k^2 v0^2 + w_b^2), fastest growth w_b / 2 at
    k v0 = (sqrt(3)/2) w_b -- independent of the radial structure for unmagnetised beams.  The measured growth rate
    of the k = 2 pi / height mode of E_z must lie within 10 % of w_b / 2 (oracle on the CPU, product on the GPU);
  * on the GPU, libfusionsim.so against the oracle bit for bit through self-consistent periodic frames.
"""
import numpy as np
import pytest

from conftest import assert_same

EPS0, QE, ME, CL = 8.8541878128e-12, 1.602e-19, 9.109e-31, 2.998e8


def two_stream_scene(n=1 << 17, nr=32, nz=64, seed=3, precision="f64"):
    R = H = 0.05
    n0, dt = 1e14, 2e-10
    wb = np.sqrt(n0 * QE * QE / (EPS0 * ME) / 2)          # plasma frequency of ONE beam
    k = 2 * np.pi / H
    v0 = np.sqrt(3) / 2 * wb / k                          # fastest-growing mode = the box length
    spec = dict(radius=R, height=H, nr=nr, nz=nz, dt=dt, nparticles=1, nparticles_total=n, particle_mass=ME,
                particle_charge=-QE, keep_moments=True, periodic_z=True, precision=precision)
    rng = np.random.Generator(np.random.PCG64(seed))
    r = R * 0.999 * np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    z = H * rng.random(n)
    z = (z + 1e-4 * H * np.sin(k * z)) % H                # seed the mode
    sgn = np.where(np.arange(n) % 2 == 0, 1.0, -1.0)
    pos = np.stack([r * np.cos(th), r * np.sin(th), z], 1)
    vel = np.stack([np.zeros(n), np.zeros(n), sgn * v0 / CL], 1)
    source = np.zeros((nr, nz))
    source[0:4, :] = 1
    scene = dict(position=pos, velocity=vel, sink_mask=np.ones((nr, nz)), source_pdf=source, rand=rng.random((n, 4)),
                 entropy=rng.random((1024 * 1024, 4)))
    weight = n0 * np.pi * R * R * H / n
    return spec, scene, dict(wb=wb, dt=dt, weight=weight, nr=nr, nz=nz)


def start(sim, scene):
    sim.set(scene)
    sim.precalc()
    sim.density()
    return sim.getField("moments01_norm")[:, 3].copy()    # immobile ions: the electrons' own density at t = 0


def growth_rate(amps, dt):
    la = np.log(np.asarray(amps))
    i1 = int(np.argmax(la > la[5:30].mean() + 2.0))       # out of the noise ...
    i2 = int(np.argmax(la > la.max() - 1.5))              # ... and before saturation
    assert i2 - i1 >= 15, (i1, i2)
    return np.polyfit(np.arange(i1, i2) * dt, la[i1:i2], 1)[0]


def mode_amplitude(E, nr, nz):
    Ez = np.asarray(E).reshape(nz, nr, -1)[:, :, 2]
    a = np.abs(np.fft.fft(Ez, axis=0)[1, :])
    return float(np.sqrt((a[4:nr - 4] ** 2).mean()))


def test_periodic_oracle_properties():
    from oracle.oracle import OraclePusher
    spec, scene, p = two_stream_scene(n=20000, nz=32)
    o = OraclePusher(spec, nthreads=2)
    o.set(scene); o.precalc()
    for _ in range(40):                                   # 0.35 cell per step: every particle wraps several times
        o.half_step()
    z = o.position[:, 2]
    assert z.min() >= 0 and z.max() < 1 and (o.position[:, 3] == 1).all()      # nothing absorbed at the ends
    o.density()
    # the deposit wraps: every particle is deposited with its full weight (no sprite is clipped at z = 0, height)
    inner = (np.hypot(o.position[:, 0], o.position[:, 1]) * p["nr"] >= 6) & (np.hypot(o.position[:, 0], o.position[:, 1]) * p["nr"] < p["nr"] - 6)
    m = o.moments01.reshape(p["nz"], p["nr"], 4)[:, :, 3]
    assert m.sum() >= 0.001 * inner.sum()
    # rolling the particles by 5 rows rolls the (periodic) moments by 5 rows -- up to the re-ordered sums
    o2 = OraclePusher(spec, nthreads=2)
    o2.set(scene); o2.precalc()
    o2.position[:] = o.position
    o2.position[:, 2] = (o.position[:, 2] + 5.0 / p["nz"]) % 1.0
    o2.velocity[:] = o.velocity
    o2.density()
    m2 = o2.moments01.reshape(p["nz"], p["nr"], 4)[:, :, 3]
    np.testing.assert_allclose(np.roll(m, 5, axis=0), m2, rtol=1e-9, atol=1e-15)


def test_two_stream_growth_rate_oracle():
    from oracle.oracle import OraclePusher
    spec, scene, p = two_stream_scene()
    o = OraclePusher(spec, nthreads=4)
    o.background = start(o, scene)
    val = dict(macro_weight=p["weight"], sweeps=200, omega=1.0, source="instant")
    amps = []
    for _ in range(260):
        o.half_step(); o.density(); o.solveFields(val)
        amps.append(mode_amplitude(o.E, p["nr"], p["nz"]))
    g = growth_rate(amps, p["dt"])
    assert abs(g / (p["wb"] / 2) - 1) < 0.10, g / (p["wb"] / 2)   # linear theory: w_b / 2
    assert max(amps) > 100 * amps[0]


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_gpu_periodic_frames_equal_the_oracle(precision):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from oracle.oracle import OraclePusher
    spec, scene, p = two_stream_scene(n=40000, nz=40, precision=precision)   # 40 rows: not a multiple of any tile
    g = makeCylindricalParticlePusher(spec)
    o = OraclePusher(spec, nthreads=4)
    bg, bo = start(g, scene), start(o, scene)
    assert_same(bg, bo, "initial density")
    g.setBackground(bg); o.background = bo.astype(o.dt)
    val = dict(macro_weight=p["weight"], sweeps=23, omega=0.9, source="instant")
    for f in range(12):
        if f % 2:
            g.step(); o.step()
        else:
            g.half_step(); o.half_step()
        g.density(); o.density()
        g.solveFields(val); o.solveFields(val)
        assert_same(g.getPosition(), o.getPosition(), f"frame {f} position")
        assert_same(g.getVelocity(), o.getVelocity(), f"frame {f} velocity")
        for nm in ("cell_count", "moments01", "moments01_norm", "moments01_avg", "phi", "E", "A"):
            assert_same(g.getField(nm), o.getField(nm), f"frame {f} {nm}")
    assert_same(g.canvas, o.canvas, "canvas")
    z = g.getPosition()[:, 2]
    assert z.min() >= 0 and z.max() < 1 and np.abs(o.getField("phi")).max() > 0


@pytest.mark.gpu
def test_gpu_two_stream_growth_rate():
    """BASELINE configs[1] in the reference's geometry: measured growth within 10 % of linear theory."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    spec, scene, p = two_stream_scene(n=1 << 20)          # 1 Mi particles, as configs[1] names
    g = makeCylindricalParticlePusher(spec)
    g.setBackground(start(g, scene))
    val = dict(macro_weight=p["weight"], sweeps=200, omega=1.0, source="instant")
    amps = []
    for _ in range(260):
        g.half_step(); g.density(); g.solveFields(val)
        amps.append(mode_amplitude(g.getField("E"), p["nr"], p["nz"]))
    rate = growth_rate(amps, p["dt"])
    assert abs(rate / (p["wb"] / 2) - 1) < 0.10, rate / (p["wb"] / 2)
    assert max(amps) > 100 * amps[0]
