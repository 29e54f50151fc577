"""CPU tests of the oracle itself (no GPU).

The reference has no tests, golden vectors or fixtures and cannot run headless (SURVEY.md
section 8c: PARITY UNPINNED).  What stands in: (i) the C oracle against a second, independently
written NumPy restatement of the shaders, bit for bit; (ii) analytic invariants of the physics
the shaders encode; (iii) committed golden outputs of the oracle (tests/golden) so that it
cannot drift unnoticed."""
import os

import numpy as np
import pytest

from conftest import assert_same, small_scene
from oracle import numpy_ref as nref
from oracle import oracle as orc
from oracle.oracle import OraclePusher
from fusion_sim_b200.scenes import apply_scene, c1_scene

PRECISIONS = ["f64", "f32"]
HERE = os.path.dirname(os.path.abspath(__file__))


def make(sc, **kw):
    o = OraclePusher(sc["spec"], **kw)
    apply_scene(o, sc)
    return o


@pytest.mark.parametrize("precision", PRECISIONS)
def test_half_step_c_vs_numpy(precision):
    sc = small_scene(precision=precision, n=3000, speed=0.2, with_E=True, blob=(0.6, 0.9))
    o = make(sc)
    pos, vel, rnd = o.position.copy(), o.velocity.copy(), o.rand.copy()
    respawns = 0
    for k in range(12):
        o.half_step()
        pos, vel, rnd = nref.half_step(pos, vel, rnd, o.entropy, o.R1, o.R2, o.R3, o.A, o.sink_mask,
                                       o.inv_cdf, o.nr, o.nz, o.step_factor)
        assert_same(o.rand, rnd, f"rand {k}")
        assert_same(o.velocity, vel, f"velocity {k}")
        assert_same(o.position, pos, f"position {k}")
        respawns += int((pos[:, 3] == 0).sum())
    assert respawns > 50


@pytest.mark.parametrize("precision", PRECISIONS)
def test_precalc_c_vs_numpy(precision):
    sc = small_scene(precision=precision, with_E=True)
    for corrected in (False, True):
        sc["spec"]["corrected_preA"] = corrected
        o = make(sc)
        R1, R2, R3, A = nref.precalc(o.E, o.B, o.h, orc.tofixed20(o.factor_r / o.factor_z),
                                     orc.tofixed20(o.factor_z / o.factor_r), orc.tofixed20(o.factor_r),
                                     orc.tofixed20(o.factor_z), corrected)
        for a, b, nm in ((o.R1, R1, "R1"), (o.R2, R2, "R2"), (o.R3, R3, "R3"), (o.A, A, "A")):
            assert_same(a, b, nm)


def test_boris_matrix_is_the_textbook_rotation():
    """programPre1-3 (empic.js:519-606) == Boris rotation for t = h*B: norm-preserving, det 1."""
    sc = small_scene(with_E=True)
    sc["spec"]["corrected_preA"] = True
    o = make(sc)
    D = np.diag([o.factor_r, o.factor_r, o.factor_z])
    Dinv = np.linalg.inv(D)
    rng = np.random.default_rng(0)
    for c in rng.integers(0, o.ncell, 200):
        Rn = np.stack([o.R1[c, :3], o.R2[c, :3], o.R3[c, :3]])
        R = Dinv @ Rn @ D  # back to physical components
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-12)
        np.testing.assert_allclose(R, nref.boris_rotation_textbook(o.B[c, :3], o.h), atol=1e-12)
        # corrected half-kick: A = (R + I) h E / c in physical components
        A = Dinv @ o.A[c, :3]
        np.testing.assert_allclose(A, (R + np.eye(3)) @ (o.h * o.E[c, :3]) / 2.998e8, rtol=1e-9, atol=1e-18)


def test_gyro_orbit_in_uniform_bz():
    """One particle in uniform Bz: each half-step rotates v by 2*atan(h*Bz) and |v| is constant;
    the orbit radius is |v| dt c / (2 sin(theta/2)) in normalised units."""
    Bz = 0.5
    spec = dict(radius=1.0, height=1.0, nr=32, nz=32, dt=2e-9, nparticles=1, particle_mass=1.67e-27,
                particle_charge=1.602e-19)
    o = OraclePusher(spec)
    o.set({"position": [[0.5, 0.0, 0.5]], "velocity": [[0.0, 1e-3, 0.0]], "sink_mask": np.ones((32, 32)),
           "source_pdf": np.ones((32, 32)), "rand": [[0.1, 0.2, 0.3, 0.4]],
           "entropy": np.full((1024 * 1024, 4), 0.5)})
    o.addBZ(Bz)
    o.precalc()
    theta = 2 * np.arctan(o.h * Bz)
    v0 = o.velocity[0, :3].copy()
    pts = [o.position[0, :2].copy()]
    for k in range(1, 200):
        o.half_step()
        v = o.velocity[0, :3]
        np.testing.assert_allclose(np.hypot(v[0], v[1]), np.hypot(v0[0], v0[1]), rtol=1e-13)
        ang = np.arctan2(v[1], v[0]) - np.arctan2(v0[1], v0[0])
        # the reference's sign convention rotates clockwise for q Bz > 0
        np.testing.assert_allclose(np.mod(-ang, 2 * np.pi), np.mod(k * theta, 2 * np.pi), atol=1e-9)
        pts.append(o.position[0, :2].copy())
    pts = np.array(pts)
    # points lie on a circle of the Boris radius
    A = np.c_[2 * pts, np.ones(len(pts))]
    sol = np.linalg.lstsq(A, (pts ** 2).sum(1), rcond=None)[0]
    radius = np.sqrt(sol[2] + sol[0] ** 2 + sol[1] ** 2)
    expect = 1e-3 * o.step_factor / (2 * np.sin(theta / 2))
    np.testing.assert_allclose(radius, expect, rtol=1e-9)


def test_rng_stays_in_unit_interval():
    sc = small_scene(n=2000)
    o = make(sc)
    for _ in range(300):
        o.half_step()
    assert o.rand.min() >= 0.0 and o.rand.max() <= 1.0
    # the additive xy channel is equidistributed enough to cover the inverse-cdf table
    assert np.unique((o.rand[:, 0] * 512).astype(int)).size > 400


@pytest.mark.parametrize("precision", PRECISIONS)
def test_deposit_identity_sprites_vs_convolution(precision):
    """The literal point-sprite raster (GLES2 coverage + gl_PointCoord) equals NGP binning followed
    by the fixed 11x11 convolution.  Same terms, different
    association, so: tolerance = rounding of a sum of |terms|."""
    sc = small_scene(precision=precision, n=5000, speed=0.02, blob=(0.5, 0.9))
    o = make(sc)
    o.step()
    o.density(literal_sprites=True)
    lit = o.moments01.astype(np.float64).copy()
    o.moments01_avg[:] = 0
    o.density()
    conv = o.moments01.astype(np.float64)
    # per cell and channel: rounding of a sum of those very terms (conftest.deposit_reorder_bound), not a global scale
    from conftest import deposit_reorder_bound
    bound = deposit_reorder_bound(o.position, o.velocity, o.nr, o.nz, o.shape, o.dt)
    assert (np.abs(lit - conv) <= bound).all(), float((np.abs(lit - conv) - bound).max())
    # every sprite away from the edges deposits 0.001 in the alpha channel (shape sums to 1)
    p = o.position
    r = np.sqrt(p[:, 0] ** 2 + p[:, 1] ** 2)
    inside = (r * o.nr >= 6) & (r * o.nr < o.nr - 6) & (p[:, 2] * o.nz >= 6) & (p[:, 2] * o.nz < o.nz - 6)
    if inside.all():
        np.testing.assert_allclose(conv[:, 3].sum(), 0.001 * o.n, rtol=1e4 * eps)
    assert int(o.cell_count.sum()) == int(((r * o.nr < o.nr) & (p[:, 2] * o.nz >= 0) & (p[:, 2] * o.nz < o.nz)).sum())


@pytest.mark.parametrize("precision", PRECISIONS)
def test_density_c_vs_numpy(precision):
    sc = small_scene(precision=precision, n=4000, speed=0.05, blob=(0.7, 0.95), nr=40, nz=56)
    o = make(sc)
    avg = o.moments01_avg.copy()
    for frame in range(3):
        o.step()
        o.density()
        S, count = nref.cell_sums(o.position, o.velocity, o.nr, o.nz)
        assert_same(o.cell_count, count, "count")
        assert_same(o.cell_sums, S, "cell sums")
        mom = nref.convolve(S, o.shape, o.nr, o.nz)
        assert_same(o.moments01, mom, "moments01")
        norm, avg = nref.normalize_ema(mom, avg, o.nr, o.nz)
        assert_same(o.moments01_norm, norm, "norm")
        assert_same(o.moments01_avg, avg, "avg")


def test_shape_table():
    s = orc.shape_table(False).reshape(11, 11)
    np.testing.assert_allclose(s.sum(), 1.0, rtol=1e-14)
    assert (s == s.T).all() and (s == s[::-1]).all() and (s == s[:, ::-1]).all()
    assert (s == 0).sum() == 40 and s[5, 5] == s.max()
    assert 0 < s[0, 5] < 1e-30  # cos(pi/2)^2 in floating point: tiny but not zero (d == 5 exactly)
    s32 = orc.shape_table(True)
    assert (s32 == s32.astype(np.float32)).all()


def test_inverse_cdf_table_of_the_demo():
    """the demo pdf gives 1023 NaN texels (column i=511 and row j=0)."""
    sc = c1_scene(1)
    t = orc.inv_cdf(sc["source_pdf"]).reshape(512, 512, 2)  # [j][i]
    nan_y = np.isnan(t[:, :, 1])
    assert nan_y.sum() == 1023 and nan_y[0, :].all() and nan_y[:, 511].all()
    assert not np.isnan(t[:, :, 0]).any()
    x = t[0, :, 0]
    assert x[0] == 0.0 and np.all(np.diff(x) >= 0) and abs(x[511] - 50 / 400) < 1e-15
    y = t[1:, 100, 1]
    assert np.all(np.diff(y) >= 0) and y.min() >= 350 / 800 and y.max() <= 450 / 800


def test_inverse_cdf_against_python_loops():
    """Independent pure-Python restatement of empic.js:1268-1339 on a small pdf."""
    rng = np.random.default_rng(3)
    pdf = rng.random((6, 9))
    pdf[2, :] = 0  # an empty row -> 0/0
    n0, n1 = pdf.shape
    cdf_y = np.zeros((n0, n1))
    cdf_x = np.zeros(n0)
    sum_x = 0.0
    for i in range(n0):  # sequential sums, as the JS loops do
        sum_y = 0.0
        for j in range(n1):
            sum_y += pdf[i, j]
            cdf_y[i, j] = sum_y
        with np.errstate(all="ignore"):
            cdf_y[i] = cdf_y[i] / sum_y
        sum_x += sum_y
        cdf_x[i] = sum_x
    cdf_x = cdf_x / sum_x
    got = orc.inv_cdf(pdf).reshape(512, 512, 2)
    for ti in (0, 1, 77, 300, 511):
        f1 = ti / 511
        i = 0
        while cdf_x[i] < f1:
            i += 1
        x = (f1 / cdf_x[0]) / n0 if i == 0 else (i + (f1 - cdf_x[i - 1]) / (cdf_x[i] - cdf_x[i - 1])) / n0
        for tj in (0, 3, 255, 511):
            f2 = tj / 511
            ii = min(n0 - 1, int(np.floor(x * n0)))
            j = 0
            while j < n1 and cdf_y[ii][j] < f2:
                j += 1
            with np.errstate(all="ignore"):
                if j == 0:
                    y = (f2 / cdf_y[ii][0]) / n1
                else:
                    y = (j + (f2 - cdf_y[ii][j - 1]) / (cdf_y[ii][j] - cdf_y[ii][j - 1])) / n1
            assert_same(got[tj, ti], np.array([x, y]), f"texel {ti},{tj}")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_loop_table_c_vs_numpy_and_on_axis_field(precision):
    nr, nz = 24, 20
    T = np.float64 if precision == "f64" else np.float32
    spec = dict(radius=1.0, height=1.0, nr=nr, nz=nz, dt=2e-9, nparticles=1, particle_mass=1.67e-27,
                particle_charge=1.602e-19, precision=precision)
    o = OraclePusher(spec)
    half, tenth = o._tables()
    cos = o.costab
    u = ((np.arange(nr * nz) % nr).astype(T) + T(0.5)) / T(nr)
    v = ((np.arange(nr * nz) // nr).astype(T) + T(0.5)) / T(nz)
    for R, tab in ((0.5, half), (0.1, tenth)):
        R = T(R)
        const = R * T(0.001) * T(1.25663706e-6) / (T(4.0) * T(3.14159265359))
        Bx = np.zeros(nr * nz, T)
        Bz = np.zeros(nr * nz, T)
        for k in range(1000):
            c = cos[k]
            r = np.sqrt(R * R + u * u + v * v - T(2.0) * u * R * c)
            f = np.where(r > 0, const * T(1.0) / (r * r * r), T(0))
            Bx = Bx + v * f * c
            Bz = Bz + f * (R - u * c)
        assert_same(tab[:, 0], Bx, "Bx")
        assert_same(tab[:, 2], Bz, "Bz")
    if precision == "f64":
        # near the axis the table approaches the on-axis field of a unit-current loop,
        # mu0 R^2 / (2 (R^2+z^2)^1.5), divided by 2*pi: the quadrature weight of empic.js:313 is
        # 0.001 (= 1/NQUAD) where the half-circle integral needs pi/NQUAD, and the factor 2 for the
        # other half circle is absent.  A quirk of the reference, reproduced as written.
        zz = v[0::nr]
        Bz_axis = half[0::nr, 2]
        expect = 1.25663706e-6 * 0.25 / (2 * (0.25 + zz ** 2) ** 1.5) / (2 * np.pi)
        np.testing.assert_allclose(Bz_axis, expect, rtol=2e-2)


def test_add_current_loop_and_uniform_fields():
    sc = small_scene(nr=24, nz=40, n=16)
    sc["current_z"], sc["bz"], sc["btheta"] = 1.0e5, 0.25, -0.125
    o = make(sc)
    B = o.B
    # cusp symmetry of two opposing loops at z = 0 and z = height: Bz antisymmetric about mid-plane
    Bz = B[:, 2].reshape(40, 24) - 0.25
    np.testing.assert_allclose(Bz, -Bz[::-1], rtol=1e-9, atol=1e-12)
    u = (np.arange(24) + 0.5) / 24
    np.testing.assert_allclose(B[:24, 1], 1.0e5 * 1.25663706e-6 / (2 * 3.14159265359 * u) - 0.125, rtol=1e-14)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_golden_vectors(precision):
    """Committed outputs of the oracle (tests/golden/make_golden.py); not reference outputs."""
    g = np.load(os.path.join(HERE, "golden", f"oracle_small_{precision}.npz"))
    from golden.make_golden import run
    out = run(precision)
    for k in g.files:
        assert_same(out[k], g[k], f"golden {k}")


def test_canvas_matches_numpy_restatement():
    sc = small_scene(n=3000, speed=0.02, blob=(0.5, 0.8))
    o = make(sc)
    for _ in range(3):
        o.step()
        o.density()
    img = o.canvas
    B, a = o.B, o.moments01_avg[:, 3]
    mag = np.sqrt(B[:, 0] ** 2 + B[:, 1] ** 2 + B[:, 2] ** 2)
    dx, dz = B[:, 0] / mag, B[:, 2] / mag
    c1 = np.stack([mag * np.abs(np.minimum(0, dz)), mag * dx, mag * np.abs(np.maximum(0, dz)), np.ones_like(mag)], 1)
    q = lambda v: np.floor(np.clip(np.nan_to_num(v), 0, 1) * 255 + 0.5)
    src = np.stack([0.5 * a, 0.5 * a, 0.5 * a, np.full_like(a, 0.5)], 1)
    out = q(np.clip(src, 0, 1) * 0.5 + q(c1) / 255)
    expect = out.reshape(o.nz, o.nr, 4)[::-1].astype(np.uint8)
    assert_same(img, expect, "canvas")
    assert img[..., 3].min() == 255


def test_window_oracle_equals_the_whole_grid_oracle():
    """tests/window_oracle.py (the checker of the C5-size GPU test) against the whole-grid oracle: counts,
    per-cell sums and running average of a window at either edge and in the middle, with NaN and
    out-of-range positions among the particles."""
    from window_oracle import oracle_window
    sc = small_scene(n=30000, nr=40, nz=96, speed=0.3, blob=(0.95, 0.99))
    o = OraclePusher(sc["spec"], nthreads=2)
    apply_scene(o, sc)
    for frame in range(3):
        prev = o.moments01_avg.copy()
        o.step()
        o.position[5::97, 2] = np.nan      # what a NaN respawn texel leaves behind
        o.position[7::101, 2] = 1.25       # beyond the top edge: clipped sprite
        o.density()
        gp, gv = o.getPosition(), o.getVelocity()
        for (w0, w1) in ((0, 16), (40, 56), (80, 96), (0, 96)):
            a, b = w0 * o.nr, w1 * o.nr
            cnt, S, avg, nsel = oracle_window(gp, gv, prev[a:b], o.nr, o.nz, w0, w1, threads=2)
            assert nsel > 100
            assert_same(cnt, o.cell_count[a:b], f"frame {frame} rows {w0}..{w1} counts")
            assert_same(S, o.cell_sums[a:b], f"frame {frame} rows {w0}..{w1} sums")
            assert_same(avg, o.moments01_avg[a:b], f"frame {frame} rows {w0}..{w1} running average")
