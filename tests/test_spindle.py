"""Spindle-cusp boundary solve: out.addSpindleCuspPlasmaField (empic.js:1369), "next" row N3 of SURVEY.md section
8f, second half.  spindle.js does not run in the reference, so the specification is written from its intent
(include/fusionsim.h) and parity is UNPINNED: what is tested is

  * the CPU oracle (oracle/fsim_oracle_spindle_impl.h) against a second, separately written NumPy statement of
    the assembly, against a direct dense solve, and against the PHYSICS the solve is for -- B.n = 0 on the
    plasma surface and hence (a curl-free, divergence-free field with no normal component on a closed surface
    vanishes inside it) no field inside the plasma, unchanged field far outside;
  * on the GPU, libfusionsim.so against that oracle, bit for bit (matrix, right-hand side, solution of the
    device weighted-Jacobi routine, node currents, B) in fp64 and fp32.
"""
import numpy as np
import pytest

from conftest import assert_same

MU0 = 1.25663706e-6
COS = np.cos(3.14159265359 * (np.arange(1000) + 0.5) / 1000.0)


def loop_field(Rl, Zl, x, z):
    """NumPy statement of the loop field of the specification (vectorised over points)."""
    x = np.asarray(x, float)[..., None]
    dz = (np.asarray(z, float) - Zl)[..., None]
    rho = np.sqrt(Rl * Rl + x * x + dz * dz - 2 * x * Rl * COS)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.where(rho > 0, Rl * 0.00628318530718 * MU0 / (4 * 3.14159265359) / rho ** 3, 0.0)
    return (dz * f * COS).sum(-1), (f * (Rl - x * COS)).sum(-1)


def numpy_system(radius, height, coil_r, coil_I, L=256, a=0.4):
    s = height / (2 * radius)
    R = radius * np.sqrt(1 + a * a)
    alpha = np.arctan(a)
    theta, arc = alpha + np.pi, 0.5 * np.pi - 2 * alpha
    phn = theta + np.arange(L + 1) * arc / L
    xn, zn = R * np.cos(-phn) + radius, s * R * np.sin(-phn)
    xn[0] = 0.0
    phm = theta + (np.arange(L) + 0.5) * arc / L
    xp, zp = R * np.cos(-phm) + radius, s * R * np.sin(-phm)
    nx, nz = -s * np.cos(-phm), -np.sin(-phm)
    nn = np.hypot(nx, nz)
    nx, nz = nx / nn, nz / nn
    nf = []
    for l in range(L + 1):
        g, m = loop_field(xn[l], zn[l], xp, zp), loop_field(xn[l], height - zn[l], xp, zp)
        nf.append((g[0] - m[0], g[1] - m[1]))
    A = np.zeros((L, L))
    for e in range(L):
        A[:, e] = nx * (nf[e][0] - nf[e + 1][0]) + nz * (nf[e][1] - nf[e + 1][1])
    g, m = loop_field(coil_r, 0.0, xp, zp), loop_field(coil_r, height, xp, zp)
    b = coil_I * (nx * (g[0] - m[0]) + nz * (g[1] - m[1]))
    return A, b, (xn, zn, xp, zp, nx, nz)


def spec_for(nr=48, nz=96, radius=1.0, height=2.0, precision="f64"):
    return dict(radius=radius, height=height, nr=nr, nz=nz, dt=2e-9, nparticles=4, particle_mass=1.67e-27,
                particle_charge=1.602e-19, precision=precision)


@pytest.fixture(scope="module")
def solved():
    from oracle.oracle import OraclePusher
    o = OraclePusher(spec_for(), nthreads=4)
    sp = o.addSpindleCuspPlasmaField(0.8, 0.5, 1.0)
    return o, sp


def test_oracle_assembly_against_numpy_and_direct_solve(solved):
    o, sp = solved
    A, b, _ = numpy_system(1.0, 2.0, 0.8, sp["coil_current"])
    L = 256
    scale = np.abs(A).max()
    assert np.abs(sp["A"][:L - 1, :L - 1] - A[:L - 1, :L - 1]).max() <= 1e-12 * scale  # different association only
    assert np.abs(sp["rhs"][:L - 1] + b[:L - 1]).max() <= 1e-12 * np.abs(b).max()
    # gauge: row and column L-1 are the identity's
    assert sp["A"][L - 1, L - 1] == 1 and not sp["A"][L - 1, :L - 1].any() and not sp["A"][:L - 1, L - 1].any()
    assert sp["rhs"][L - 1] == 0 and sp["x"][L - 1] == 0
    # the constant vector is (numerically) a null vector of the UNgauged matrix: element strengths are a stream function
    assert np.abs(A @ np.ones(L)).max() <= 1e-9 * scale
    # weighted Jacobi (the reference's solver) converged to the direct solution
    x = np.linalg.solve(sp["A"], sp["rhs"])
    assert sp["diff"] <= 1e-9 and 0 < sp["iterations"] < 4000
    assert np.abs(sp["x"] - x).max() <= 2e-6 * np.abs(x).max()
    # node l carries x_l - x_{l-1}
    cur = sp["currents"]
    assert cur[0] == sp["x"][0] and cur[256] == -sp["x"][255]
    assert_same(cur[1:256], sp["x"][1:] - sp["x"][:-1], "node currents")


def test_field_is_excluded_from_the_plasma(solved):
    """B.n = 0 on the surface, no field inside, vacuum field far outside."""
    o, sp = solved
    from oracle.oracle import OraclePusher
    vac = OraclePusher(spec_for(), nthreads=4)
    vac.addSpindleCuspPlasmaField(0.8, 0.5, 0.0)  # beta 0: surface currents scaled to nothing = the coils alone
    B = o.B.reshape(96, 48, 4)
    V = vac.B.reshape(96, 48, 4)
    mag, vmag = np.hypot(B[..., 0], B[..., 2]), np.hypot(V[..., 0], V[..., 2])
    # the field of one coil at its own centre is B_c = 0.5 T up to the other coil and the half-cell offset
    assert abs(V[0, 0, 2] - 0.5) < 0.05
    # well inside the plasma (axis side of the arc from (0, 0.4) to (0.6, 1.0), and its mirror image)
    inside = [(2, 40), (5, 45), (10, 47), (14, 46), (2, 55), (10, 48), (5, 50)]
    for (i, j) in inside:
        assert mag[j, i] < 0.01 * vmag[j, i], (i, j, mag[j, i] / vmag[j, i])
    # far outside: near the coils and the wall the plasma currents change little
    for (i, j) in [(14, 10), (40, 5), (40, 90), (14, 85)]:
        assert abs(mag[j, i] / vmag[j, i] - 1) < 0.05, (i, j, mag[j, i] / vmag[j, i])
    # the cusp is antisymmetric about the mid-plane: B_r(z) = B_r(H - z), B_z(z) = -B_z(H - z)
    np.testing.assert_allclose(B[::-1, :, 0], B[:, :, 0], rtol=0, atol=1e-9 * np.abs(B).max())
    np.testing.assert_allclose(B[::-1, :, 2], -B[:, :, 2], rtol=0, atol=1e-9 * np.abs(B).max())
    # B.n on the surface, evaluated with the NumPy loop field from all the loops the solve superposed
    _, _, (xn, zn, xp, zp, nx, nz) = numpy_system(1.0, 2.0, 0.8, sp["coil_current"])
    br, bz = np.zeros(256), np.zeros(256)
    for (Rl, Zl, I) in sp["loops"]:
        g = loop_field(Rl, Zl, xp, zp)
        br += I * g[0]
        bz += I * g[1]
    g, m = loop_field(0.8, 0.0, xp, zp), loop_field(0.8, 2.0, xp, zp)
    coil_n = np.abs(sp["coil_current"] * (nx * (g[0] - m[0]) + nz * (g[1] - m[1])))
    assert np.abs(nx * br + nz * bz)[:255].max() < 1e-5 * coil_n.max()


def test_partial_exclusion_and_other_aspect_ratio():
    from oracle.oracle import OraclePusher
    full = OraclePusher(spec_for(32, 48, radius=0.7, height=1.1), nthreads=4)
    half = OraclePusher(spec_for(32, 48, radius=0.7, height=1.1), nthreads=4)
    none = OraclePusher(spec_for(32, 48, radius=0.7, height=1.1), nthreads=4)
    sf = full.addSpindleCuspPlasmaField(0.6, 0.3, 1.0)
    sh = half.addSpindleCuspPlasmaField(0.6, 0.3, 0.75)   # 1 - sqrt(1 - 0.75) = 0.5
    none.addSpindleCuspPlasmaField(0.6, 0.3, 0.0)
    assert sf["diff"] <= 1e-9
    np.testing.assert_allclose(sh["currents"], 0.5 * sf["currents"], rtol=1e-14, atol=0)
    # linear in the surface currents: the half-excluded field lies half way between vacuum and full exclusion
    np.testing.assert_allclose(half.B, 0.5 * (full.B + none.B), rtol=0, atol=1e-12 * np.abs(none.B).max())
    B, V = full.B.reshape(48, 32, 4), none.B.reshape(48, 32, 4)
    i, j = 3, 23  # inside the plasma of the squashed (s = height / 2 radius = 0.786) surface
    assert np.hypot(B[j, i, 0], B[j, i, 2]) < 0.02 * np.hypot(V[j, i, 0], V[j, i, 2])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_gpu_spindle_solve_equals_the_oracle(precision):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from oracle.oracle import OraclePusher
    spec = spec_for(precision=precision)
    g = makeCylindricalParticlePusher(spec)
    o = OraclePusher(spec, nthreads=4)
    # on top of a field that is already there (blended ONE, ONE like every field builder)
    g.addBZ(0.01); o.addBZ(0.01)
    got = g.addSpindleCuspPlasmaField(0.8, 0.5, 0.9)
    want = o.addSpindleCuspPlasmaField(0.8, 0.5, 0.9)
    assert_same(got["A"], want["A"], "boundary matrix")
    assert_same(got["rhs"], want["rhs"], "right-hand side")
    assert got["iterations"] == want["iterations"]
    assert_same(got["x"], want["x"], "element strengths (device weighted-Jacobi routine)")
    assert_same(got["currents"], want["currents"], "node currents")
    assert_same(g.getField("B"), o.getField("B"), "B")
    g.precalc(); o.precalc()
    assert_same(g.getField("R1"), o.getField("R1"), "Boris rows in the spindle-cusp field")


@pytest.mark.gpu
def test_gpu_spindle_arguments_are_validated():
    from fusion_sim_b200 import Error, makeCylindricalParticlePusher
    g = makeCylindricalParticlePusher(spec_for(16, 16))
    with pytest.raises(Error, match="beta_c"):
        g.addSpindleCuspPlasmaField(0.8, 0.5, 1.5)
    with pytest.raises(Error, match="coil radius"):
        g.addSpindleCuspPlasmaField(-1.0, 0.5, 1.0)
