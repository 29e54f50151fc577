"""EXTENSION (no reference counterpart): the self-consistent
electrostatic field solve.  CPU: the C oracle (oracle/fsim_oracle_fields_impl.h, the written
specification) against an independent NumPy restatement bit for bit, and against a direct sparse
solve of the same discrete operator.  GPU: libfusionsim.so against the oracle, bit-exact."""
import numpy as np
import pytest

from conftest import assert_same, small_scene


def _oracle(sc):
    from fusion_sim_b200.scenes import apply_scene
    from oracle.oracle import OraclePusher
    o = OraclePusher(sc["spec"], nthreads=4)
    apply_scene(o, sc)
    return o


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_oracle_matches_numpy_restatement(precision):
    from oracle import numpy_ref as nref
    sc = small_scene(n=20000, speed=0.02, blob=(0.5, 0.8), precision=precision, nr=37, nz=53)
    o = _oracle(sc)
    o.step(); o.density()
    sp = o.spec
    dr, dz = sp["radius"] / o.nr, sp["height"] / o.nz
    coef = nref.relax_coeffs(o.nr, dr, dz)
    phi = np.zeros(o.ncell, o.dt)
    for sweeps, omega in ((1, 1.0), (7, 0.9), (4, 0.5)):
        o.solveFields({"macro_weight": 1e9, "sweeps": sweeps, "omega": omega})
        phi = nref.relax(phi, o.rho_src, coef, omega, sweeps, o.nr, o.nz)  # warm start, like the oracle
        assert_same(o.phi, phi, f"phi after {sweeps} sweeps")
        assert_same(o.E, nref.efield(phi, o.nr, o.nz, 1 / (2 * dr), 1 / (2 * dz)), "E")
    assert np.abs(o.phi).max() > 0 and np.abs(o.E[:, 0]).max() > 0
    assert np.all(o.E[:, 1] == 0)


def test_relaxation_converges_to_the_sparse_solution():
    """Many sweeps reach the solution of the discrete system A phi = src (scipy sparse LU)."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spl
    from oracle import numpy_ref as nref
    nr, nz, dr, dz = 12, 10, 0.05, 0.07
    coef = nref.relax_coeffs(nr, dr, dz)
    rng = np.random.Generator(np.random.PCG64(1))
    src = rng.random(nr * nz)
    # rows of the system in the Jacobi-normalised form: phi - cE phi_E - cW phi_W - cZ (phi_N + phi_S) = cB src
    A = sps.lil_matrix((nr * nz, nr * nz))
    for j in range(nz):
        for i in range(nr):
            c = i + j * nr
            A[c, c] = 1.0
            if i + 1 < nr: A[c, c + 1] = -coef[i, 0]
            if i > 0: A[c, c - 1] = -coef[i, 1]
            if j + 1 < nz: A[c, c + nr] = -coef[i, 2]
            if j > 0: A[c, c - nr] = -coef[i, 2]
    exact = spl.spsolve(A.tocsc(), np.tile(coef[:, 3], nz) * src)
    phi = nref.relax(np.zeros(nr * nz), src, coef, 1.0, 3000, nr, nz)
    assert np.abs(phi - exact).max() <= 1e-10 * np.abs(exact).max()
    # the operator is the finite-volume cylindrical Laplacian: a uniform source gives phi > 0 inside
    # grounded walls, largest on the axis at mid-height
    phi_u = nref.relax(np.zeros(nr * nz), np.ones(nr * nz), coef, 1.0, 3000, nr, nz).reshape(nz, nr)
    assert phi_u.min() > 0 and np.unravel_index(phi_u.argmax(), phi_u.shape)[1] == 0


def test_repulsion_sign():
    """A positive charge cloud on the axis pushes positive particles outwards: E_r > 0 outside it."""
    sc = small_scene(n=20000, speed=0.0, blob=(0.1, 0.1))
    o = _oracle(sc)
    o.density()
    o.solveFields({"macro_weight": 1e10, "sweeps": 400, "source": "instant"})
    E = o.E.reshape(o.nz, o.nr, 4)
    assert E[o.nz // 2, o.nr // 2, 0] > 0
    assert E[o.nz // 4, 2, 2] < 0 < E[3 * o.nz // 4, 2, 2]  # E_z points away from the cloud at z = height/2


# ---- GPU parity ---------------------------------------------------------------------------------
def _pair(sc):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene
    g = makeCylindricalParticlePusher(sc["spec"])
    apply_scene(g, sc)
    return g, _oracle(sc)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("shape", [(48, 96), (37, 53), (130, 67)])
def test_gpu_field_solve_bit_exact(precision, shape):
    """Grids that are not multiples of the 64x32 tile, every sweeps-per-launch split (4+2+1), warm
    starts, both densities; phi, rho/eps0, E and the Boris records must match the oracle bit for bit."""
    sc = small_scene(n=20000, speed=0.02, blob=(0.5, 0.8), precision=precision, nr=shape[0], nz=shape[1])
    g, o = _pair(sc)
    g.step(); o.step()
    g.density(); o.density()
    for sweeps, omega, source in ((1, 1.0, "avg"), (2, 0.9, "instant"), (7, 0.8, "avg"), (4, 1.0, "instant"),
                                  (11, 1.2, "avg"), (0, 1.0, "avg")):
        v = {"macro_weight": 3e9, "sweeps": sweeps, "omega": omega, "source": source}
        g.solveFields(v); o.solveFields(v)
        what = f"{sweeps} sweeps, omega {omega}, {source}"
        assert_same(g.getField("rho_src"), o.getField("rho_src"), what + " rho_src")
        assert_same(g.getField("phi"), o.getField("phi"), what + " phi")
        for nm in ("E", "R1", "R2", "R3", "A"):
            assert_same(g.getField(nm), o.getField(nm), what + " " + nm)
    assert np.abs(g.getField("phi")).max() > 0


@pytest.mark.gpu
def test_gpu_self_consistent_loop_bit_exact():
    """step -> density -> solveFields for several frames: the particles feel the field they made."""
    sc = small_scene(n=20000, speed=0.01, blob=(0.3, 0.5))
    g, o = _pair(sc)
    v = {"macro_weight": 5e11, "sweeps": 9, "omega": 1.0}
    for frame in range(5):
        g.step(); o.step()
        g.density(); o.density()
        g.solveFields(v); o.solveFields(v)
    assert_same(g.getPosition(), o.getPosition(), "position")
    assert_same(g.getVelocity(), o.getVelocity(), "velocity")
    assert_same(g.getField("phi"), o.getField("phi"), "phi")
    assert np.abs(o.A).max() > 0  # the half-kick constant is live


@pytest.mark.gpu
def test_gpu_field_solve_errors():
    from fusion_sim_b200 import Error
    sc = small_scene(n=256)
    sc["spec"]["keep_moments"] = False
    g, _ = _pair(sc)
    g.density()
    with pytest.raises(Error):
        g.solveFields({"macro_weight": 1.0, "sweeps": 1, "source": "instant"})  # needs keep_moments
    with pytest.raises(Error):
        g.solveFields({"macro_weight": 1.0, "sweeps": 1, "omega": 2.5})
    with pytest.raises(Error):
        g.solveFields({"sweeps": 1})
    g.solveFields({"macro_weight": 1.0, "sweeps": 3})
