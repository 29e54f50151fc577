"""Dense weighted-Jacobi solver (matrix_webgl.makeSORIterative, "next" row N3).
CPU: the C oracle against a NumPy restatement of the packed reduction tree, convergence to
numpy.linalg.solve in the intended mode, the author's diagonal smoke test (fusionsim.js:35-67) in
literal mode.  GPU: the CUDA routine against the oracle, bit for bit, through the C ABI."""
import numpy as np
import pytest

from conftest import assert_same

PRECISIONS = ["f64", "f32"]


def system(n_power, seed=0, dominance=2.0):
    L = 4 * (2 ** n_power) ** 2
    rng = np.random.default_rng(seed)
    A = rng.random((L, L)) - 0.5
    A[np.arange(L), np.arange(L)] = dominance * np.abs(A).sum(1) * (1 + rng.random(L))
    return A, rng.random(L), rng.random(L)


def numpy_row_sum(prod, vh):
    """The reference's order: 2x2 texel blocks (+x,+y),(-x,+y),(+x,-y),(-x,-y), then ((r+g)+b)+a."""
    t = prod.reshape(vh, vh, 4)  # [y][x][channel]
    while t.shape[0] > 1:
        t = ((t[1::2, 1::2] + t[1::2, 0::2]) + t[0::2, 1::2]) + t[0::2, 0::2]
    v = t[0, 0]
    return ((v[0] + v[1]) + v[2]) + v[3]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("n_power", [1, 2, 3])
def test_oracle_matches_numpy_tree(precision, n_power):
    from oracle.jacobi import OracleSOR
    from oracle.oracle import tofixed20
    A, b, x0 = system(n_power)
    for omega in (1.0, 0.75):
        o = OracleSOR({"n_power": n_power, "relaxation": omega, "precision": precision})
        o.set_matrix(A).set_b(b).init_vector(x0)
        res = o.solve({"tolerance": 0.0, "max_iterations": 2})
        T = o.dt
        Ad, bd = A.astype(T), b.astype(T)
        d = np.diag(Ad).copy()
        R = -Ad / d[:, None]
        np.fill_diagonal(R, 0)
        C = bd / d
        if omega != 1.0:
            R, C = T(tofixed20(omega)) * R, T(tofixed20(omega)) * C
        x = x0.astype(T)
        for _ in range(2):
            s = np.array([numpy_row_sum(R[r] * x, o.vec_height) for r in range(o.vec_length)], T)
            xn = s + C
            if omega != 1.0:
                xn = xn + T(tofixed20(1.0 - omega)) * x
            x = xn
        assert_same(res["result"], x.astype(np.float64), f"omega {omega}")
        assert res["iterations"] == 2


def test_intent_mode_solves_the_system():
    from oracle.jacobi import OracleSOR
    A, b, x0 = system(2, seed=3)
    o = OracleSOR({"n_power": 2})
    res = o.set_matrix(A).set_b(b).init_vector(x0).solve({"tolerance": 1e-13, "max_iterations": 200})
    assert res["iterations"] < 200 and res["diff"] <= 1e-13
    np.testing.assert_allclose(res["result"], np.linalg.solve(A, b), rtol=1e-10)
    assert abs(res["correlation"]) <= 1.0 + 1e-9 or np.isnan(res["correlation"])


def test_literal_mode_reproduces_the_reference_defects():
    """fusionsim.js:35-67: a diagonal matrix hides the row-gather defect (x = b/diag after one
    iteration); a dense matrix exposes it (literal != intent)."""
    from oracle.jacobi import OracleSOR
    L = 16
    rng = np.random.default_rng(1)
    A = np.diag(rng.random(L) + 0.5)
    b = rng.random(L)
    lit = OracleSOR({"n_power": 1, "literal": True}).set_matrix(A).set_b(b)
    res = lit.solve({"tolerance": 1e-3, "substep": 1, "max_iterations": 100})
    np.testing.assert_allclose(res["result"], b / np.diag(A), rtol=1e-15)
    A2, b2, x0 = system(1, seed=5)
    r_lit = OracleSOR({"n_power": 1, "literal": True}).set_matrix(A2).set_b(b2).init_vector(x0).solve(
        {"tolerance": 0.0, "max_iterations": 1})["result"]
    r_int = OracleSOR({"n_power": 1}).set_matrix(A2).set_b(b2).init_vector(x0).solve(
        {"tolerance": 0.0, "max_iterations": 1})["result"]
    assert not np.array_equal(r_lit, r_int)
    # entry e receives the sum of row {2px, 2px+1, 2px+2vh, 2px+2vh+1} + 4 vh py (programResult :408-411)
    vh = 2
    for e in range(L):
        pix, k = divmod(e, 4)
        px, py = pix % vh, pix // vh
        row = 2 * px + 4 * vh * py + (k & 1) + (2 * vh if k >> 1 else 0)
        d = np.diag(A2)
        Rrow = -A2[row] / d[row]
        Rrow[row] = 0
        want = numpy_row_sum(Rrow * x0, vh) + b2[e] / d[e]
        assert r_lit[e] == want
    # no max_iterations => the loop never runs (iteration < undefined is false)
    assert OracleSOR({"n_power": 1}).set_matrix(A2).set_b(b2).solve({"tolerance": 1e-3})["iterations"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("n_power", [1, 2, 3, 4])
def test_cuda_matches_oracle(precision, n_power):
    from fusion_sim_b200.matrix import makeSORIterative
    from oracle.jacobi import OracleSOR
    A, b, x0 = system(n_power, seed=n_power)
    for spec in ({"relaxation": 1.0}, {"relaxation": 0.8}, {"relaxation": 0.8, "literal": True}):
        spec = dict(spec, n_power=n_power, precision=precision)
        params = {"tolerance": 1e-9, "substep": 2, "max_iterations": 6}
        g = makeSORIterative(spec).set_matrix(A).set_b(b).init_vector(x0)
        o = OracleSOR(spec, nthreads=4).set_matrix(A).set_b(b).init_vector(x0)
        rg, ro = g.solve(params), o.solve(params)
        assert rg["iterations"] == ro["iterations"]
        assert_same(rg["result"], ro["result"], f"result {spec}")
        assert_same(np.array([rg["diff"], rg["correlation"]]), np.array([ro["diff"], ro["correlation"]]), "stats")
        assert_same(g.x_result_tex(), ro["result"], "x_result_tex")


@pytest.mark.gpu
def test_cuda_solves_a_large_system():
    """n_power = 5: 4096 unknowns, 134 MB matrix -- the solution satisfies A x = b."""
    from fusion_sim_b200.matrix import makeSORIterative
    A, b, _ = system(5, seed=9)
    g = makeSORIterative({"n_power": 5}).set_matrix(A).set_b(b)
    res = g.solve({"tolerance": 1e-14, "substep": 4, "max_iterations": 50})
    x = res["result"]
    assert np.abs(A @ x - b).max() < 1e-10 * np.abs(b).max() * 4096
    np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-9)


@pytest.mark.gpu
def test_errors():
    from fusion_sim_b200 import Error
    from fusion_sim_b200.matrix import makeSORIterative
    with pytest.raises(Error, match=r"\.n_power <- Non-optional property is undefined!"):
        makeSORIterative({})
    g = makeSORIterative({"n_power": 1})
    with pytest.raises(Error, match="set_matrix and set_b"):
        g.solve({"tolerance": 1e-3, "max_iterations": 3})
    with pytest.raises(Error, match=r"\.tolerance"):
        g.solve({})
