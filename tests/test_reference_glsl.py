"""The CPU oracle against the REFERENCE'S OWN SHADER SOURCE.

tests/golden/reference_glsl_{f64,f32}.npz hold the inputs and outputs of executing the GLSL text of
every program on the hot path -- read at generation time out of
/root/reference/public/javascripts/empic.js, never copied here -- with the small GLSL ES 1.00
interpreter in oracle/glsl_interp.py (tests/golden/make_reference_vectors.py).  The oracle must
reproduce them BIT FOR BIT: every operand, sign, constant and the order of operations of
programCurrentLoopShape/CurrentLoop/CurrentZ/BZ/BTheta, programPre1/2/3/A, programStepRand,
step_velocity_frag, step_position_frag, programMoments01 (vertex + fragment),
programNormalizeMoments01 and avg_frag are thereby pinned to the reference's text.  What stays
unpinned is what GLSL ES 1.00 itself leaves open (rounding of / and sqrt, evaluation order, NaN
texture coordinates): both sides use the documented IEEE / left-to-right choices."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import assert_same

HERE = os.path.dirname(os.path.abspath(__file__))
PRECISIONS = ["f64", "f32"]
SPEC = dict(radius=1.0, height=2.0, nr=24, nz=40, dt=2e-9, nparticles=16,
            particle_mass=1.67e-27, particle_charge=1.602e-19)


def load(precision):
    return np.load(os.path.join(HERE, "golden", f"reference_glsl_{precision}.npz"))


def entropy_table():
    k = np.arange(1024 * 1024 * 4, dtype=np.float64)
    return np.mod(k * 0.6180339887498949, 1.0).reshape(-1, 4)


def oracle_for(precision):
    from oracle.oracle import OraclePusher
    return OraclePusher(dict(SPEC, precision=precision))


@pytest.mark.parametrize("precision", PRECISIONS)
def test_static_fields_match_the_reference_shaders(precision):
    d = load(precision)
    o = oracle_for(precision)
    half, tenth = o._tables()
    assert_same(half, d["loop_half"], "programCurrentLoopShape, u_R = 0.5")
    assert_same(tenth, d["loop_tenth"], "programCurrentLoopShape, u_R = 0.1")
    for r, z, I in d["loops"]:
        o.addCurrentLoop(float(r), float(z), float(I))
    assert_same(o.B, d["B_loops"], "programCurrentLoop x 2, blended")
    cz, bz, bt = (float(v) for v in d["uniform_terms"])
    o.addCurrentZ(cz); o.addBZ(bz); o.addBTheta(bt)
    assert_same(o.B, d["B"], "programCurrentZ + programBZ + programBTheta")
    assert np.abs(d["B"][:, :3]).max() > 0


@pytest.mark.parametrize("precision", PRECISIONS)
def test_precalc_matches_the_reference_shaders(precision):
    d = load(precision)
    o = oracle_for(precision)
    o.B[:] = d["B"]
    o.E[:] = d["E"]
    o.precalc()
    for nm in ("R1", "R2", "R3", "A"):
        assert_same(getattr(o, nm), d[nm], "programPre" + nm[-1])
    assert np.abs(d["A"][:, :3]).max() > 0  # E != 0: the scalar-added-to-vector term of :645 is live


@pytest.mark.parametrize("precision", PRECISIONS)
def test_half_steps_match_the_reference_shaders(precision):
    """Eight half-steps of programStepRand + step_velocity_frag + step_position_frag, from a state
    with just-respawned particles, r = 0, absorption at the wall and NaN inverse-cdf texels."""
    from oracle import oracle as orc
    d = load(precision)
    o = oracle_for(precision)
    for nm in ("R1", "R2", "R3", "A"):
        getattr(o, nm)[:] = d[nm]
    o.sink_mask[:] = d["sink"]
    o.inv_cdf[:, :2] = orc.inv_cdf(d["source_pdf"])
    o.entropy[:] = entropy_table()
    o.position[:], o.velocity[:], o.rand[:] = d["position"], d["velocity"], d["rand"]
    respawned = 0
    live = np.ones(o.n, bool)  # particles still compared
    deviated = 0
    for k in range(d["step_position"].shape[0]):
        o.half_step()
        ref_pos = d["step_position"][k]
        # The one DOCUMENTED deviation (oracle header): a pushed position with NaN r or
        # z makes the sink lookup a NaN texture coordinate -- undefined in GL; read literally as "texel 0"
        # the particle would stay NaN for ever wherever sink[0][0] = 1.  Oracle and product absorb and
        # respawn it instead.  Such a particle leaves the comparison at that half-step.
        nan_kept = live & (ref_pos[:, 3] == 1) & np.isnan(ref_pos[:, :3]).any(1)
        assert (o.position[nan_kept, 3] == 0).all()  # ... and the oracle did respawn it
        deviated += int(nan_kept.sum())
        live &= ~nan_kept
        assert_same(o.rand[live], d["step_rand"][k][live], f"half-step {k}: programStepRand")
        assert_same(o.velocity[live], d["step_velocity"][k][live], f"half-step {k}: step_velocity_frag")
        assert_same(o.position[live], ref_pos[live], f"half-step {k}: step_position_frag")
        # keep the two states identical for the excluded particles too, so later steps stay comparable
        o.position[~live], o.velocity[~live], o.rand[~live] = ref_pos[~live], d["step_velocity"][k][~live], d["step_rand"][k][~live]
        respawned += int((o.position[live, 3] == 0).sum())
    assert respawned > 20 and np.isnan(d["step_position"]).any()
    assert 1 <= deviated <= 8 and live.sum() >= o.n - 8  # the r = 0 particle, and respawns into NaN texels


@pytest.mark.parametrize("precision", PRECISIONS)
def test_deposit_matches_the_reference_shaders(precision):
    """programMoments01: the vertex shader's colour and the fragment shader's product on every pixel a
    size-11 sprite covers, blended in particle order (literal form), then the canonical NGP-sums (*)
    footprint form; programNormalizeMoments01 and avg_frag."""
    from oracle import oracle as orc
    d = load(precision)
    o = oracle_for(precision)
    o.position[:], o.velocity[:] = d["step_position"][-1], d["step_velocity"][-1]
    o.density(literal_sprites=True)
    assert_same(o.moments01, d["sprite_moments01"], "programMoments01 sprites")
    # colours: with the sprite weights set to one the literal deposit of a single particle is 0 + v_color
    ones = np.ones(121, o.dt)
    for p in (0, 17, 100, 255):
        mom = np.zeros((o.ncell, 4), o.dt)
        o._f("orc_deposit_sprites")(C.c_int64(1), orc._p(np.ascontiguousarray(o.position[p:p + 1])),
                                    orc._p(np.ascontiguousarray(o.velocity[p:p + 1])), orc._p(ones), C.c_int64(o.nr),
                                    C.c_int64(o.nz), orc._p(mom))
        hit = mom[np.abs(mom[:, 3]) > 0]
        if len(hit):
            assert_same(hit[0], d["sprite_color"][p], f"v_color of particle {p}")
    # the canonical convolution form agrees with the literal one to rounding (different association)
    lit = o.moments01.copy()
    o.density()
    # -- per cell and channel within the rounding of a re-ordered sum of the same terms (conftest.deposit_reorder_bound)
    from conftest import deposit_reorder_bound
    bound = deposit_reorder_bound(o.position, o.velocity, o.nr, o.nz, o.shape, o.dt)
    ok = ~np.isnan(lit)
    diff = np.abs(o.moments01.astype(np.float64) - lit.astype(np.float64))
    assert (diff[ok] <= bound[ok]).all(), float((diff[ok] - bound[ok]).max())
    # normalise + running average
    o.moments01[:] = d["moments01"]
    o.moments01_avg[:] = d["avg0"]
    o._f("orc_normalize_ema")(C.c_int64(o.nr), C.c_int64(o.nz), orc._p(o.moments01), orc._p(o.moments01_norm),
                              orc._p(o.moments01_avg), C.c_int(1))
    assert_same(o.moments01_norm, d["moments01_norm"], "programNormalizeMoments01")
    assert_same(o.moments01_avg, d["moments01_avg"], "avg_frag")
    # the canvas the page draws: programBMag + programDensity (the fixed-function clamp / round / blend are GL's)
    o.B[:] = d["B"]
    assert_same(o.canvas, d["canvas"], "programBMag + programDensity -> canvas")
    assert len(np.unique(d["canvas"][..., :3])) > 8


def test_vectors_regenerate_from_the_reference_tree():
    """Where the reference tree is present (this container, not the GPU box) the committed vectors
    must be exactly what executing its shader source gives today."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_reference_vectors as gen
    if not os.path.exists(gen.REF_JS):
        pytest.skip("reference tree not present")
    for precision, dt in (("f64", np.float64), ("f32", np.float32)):
        d, res = load(precision), gen.run(dt)  # full=False: the demo-scene inverse cdf (minutes) is not re-run
        assert set(d.files) - set(res) <= {"invcdf_demo_digest", "invcdf_demo_nan_texels"} and set(res) <= set(d.files)
        for k in res:
            assert_same(np.asarray(res[k]), d[k], f"{precision} {k}")
        d, res = np.load(os.path.join(HERE, "golden", f"reference_glsl_odd_{precision}.npz")), gen.run(dt, sp=gen.SPEC_ODD)
        for k in d.files:
            assert_same(np.asarray(res[k]), d[k], f"odd units, {precision} {k}")


@pytest.mark.gpu
@pytest.mark.parametrize("precision", PRECISIONS)
def test_gpu_fields_match_the_reference_shaders(precision):
    """libfusionsim.so straight against the reference-shader vectors (static fields and precalc; the
    step and deposit kernels are held to the oracle, which the tests above hold to these vectors)."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    d = load(precision)
    g = makeCylindricalParticlePusher(dict(SPEC, precision=precision))
    nr, nz = SPEC["nr"], SPEC["nz"]
    E = d["E"][:, :3].astype(np.float64).reshape(nz, nr, 3).transpose(1, 0, 2)  # value.E[i][j][k]
    g.set({"E": np.ascontiguousarray(E)})
    for r, z, I in d["loops"]:
        g.addCurrentLoop(float(r), float(z), float(I))
    assert_same(g.getField("B"), d["B_loops"][:, :3].astype(np.float64), "B after the loops")
    cz, bz, bt = (float(v) for v in d["uniform_terms"])
    g.addCurrentZ(cz); g.addBZ(bz); g.addBTheta(bt)
    assert_same(g.getField("B"), d["B"][:, :3].astype(np.float64), "B")
    g.precalc()
    for nm in ("R1", "R2", "R3", "A"):
        assert_same(g.getField(nm), d[nm][:, :3].astype(np.float64), nm)


# ---- matrix_webgl.makeSORIterative (row N3) against ITS shader source ----------------------------
JACOBI_CASES = [(1, 1.0), (2, 1.0), (2, 0.8)]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("n_power,omega", JACOBI_CASES)
def test_jacobi_oracle_matches_the_reference_shaders(precision, n_power, omega):
    """programR, programC, programMVproduct + sum_frag chain + programResult (two iterations) and
    programStats of matrix_webgl.js, executed from source, against the LITERAL mode of the oracle
    (the mode that keeps the row numbering of programResult :408-411 as written)."""
    from oracle import oracle as orc
    from oracle.jacobi import OracleSOR
    d = np.load(os.path.join(HERE, "golden", "reference_glsl_jacobi.npz"))
    g = lambda k: d[f"{precision}_p{n_power}_w{int(10 * omega)}_{k}"]
    o = OracleSOR({"n_power": n_power, "relaxation": omega, "literal": True, "precision": precision})
    vh, L = o.vec_height, o.vec_length
    o.set_matrix(g("A")).set_b(g("b"))
    o._f("orcj_setup")(C.c_int64(L), orc._p(o.A), orc._p(o.b), C.c_double(orc.tofixed20(o.omega)),
                       C.c_int(1 if o.omega == 1.0 else 0), orc._p(o.R), orc._p(o.Cv), C.c_int(1))
    # R texture (mat_height^2 RGBA) -> natural [row][col]: programR :238-243
    mat_h = 2 * vh * vh
    Rt = g("R").reshape(mat_h, mat_h, 4)
    Rnat = np.empty((L, L), Rt.dtype)
    for py in range(mat_h):
        for px in range(mat_h):
            row = px // vh + 2 * vh * (py // vh)
            col = 4 * (px % vh + vh * (py % vh))
            Rnat[row, col:col + 4] = Rt[py, px]
    assert_same(o.R, Rnat, "programR")
    assert_same(o.Cv, g("C").reshape(-1), "programC")
    o.x_guess = g("x").copy()
    x1 = o.mv_product()
    assert_same(x1, g("x1").reshape(-1), "mv_product, first iteration")
    o.x_guess = x1.copy()
    x2 = o.mv_product()
    assert_same(x2, g("x2").reshape(-1), "mv_product, second iteration")
    stats = np.zeros(L, o.dt)
    o._f("orcj_stats")(C.c_int64(L // 4), orc._p(x1), orc._p(x2), orc._p(stats))
    assert_same(stats, g("stats").reshape(-1), "programStats")
    if vh >= 2:  # the defect the literal mode keeps: x' is NOT omega (R x + C) + (1 - omega) x in natural row order
        intended = OracleSOR({"n_power": n_power, "relaxation": omega, "literal": False, "precision": precision})
        intended.R, intended.Cv, intended.x_guess = o.R, o.Cv, g("x").copy()
        assert not np.array_equal(intended.mv_product(), x1)


def test_jacobi_vectors_regenerate_from_the_reference_tree():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_reference_vectors_jacobi as gen
    if not os.path.exists(gen.REF_JS):
        pytest.skip("reference tree not present")
    d = np.load(os.path.join(HERE, "golden", "reference_glsl_jacobi.npz"))
    for name, dt in (("f64", np.float64), ("f32", np.float32)):
        for n_power, omega in gen.CASES:
            for k, v in gen.run_case(n_power, omega, dt).items():
                assert_same(v, d[f"{name}_p{n_power}_w{int(10 * omega)}_{k}"], f"{name} {n_power} {omega} {k}")


# ---- the host-side JavaScript of set(): inverse-cdf table (empic.js:1263-1339) -----------------------
def _digest(t):
    import hashlib
    c = np.array(t, np.float64)
    c[np.isnan(c)] = np.nan
    return np.frombuffer(hashlib.sha256(c.tobytes()).digest(), np.uint8)


def _pdfs():
    small = np.zeros((SPEC["nr"], SPEC["nz"]))
    small[0:3, 17:23] = 1.0
    demo = np.zeros((400, 800))  # fusionsim.js:116-122
    demo[0:50, 350:450] = 1.0
    return small, demo


def test_inverse_cdf_matches_the_reference_javascript():
    """The table the reference's own source_pdf code builds (its JavaScript transliterated mechanically
    and executed, tests/golden/js_transliterate.py) against the oracle's restatement: every one of the
    2 x 512^2 doubles, for the small scene and for the demo scene -- whose table holds exactly 1023 NaN
    texels, now confirmed by running the reference's code."""
    from oracle import oracle as orc
    d = load("f64")
    small, demo = _pdfs()
    assert_same(d["source_pdf"], small, "the pdf the vectors were made with")
    t = orc.inv_cdf(small)
    assert_same(t.reshape(512, 512, 2)[::5, ::5], d["invcdf_small_sub"], "inverse cdf (every 5th texel)")
    assert_same(_digest(t), d["invcdf_small_digest"], "inverse cdf, all texels (sha256)")
    t = orc.inv_cdf(demo)
    assert int(np.isnan(t).any(1).sum()) == int(d["invcdf_demo_nan_texels"]) == 1023
    assert_same(_digest(t), d["invcdf_demo_digest"], "demo-scene inverse cdf, all texels (sha256)")


def test_demo_scene_matches_the_reference_javascript():
    """fusion_sim_b200/scenes.py (config C1) against the scene construction of the reference's page
    (fusionsim.js:94-128), executed: sink mask, source pdf, and the particle formulas fed the same
    uniform numbers in the same order (x, y, z, vx, vy, vz per particle)."""
    from fusion_sim_b200.scenes import c1_sink_source
    d = load("f64")
    sink, source = c1_sink_source(400, 800)
    assert_same(np.packbits(sink.astype(np.uint8)), d["demo_sink_packed"], "sink mask")
    assert_same(np.packbits(source.astype(np.uint8)), d["demo_source_packed"], "source pdf")
    assert sink[0, 0] == 1 and sink[0, 799] == 1 and sink[399, 5] == 0 and sink[7, 0] == 0  # :105-112
    u = d["demo_uniforms"].reshape(-1, 6)
    position = 0.2 * (u[:, :3] - 0.5)   # scenes.c1_scene
    position[:, 2] += 1
    velocity = 0.002 * (u[:, 3:] - 0.5)
    assert_same(position, d["demo_position"], "initial positions (fusionsim.js:126)")
    assert_same(velocity, d["demo_velocity"], "initial velocities (fusionsim.js:127)")


@pytest.mark.gpu
def test_gpu_inverse_cdf_matches_the_reference_javascript():
    """fsim_set_source_pdf (api.cu: the product's own restatement) against the executed reference code."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    d = load("f64")
    small, demo = _pdfs()
    g = makeCylindricalParticlePusher(dict(SPEC, precision="f64"))
    g.set({"source_pdf": small})
    assert_same(_digest(g.getField("inv_cdf")), d["invcdf_small_digest"], "inverse cdf (sha256)")
    g = makeCylindricalParticlePusher(dict(SPEC, nr=400, nz=800, precision="f64"))
    g.set({"source_pdf": demo})
    assert_same(_digest(g.getField("inv_cdf")), d["invcdf_demo_digest"], "demo-scene inverse cdf (sha256)")


@pytest.mark.gpu
@pytest.mark.parametrize("precision", PRECISIONS)
def test_gpu_half_steps_match_the_reference_shaders(precision):
    """libfusionsim.so straight against the outputs of the reference's step shaders: the state the
    vectors start from (just-respawned particles included) is restored with fsim_set_state, then 8
    half-steps; same documented NaN deviation as in the oracle test above."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    d = load(precision)
    nr, nz = SPEC["nr"], SPEC["nz"]
    g = makeCylindricalParticlePusher(dict(SPEC, precision=precision))
    cells = lambda a: np.ascontiguousarray(a[:, :3].astype(np.float64).reshape(nz, nr, 3).transpose(1, 0, 2))
    g.set({"E": cells(d["E"]), "B": cells(d["B"]), "sink_mask": d["sink"][:, 0].astype(np.float64).reshape(nz, nr).T,
           "source_pdf": d["source_pdf"], "entropy": entropy_table()})
    g.precalc()
    for nm in ("R1", "R2", "R3", "A"):
        assert_same(g.getField(nm), d[nm][:, :3].astype(np.float64), nm)
    f64 = lambda a: a.astype(np.float64)
    g.setState(f64(d["position"]), f64(d["velocity"][:, :3]), f64(d["rand"]))
    live = np.ones(g.n, bool)
    for k in range(d["step_position"].shape[0]):
        g.half_step()
        ref_pos = f64(d["step_position"][k])
        pos = g.getPosition()
        nan_kept = live & (ref_pos[:, 3] == 1) & np.isnan(ref_pos[:, :3]).any(1)
        assert (pos[nan_kept, 3] == 0).all()
        live &= ~nan_kept
        assert_same(g.getRand()[live], f64(d["step_rand"][k])[live], f"half-step {k}: programStepRand")
        assert_same(g.getVelocity()[live], f64(d["step_velocity"][k])[live, :3], f"half-step {k}: step_velocity_frag")
        assert_same(pos[live], ref_pos[live], f"half-step {k}: step_position_frag")
        if (~live).any():  # keep the excluded particles on the reference's track
            p, v, r = pos.copy(), g.getVelocity(), g.getRand()
            p[~live], v[~live], r[~live] = ref_pos[~live], f64(d["step_velocity"][k])[~live, :3], f64(d["step_rand"][k])[~live]
            g.setState(p, v, r)
    assert live.sum() >= g.n - 8


# ---- a scene whose unit factors are not powers of two: the toFixed(20) literals round ---------------
SPEC_ODD = dict(SPEC, radius=0.7, height=1.3)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_odd_units_match_the_reference_shaders(precision):
    """radius = 0.7, height = 1.3: factor_r/factor_z, factor_z/factor_r, factor_r and factor_z enter the
    shaders as N(x) = x.toFixed(20) literals (empic.js:527,566,606,647) that are not dyadic any more (and
    round when the shader runs in fp32); fields, precalc and 8 half-steps must still be bit-identical to
    the executed shader text."""
    from oracle import oracle as orc
    from oracle.oracle import OraclePusher
    d = np.load(os.path.join(HERE, "golden", f"reference_glsl_odd_{precision}.npz"))
    o = OraclePusher(dict(SPEC_ODD, precision=precision))
    for r, z, I in d["loops"]:
        o.addCurrentLoop(float(r), float(z), float(I))
    cz, bz, bt = (float(v) for v in d["uniform_terms"])
    o.addCurrentZ(cz); o.addBZ(bz); o.addBTheta(bt)
    assert_same(o.B, d["B"], "B")
    o.E[:] = d["E"]
    o.precalc()
    for nm in ("R1", "R2", "R3", "A"):
        assert_same(getattr(o, nm), d[nm], "programPre" + nm[-1])
    assert o.factor_r / o.factor_z != 2.0 and orc.tofixed20(o.factor_r) == o.factor_r  # 20 decimals keep a double of this size
    o.sink_mask[:] = d["sink"]
    o.inv_cdf[:, :2] = orc.inv_cdf(d["source_pdf"])
    o.entropy[:] = entropy_table()
    o.position[:], o.velocity[:], o.rand[:] = d["position"], d["velocity"], d["rand"]
    live = np.ones(o.n, bool)
    for k in range(d["step_position"].shape[0]):
        o.half_step()
        ref_pos = d["step_position"][k]
        nan_kept = live & (ref_pos[:, 3] == 1) & np.isnan(ref_pos[:, :3]).any(1)  # documented deviation, see above
        live &= ~nan_kept
        assert_same(o.rand[live], d["step_rand"][k][live], f"half-step {k}: programStepRand")
        assert_same(o.velocity[live], d["step_velocity"][k][live], f"half-step {k}: step_velocity_frag")
        assert_same(o.position[live], ref_pos[live], f"half-step {k}: step_position_frag")
        o.position[~live], o.velocity[~live], o.rand[~live] = ref_pos[~live], d["step_velocity"][k][~live], d["step_rand"][k][~live]
    assert live.sum() >= o.n - 8


@pytest.mark.gpu
@pytest.mark.parametrize("precision", PRECISIONS)
def test_gpu_odd_units_match_the_reference_shaders(precision):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    d = np.load(os.path.join(HERE, "golden", f"reference_glsl_odd_{precision}.npz"))
    nr, nz = SPEC["nr"], SPEC["nz"]
    g = makeCylindricalParticlePusher(dict(SPEC_ODD, precision=precision))
    cells = lambda a: np.ascontiguousarray(a[:, :3].astype(np.float64).reshape(nz, nr, 3).transpose(1, 0, 2))
    g.set({"E": cells(d["E"]), "sink_mask": d["sink"][:, 0].astype(np.float64).reshape(nz, nr).T,
           "source_pdf": d["source_pdf"], "entropy": entropy_table()})
    for r, z, I in d["loops"]:
        g.addCurrentLoop(float(r), float(z), float(I))
    cz, bz, bt = (float(v) for v in d["uniform_terms"])
    g.addCurrentZ(cz); g.addBZ(bz); g.addBTheta(bt)
    assert_same(g.getField("B"), d["B"][:, :3].astype(np.float64), "B")
    g.precalc()
    for nm in ("R1", "R2", "R3", "A"):
        assert_same(g.getField(nm), d[nm][:, :3].astype(np.float64), nm)
    f64 = lambda a: a.astype(np.float64)
    g.setState(f64(d["position"]), f64(d["velocity"][:, :3]), f64(d["rand"]))
    live = np.ones(g.n, bool)
    for k in range(d["step_position"].shape[0]):
        g.half_step()
        ref_pos, pos = f64(d["step_position"][k]), g.getPosition()
        live &= ~(live & (ref_pos[:, 3] == 1) & np.isnan(ref_pos[:, :3]).any(1))
        assert_same(g.getRand()[live], f64(d["step_rand"][k])[live], f"half-step {k}: rand")
        assert_same(g.getVelocity()[live], f64(d["step_velocity"][k])[live, :3], f"half-step {k}: velocity")
        assert_same(pos[live], ref_pos[live], f"half-step {k}: position")
        if (~live).any():
            p, v, r = pos.copy(), g.getVelocity(), g.getRand()
            p[~live], v[~live], r[~live] = ref_pos[~live], f64(d["step_velocity"][k])[~live, :3], f64(d["step_rand"][k])[~live]
            g.setState(p, v, r)


# ---- out.set(value), empic.js:1157-1261: layout conversions and unit factors --------------------------
@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("scene", ["", "odd_"])
def test_set_conversions_match_the_reference_javascript(precision, scene):
    """The typed arrays the reference's set() fills (its JavaScript executed, Float32Array stores emulated
    in fp32 mode) against the oracle's set(): texel order i + j*nr, the unit factors, w = 1."""
    from oracle.oracle import OraclePusher
    d = np.load(os.path.join(HERE, "golden", f"reference_glsl_{scene}{precision}.npz"))
    o = OraclePusher(dict(SPEC_ODD if scene else SPEC, precision=precision))
    o.set({k: d["setin_" + k] for k in ("E", "B", "position", "velocity", "sink_mask")})
    assert_same(o.E, d["setout_E"], "E_arr")
    assert_same(o.B, d["setout_B"], "B_arr")
    assert_same(o.position, d["setout_position"], "position_arr")
    assert_same(o.velocity, d["setout_velocity"], "velocity_arr")
    assert_same(o.sink_mask[:, 0], d["setout_sink_mask"][:, 0], "sink_mask_arr")


@pytest.mark.gpu
@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("scene", ["", "odd_"])
def test_gpu_set_conversions_match_the_reference_javascript(precision, scene):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    d = np.load(os.path.join(HERE, "golden", f"reference_glsl_{scene}{precision}.npz"))
    g = makeCylindricalParticlePusher(dict(SPEC_ODD if scene else SPEC, precision=precision))
    g.set({k: d["setin_" + k] for k in ("E", "B", "position", "velocity", "sink_mask")})
    f64 = lambda a: a.astype(np.float64)
    assert_same(g.getField("E"), f64(d["setout_E"][:, :3]), "E")
    assert_same(g.getField("B"), f64(d["setout_B"][:, :3]), "B")
    assert_same(g.getPosition(), f64(d["setout_position"]), "position (w = alive = 1)")
    assert_same(g.getVelocity(), f64(d["setout_velocity"][:, :3]), "velocity")
    assert_same(g.getField("sink_mask"), (d["setout_sink_mask"][:, 0] > 0.5).astype(np.uint8), "sink mask")
