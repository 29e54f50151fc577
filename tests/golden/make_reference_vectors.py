"""Generates tests/golden/reference_glsl_{f64,f32}.npz by EXECUTING THE REFERENCE'S OWN SHADER SOURCE.

Runs only where /root/reference exists (this container).  The GLSL text of every program on the
hot path is read out of /root/reference/public/javascripts/empic.js at run time (string arrays,
with the `N(expr)` = expr.toFixed(20) insertions of empic.js:23-25 evaluated as the JS would) and
executed fragment by fragment with the interpreter in oracle/glsl_interp.py.  Nothing of the
reference's source is copied into this repository: only the numeric inputs and outputs of that
execution are saved, and tests/test_reference_glsl.py holds the CPU oracle to them bit for bit.

The orchestration comes from the reference too: which program draws when and into which buffer
(out.precalc :1413-1434, out.step :1436-1469, the quad draws of out.density :1480-1504) and which buffer
feeds which uniform (the `.set({...})` after every `webgl.linkProgram`, :506-659, :783-928, :1042-1116) are
PARSED out of empic.js (js_shader_source.draw_sequence / program_bindings).  Restated by hand, with the
lines cited: the uniforms addCurrentLoop/Z/BZ/BTheta set per call (:1352-1411) and the point-sprite draw.
Fixed-function GL behaviour (additive blending ONE,ONE = a floating-point add in draw order, NEAREST
+ CLAMP_TO_EDGE sampling, the identity quad mapping of empic.js:66-92, GLES2 point-sprite coverage)
is not shader text; it is emulated here as the oracle documents it.

    python tests/golden/make_reference_vectors.py
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF_JS = "/root/reference/public/javascripts/empic.js"

from oracle.glsl_interp import Shader, Texture  # noqa: E402

C_LIGHT = 2.998e8  # empic.js:27
SPEC = dict(radius=1.0, height=2.0, nr=24, nz=40, dt=2e-9, nparticles=16,
            particle_mass=1.67e-27, particle_charge=1.602e-19)


from js_shader_source import JsExpr, draw_sequence, program_bindings  # noqa: E402
from js_shader_source import shader_sources as _shader_sources  # noqa: E402


def shader_sources(env):
    return _shader_sources(REF_JS, env)


# ---- the host-side JavaScript of set(): inverse-cdf table (empic.js:1263-1339), executed ----------
def inverse_cdf_js(pdf):
    """Runs the reference's own source_pdf code (transliterated mechanically, js_transliterate.py) and
    returns the [512*512][2] table of (x, y) as doubles -- the values BEFORE the Float32Array store."""
    import types
    import js_transliterate as jt
    text = open(REF_JS).read()
    a = text.index("if (value.source_pdf) {")
    a = text.index("\n", a) + 1
    b = text.index("inv_cdf_tex.update();", a)
    rows = [[np.float64(v) for v in row] for row in np.asarray(pdf, np.float64)]
    arr = np.zeros(4 * 512 * 512, np.float64)
    with np.errstate(all="ignore"):
        jt.run(text[a:b], dict(value=types.SimpleNamespace(source_pdf=rows), inv_cdf_arr=arr))
    return arr.reshape(-1, 4)[:, :2].copy()


FUSIONSIM_JS = "/root/reference/public/javascripts/fusionsim.js"


def demo_scene_js(nparticles, uniforms):
    """The scene construction of the reference's page (fusionsim.js:94-128: sink mask, source pdf,
    initial positions and velocities), executed; Math.random() returns `uniforms` in call order."""
    import types
    import js_transliterate as jt
    text = open(FUSIONSIM_JS).read()
    a = text.index("for(i = 0; i < spec.nr; i++) {")
    b = text.index("simulation.set({", a)
    it = iter(uniforms)
    spec = types.SimpleNamespace(nr=400, nz=800)
    scope = jt.run(text[a:b], dict(spec=spec, nparticles=nparticles, js_random=lambda: float(next(it)),
                                   source=jt.JsArray(), sink=jt.JsArray(), init_position=jt.JsArray(),
                                   init_velocity=jt.JsArray()))
    return (np.array(scope["sink"], np.float64), np.array(scope["source"], np.float64),
            np.array(scope["init_position"], np.float64), np.array(scope["init_velocity"], np.float64))


def set_js(sp, dtype, value):
    """out.set(value) of the reference (empic.js:1157-1261: E, B, position, velocity, sink_mask), executed.
    The typed arrays are NumPy arrays of `dtype`: float32 reproduces the Float32Array stores, float64 shows
    the products before that rounding.  GL plumbing lines (tex.update(), programSet...) are dropped."""
    import types
    import js_transliterate as jt
    text = open(REF_JS).read()
    a = text.index("out.set = function(value) {")
    a = text.index("\n", a) + 1
    b = text.index("if (value.source_pdf) {", a)
    body = re.sub(r"programSet\.draw\(\{.*?\}\);", "", text[a:b], flags=re.S)
    body = re.sub(r"^\s*(\w+_tex\.update\(\);|programSet\.set\(.*\);)\s*$", "", body, flags=re.M)
    nr, nz, n = sp["nr"], sp["nz"], sp["nparticles"] ** 2
    arrs = {k: np.zeros(4 * (nr * nz if k in ("E_arr", "B_arr", "sink_mask_arr") else n), dtype)
            for k in ("E_arr", "B_arr", "position_arr", "velocity_arr", "sink_mask_arr")}
    ns = types.SimpleNamespace(E=None, B=None, position=None, velocity=None, sink_mask=None, source_pdf=None)
    for k, v in value.items():
        setattr(ns, k, [[list(map(np.float64, c)) if np.ndim(c) else np.float64(c) for c in row] for row in v]
                if np.ndim(v) == 3 else [list(map(np.float64, row)) for row in v])
    jt.run(body, dict(value=ns, spec=types.SimpleNamespace(nr=nr, nz=nz), nparticles=n, i=0, j=0,
                      factor_r=1 / sp["radius"], factor_z=1 / sp["height"], **arrs))
    return {k[:-4]: v.reshape(-1, 4) for k, v in arrs.items()}


def table_digest(t):
    import hashlib
    c = np.array(t, np.float64)
    c[np.isnan(c)] = np.nan  # one NaN bit pattern
    return np.frombuffer(hashlib.sha256(c.tobytes()).digest(), np.uint8).copy()


def demo_pdf():
    source = np.zeros((400, 800))  # fusionsim.js:116-122
    source[0:50, 350:450] = 1.0
    return source


# ---- the scene ------------------------------------------------------------------------------------
def entropy_table():
    k = np.arange(1024 * 1024 * 4, dtype=np.float64)
    return np.mod(k * 0.6180339887498949, 1.0).reshape(-1, 4)  # no 32 MB fixture: a formula both sides know


def scene(dtype):
    rng = np.random.Generator(np.random.PCG64(20261018))
    nr, nz, side = SPEC["nr"], SPEC["nz"], SPEC["nparticles"]
    n, nc = side * side, nr * nz
    pos = np.ones((n, 4))
    pos[:, 0] = 0.55 * (rng.random(n) - 0.5) * 2
    pos[:, 1] = 0.55 * (rng.random(n) - 0.5) * 2
    pos[:, 2] = 0.5 + 0.45 * (rng.random(n) - 0.5) * 2
    pos[:8, 3] = 0.0                      # "just respawned": fresh random velocity (empic.js:772)
    pos[8, 0] = pos[8, 1] = 0.0           # r = 0: direction = 0/0
    vel = np.ones((n, 4))
    vel[:, :3] = 0.3 * (rng.random((n, 3)) - 0.5)   # fast: absorption and respawn within a few half-steps
    rnd = rng.random((n, 4))
    E = np.ones((nc, 4))
    E[:, :3] = 1.0e5 * (rng.random((nc, 3)) - 0.5)
    sink = np.ones((nr, nz))
    sink[nr - 1, :] = 0
    sink[1:nr - 1, 0] = 0
    sink[1:nr - 1, nz - 1] = 0
    sink_tex = np.zeros((nc, 4))
    sink_tex[:, 0] = sink.T.reshape(-1)
    source = np.zeros((nr, nz))
    source[0:3, 17:23] = 1.0
    mom = np.zeros((nc, 4))
    mom[:, :3] = rng.random((nc, 3)) - 0.5
    mom[:, 3] = np.where(rng.random(nc) > 0.3, rng.random(nc), 0.0)
    avg = rng.random((nc, 4))
    cast = lambda a: a.astype(dtype)
    return dict(position=cast(pos), velocity=cast(vel), rand=cast(rnd), E=cast(E), sink=cast(sink_tex),
                source_pdf=source, moments01=cast(mom), avg0=cast(avg))


def quad_coords(w, h, dtype):
    """v_texCoord of the fragment at texel (i, j): ((i+.5)/w, (j+.5)/h), fragments in texel order i + j*w."""
    T = np.dtype(dtype).type
    i = np.tile(np.arange(w), h).astype(dtype)
    j = np.repeat(np.arange(h), w).astype(dtype)
    return np.stack([(i + T(0.5)) / T(w), (j + T(0.5)) / T(h)], 1)


# a second scene whose unit factors are NOT powers of two: N(factor_r/factor_z).toFixed(20) etc. then
# round the literals the shaders multiply with (empic.js:527,566,606,647), which SPEC cannot show
SPEC_ODD = dict(SPEC, radius=0.7, height=1.3)


def run(dtype, full=False, sp=None):
    from oracle import oracle as orc
    dtype = np.dtype(dtype)
    odd = sp is not None
    sp = SPEC if sp is None else sp
    nr, nz, side = sp["nr"], sp["nz"], sp["nparticles"]
    n, nc = side * side, nr * nz
    h = sp["particle_charge"] * sp["dt"] / (2 * sp["particle_mass"])  # empic.js:44
    env = dict(factor_r=1 / sp["radius"], factor_z=1 / sp["height"], h=h, speed_of_light=C_LIGHT)
    src = shader_sources(env)
    sh = lambda name, k=0: Shader(src[name][k], dtype)
    sc = scene(dtype)
    out = {k: v for k, v in sc.items()}
    grid = quad_coords(nr, nz, dtype)
    part = quad_coords(side, side, dtype)
    tex = lambda a, w, hh: Texture(a, w, hh)
    frag = lambda prog, coords, **u: prog.run(len(coords), dict(u, v_texCoord=coords))["gl_FragColor"]

    # -- static field: loop tables (:295-345), two opposing loops (:1352-1363, blend ONE,ONE), uniform terms
    shape_prog = sh("programCurrentLoopShape")
    half = frag(shape_prog, grid, u_R=0.5)
    tenth = frag(shape_prog, grid, u_R=0.1)
    out["loop_half"], out["loop_tenth"] = half, tenth
    B = np.zeros((nc, 4), dtype)
    loops = [(0.8, 2.0, -1.0e7), (0.8, 0.0, 1.0e7)]
    for r, z, I in loops:
        B = B + frag(sh("programCurrentLoop"), grid, u_R=r * env["factor_r"], u_Z=z * env["factor_z"], u_I=I,
                     u_shape_half=tex(half, nr, nz), u_shape_tenth=tex(tenth, nr, nz))
    out["B_loops"] = B.copy()
    # gl_FragColor += on an unwritten colour: taken as "=" (documented), then blended into B
    B = B + frag(sh("programCurrentZ"), grid, u_I=3.0e5)
    B = B + frag(sh("programBZ"), grid, u_Bz=0.02)
    B = B + frag(sh("programBTheta"), grid, u_Btheta=-0.01)
    out["B"] = B
    out["uniform_terms"] = np.array([3.0e5, 0.02, -0.01])
    out["loops"] = np.array(loops)

    # -- the draw calls of out.precalc (:1413-1434) and out.step (:1436-1469) are PARSED out of the reference:
    # which program runs when, into which buffer, and which buffer feeds which uniform (:506-659, :783-928)
    bind = program_bindings(REF_JS)
    jsenv = {"spec.dt": sp["dt"], "speed_of_light": C_LIGHT, "h": h}
    invcdf = np.zeros((512 * 512, 4), dtype)
    invcdf[:, :2] = orc.inv_cdf(sc["source_pdf"]).astype(dtype)  # host JS code (:1268-1339), an INPUT here
    zeros = lambda m: np.zeros((m, 4), dtype)
    buf = {"B": (B, nr, nz), "E": (sc["E"], nr, nz), "R1": (zeros(nc), nr, nz), "R2": (zeros(nc), nr, nz),
           "R3": (zeros(nc), nr, nz), "A": (zeros(nc), nr, nz), "sink_mask": (sc["sink"], nr, nz),
           "inv_cdf": (invcdf, 512, 512), "entropy_tex": (entropy_table().astype(dtype), 1024, 1024),
           "position_A": (sc["position"], side, side), "velocity_A": (sc["velocity"], side, side),
           "rand_A": (sc["rand"], side, side), "position_B": (zeros(n), side, side),
           "velocity_B": (zeros(n), side, side), "rand_B": (zeros(n), side, side)}

    def draw(program, target, extra_uniforms=None):
        """One `program.draw({triangles: 6, target})`: the quad covers the target, one fragment per texel."""
        info = bind[program]
        uniforms = dict(info["uniforms"], **(extra_uniforms or {}))
        args = {}
        for name, js in uniforms.items():
            if name.startswith("a_"):
                continue  # the quad's vertex attributes
            if js in buf:
                args[name] = tex(*buf[js])
            else:
                args[name] = JsExpr(js, jsenv).value()
        _, w, hh = buf[target]
        text = src[info["fragment"]][-1]
        buf[target] = (frag(Shader(text, dtype), quad_coords(w, hh, dtype), **args), w, hh)

    for program, opts, sets in draw_sequence(REF_JS, "precalc"):
        draw(program, opts["target"], sets)
    R1, R2, R3, A = (buf[k][0] for k in ("R1", "R2", "R3", "A"))
    out.update(R1=R1, R2=R2, R3=R3, A=A)

    assert src["programStepRandA"] == src["programStepRandB"]
    step_seq = draw_sequence(REF_JS, "step")
    assert len(step_seq) == 6
    states = []
    for k in range(4):                      # four out.step() = eight half-steps
        for q, (program, opts, sets) in enumerate(step_seq):
            draw(program, opts["target"], sets)
            if q % 3 == 2:                  # a half-step is complete: its three targets are the new state
                suffix = step_seq[q][1]["target"][-2:]
                states.append(tuple(buf[nm + suffix][0].copy() for nm in ("position", "velocity", "rand")))
    pos, vel, rnd = states[-1]
    out["step_position"] = np.stack([s[0] for s in states])
    out["step_velocity"] = np.stack([s[1] for s in states])
    out["step_rand"] = np.stack([s[2] for s in states])

    # -- density (:1471-1495): vertex shader of the sprites, normalise, running average
    vert = sh("programMoments01", 0).run(n, dict(a_particleTexCoord=part, u_position=tex(pos, side, side),
                                                 u_velocity=tex(vel, side, side), u_pointsize=11.0),
                                         outputs=("gl_Position", "gl_PointSize", "v_color"))
    out["sprite_color"], out["sprite_position"] = vert["v_color"], vert["gl_Position"]
    # fragment shader of the sprites (:1022) on every pixel GLES2 lets a size-11 sprite cover, additive
    # blend in particle order; window coordinates from the vertex shader's gl_Position as the viewport maps them
    shape = orc.shape_table(dtype == np.float32).astype(dtype)
    t_shape = tex(np.repeat(shape[:, None], 4, 1), 11, 11)
    frag_prog = sh("programMoments01", 1)
    mom = np.zeros((nz, nr, 4), dtype)
    T = dtype.type
    for p in range(n):
        r = np.sqrt(pos[p, 0] * pos[p, 0] + pos[p, 1] * pos[p, 1])
        xw, yw = r * T(nr), pos[p, 2] * T(nz)   # (ndc + 1)/2 * size with ndc = 2 r - 1 as written at :997
        if not (xw >= 0 and xw < nr and yw >= 0 and yw < nz):
            continue
        xs = [x for x in range(int(np.floor(xw - T(5.5))) - 1, int(np.floor(xw - T(5.5))) + 14)
              if 0 <= x < nr and (T(x) + T(0.5)) >= xw - T(5.5) and (T(x) + T(0.5)) < xw + T(5.5)]
        ys = [y for y in range(int(np.floor(yw - T(5.5))) - 1, int(np.floor(yw - T(5.5))) + 14)
              if 0 <= y < nz and (T(y) + T(0.5)) >= yw - T(5.5) and (T(y) + T(0.5)) < yw + T(5.5)]
        if not xs or not ys:
            continue
        px = np.array([[x, y] for y in ys for x in xs])
        pc = np.stack([T(0.5) + ((px[:, 0].astype(dtype) + T(0.5)) - xw) / T(11),
                       T(0.5) + ((px[:, 1].astype(dtype) + T(0.5)) - yw) / T(11)], 1)
        col = frag_prog.run(len(px), dict(v_color=np.repeat(vert["v_color"][p:p + 1], len(px), 0), u_shape=t_shape,
                                          gl_PointCoord=pc))["gl_FragColor"]
        for (x, y), c in zip(px, col):
            mom[y, x] = mom[y, x] + c
    out["sprite_moments01"] = mom.reshape(nc, 4)
    # the quad draws of out.density (:1480-1495), parsed: normalise -> running average -> copy avgA to avgB
    buf.update(moments01=(sc["moments01"], nr, nz), moments01_norm=(zeros(nc), nr, nz),
               moments01_avgA=(zeros(nc), nr, nz), moments01_avgB=(sc["avg0"], nr, nz))
    dens_seq = draw_sequence(REF_JS, "density")
    assert [d[0] for d in dens_seq] == ["programMoments01", "programNormalizeMoments01", "programAvgMoments",
                                        "programSet", "programBMag", "programDensity"]
    for program, opts, sets in dens_seq[1:4]:
        draw(program, opts["target"], sets)
    norm, avg = buf["moments01_norm"][0], buf["moments01_avgA"][0]
    assert np.array_equal(buf["moments01_avgB"][0], avg, equal_nan=True)  # programSet: a copy
    out.update(moments01_norm=norm, moments01_avg=avg)

    # -- set({source_pdf}) (:1263-1339): the reference's host-side JavaScript, executed
    if dtype == np.float64 and not odd:
        small = inverse_cdf_js(sc["source_pdf"])
        out["invcdf_small_digest"] = table_digest(small)
        out["invcdf_small_sub"] = small.reshape(512, 512, 2)[::5, ::5].copy()
        u = np.random.Generator(np.random.PCG64(7)).random(6 * 500)
        sink, source, p0, v0 = demo_scene_js(500, u)
        out.update(demo_uniforms=u, demo_position=p0, demo_velocity=v0,
                   demo_sink_packed=np.packbits(sink.astype(np.uint8)), demo_source_packed=np.packbits(source.astype(np.uint8)))
        assert set(np.unique(sink)) <= {0.0, 1.0} and set(np.unique(source)) <= {0.0, 1.0}
        if full:  # the demo scene's pdf: ~2 minutes of interpreted loops, not part of the regeneration check
            demo = inverse_cdf_js(demo_pdf())
            out["invcdf_demo_digest"] = table_digest(demo)
            out["invcdf_demo_nan_texels"] = np.array(int(np.isnan(demo).any(1).sum()))

    # -- set({E, B, position, velocity, sink_mask}) (:1157-1261): the reference's host-side JavaScript, executed
    rng = np.random.Generator(np.random.PCG64(99))
    setv = dict(E=rng.random((nr, nz, 3)) - 0.5, B=rng.random((nr, nz, 3)) - 0.5, position=rng.random((n, 3)) - 0.5,
                velocity=1e-3 * (rng.random((n, 3)) - 0.5), sink_mask=(rng.random((nr, nz)) > 0.3).astype(np.float64))
    for k, v in setv.items():
        out["setin_" + k] = v
    for k, v in set_js(sp, dtype, setv).items():
        out["setout_" + k] = v

    # -- canvas (:1497-1504): programBMag, then programDensity blended SRC_ALPHA,ONE into the RGBA8 canvas.
    # The two colours are the reference's shader text; clamping to [0,1], rounding to k/255 and the blend are
    # fixed-function GL, emulated as the oracle documents them; canvas rows run top to bottom.
    def canvas_colour(program):  # a draw without target: the canvas, nr x nz; bindings parsed like the others
        info = bind[program]
        args = {name: tex(*buf[js]) for name, js in info["uniforms"].items() if not name.startswith("a_")}
        return frag(Shader(src[info["fragment"]][-1], dtype), grid, **args)

    c1, c2 = canvas_colour(dens_seq[4][0]), canvas_colour(dens_seq[5][0])
    assert "SRC_ALPHA" in dens_seq[5][1]["blend"] and "blend" not in dens_seq[4][1]
    out["bmag_color"], out["density_color"] = c1, c2
    T = dtype.type
    with np.errstate(invalid="ignore"):
        clamp = lambda v: np.where(v > 0, np.where(v > 1, T(1), v), T(0)).astype(dtype)
        quant = lambda v: np.floor(v * T(255) + T(0.5))
        dst = quant(clamp(c1)) / T(255)
        blended = clamp(c2) * clamp(c2[:, 3:4]) + dst
        img = quant(clamp(blended)).astype(np.uint8).reshape(nz, nr, 4)
    out["canvas"] = img[::-1].copy()
    return out


if __name__ == "__main__":
    if not os.path.exists(REF_JS):
        sys.exit("the reference tree is not present: the committed vectors cannot be regenerated here")
    for name, dt in (("f64", np.float64), ("f32", np.float32)):
        res = run(dt, full=True)
        np.savez_compressed(os.path.join(HERE, f"reference_glsl_{name}.npz"), **res)
        print("wrote", name, {k: v.shape for k, v in res.items() if hasattr(v, "shape")})
        res = run(dt, sp=SPEC_ODD)
        keep = ("position", "velocity", "rand", "E", "sink", "source_pdf", "B", "loops", "uniform_terms", "R1", "R2", "R3", "A",
                "step_position", "step_velocity", "step_rand")
        np.savez_compressed(os.path.join(HERE, f"reference_glsl_odd_{name}.npz"),
                            **{k: v for k, v in res.items() if k in keep or k.startswith(("setin_", "setout_"))})
