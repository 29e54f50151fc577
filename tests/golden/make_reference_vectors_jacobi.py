"""Generates tests/golden/reference_glsl_jacobi.npz by EXECUTING the shader source of
matrix_webgl.makeSORIterative (/root/reference/public/javascripts/matrix_webgl.js:35-711) with the
GLSL interpreter of oracle/glsl_interp.py -- row N3 of SURVEY.md section 8f.  Runs only where the
reference tree exists; nothing of its source is copied here.  Draw order and bindings as in
solve() :576-640 and mv_product() :539-562:
    programR(A) -> R;  programC(A, b) -> C;
    programMVproduct(R, x) -> n_power x sum_frag (2x2 texel sums) -> programResult(+ C [+ (1-w) x]) -> x';
    programStats(x, x') -> (x.x'/4, x.x/4, x'.x'/4, max|x'-x|).

    python tests/golden/make_reference_vectors_jacobi.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF_JS = "/root/reference/public/javascripts/matrix_webgl.js"

from js_shader_source import shader_sources  # noqa: E402
from oracle.glsl_interp import Shader, Texture  # noqa: E402

CASES = [(1, 1.0), (2, 1.0), (2, 0.8)]  # (n_power, relaxation)


def quad(w, h, dtype):
    T = np.dtype(dtype).type
    i = np.tile(np.arange(w), h).astype(dtype)
    j = np.repeat(np.arange(h), w).astype(dtype)
    return np.stack([(i + T(0.5)) / T(w), (j + T(0.5)) / T(h)], 1)


def run_case(n_power, omega, dtype):
    dtype = np.dtype(dtype)
    vh = 2 ** n_power
    L = 4 * vh * vh
    mat_h = 2 * vh * vh
    env = dict(vec_height=vh, vec_length=L, mat_height=mat_h, omega=omega)
    src = shader_sources(REF_JS, env)
    sh = lambda text: Shader(text, dtype)
    frag = lambda prog, coords, **u: prog.run(len(coords), dict(u, v_texCoord=coords))["gl_FragColor"]
    rng = np.random.Generator(np.random.PCG64(100 * n_power + int(10 * omega)))
    A = rng.random((L, L)) - 0.5
    A[np.arange(L), np.arange(L)] = 1.5 * np.abs(A).sum(1)
    b, x = rng.random(L), rng.random(L)
    A, b, x = A.astype(dtype), b.astype(dtype), x.astype(dtype)
    a_tex = np.zeros((L, L, 4), dtype)
    a_tex[:, :, 0] = A                                   # m_set_arr[4 (col + row L)] = matrix[row][col], :462-466
    tA = Texture(a_tex)
    tb, tx = Texture(b.reshape(vh, vh, 4)), Texture(x.reshape(vh, vh, 4))
    R = frag(sh(src["programR"][0]), quad(mat_h, mat_h, dtype), u_A=tA)
    C = frag(sh(src["programC"][0]), quad(vh, vh, dtype), u_A=tA, u_b=tb)
    tR, tC = Texture(R, mat_h, mat_h), Texture(C, vh, vh)

    def mv_product(xt):
        cur, size = frag(sh(src["programMVproduct"][0]), quad(mat_h, mat_h, dtype), u_M=tR, u_v=xt), mat_h
        for i in range(n_power):
            num_x = mat_h / 2 ** i                      # sum_frag(mat_height / Math.pow(2, i)), :374
            text = shader_sources(REF_JS, env, {"sum_frag": {"num_x": num_x}})["sum_frag"][0]
            half = size // 2
            cur = frag(sh(text), quad(half, half, dtype), u_M=Texture(cur, size, size))
            size = half
        return frag(sh(src["programResult"][0]), quad(vh, vh, dtype), u_Vsum=Texture(cur, size, size), u_C=tC, u_X=xt)

    x1 = mv_product(tx)
    x2 = mv_product(Texture(x1, vh, vh))
    stats = frag(sh(src["programStats"][0]), quad(vh, vh, dtype), u_X1=Texture(x1, vh, vh), u_X2=Texture(x2, vh, vh))
    return dict(A=A, b=b, x=x, R=R, C=C, x1=x1, x2=x2, stats=stats)


if __name__ == "__main__":
    if not os.path.exists(REF_JS):
        sys.exit("the reference tree is not present: the committed vectors cannot be regenerated here")
    out = {}
    for name, dt in (("f64", np.float64), ("f32", np.float32)):
        for n_power, omega in CASES:
            for k, v in run_case(n_power, omega, dt).items():
                out[f"{name}_p{n_power}_w{int(10 * omega)}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "reference_glsl_jacobi.npz"), **out)
    print("wrote", len(out), "arrays")
