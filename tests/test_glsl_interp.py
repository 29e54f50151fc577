"""Self-tests of the test tooling that executes the reference's code: the GLSL ES 1.00 subset interpreter
(oracle/glsl_interp.py), the JS shader-source reader and the JS transliterator (tests/golden/).  Small
hand-written programs with known answers; nothing here needs the reference tree or a GPU."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from oracle.glsl_interp import Shader, Texture  # noqa: E402


def run(src, n=4, dtype=np.float64, **inputs):
    return Shader(src, dtype).run(n, inputs)["gl_FragColor"]


def test_swizzles_constructors_and_left_to_right_arithmetic():
    src = """precision highp float;
    uniform float u_a;
    varying vec2 v_texCoord;
    void main() {
        vec4 p = vec4(v_texCoord, 3.0, 4.0);
        vec3 q = p.zyx * 2.0 - vec3(1.0);
        float s = 0.1 + 0.2 + 0.3;           // ((0.1 + 0.2) + 0.3), not 0.1 + (0.2 + 0.3)
        gl_FragColor = vec4(q.x, q.yz, s) + u_a * vec4(0.0, 0.0, 0.0, 1.0);
    }"""
    tc = np.array([[0.5, 0.25]] * 4)
    out = run(src, v_texCoord=tc, u_a=10.0)
    assert np.array_equal(out[0], [5.0, -0.5, 0.0, (0.1 + 0.2) + 0.3 + 10.0])
    assert (0.1 + 0.2) + 0.3 != 0.1 + (0.2 + 0.3)  # the order is observable


def test_select_if_else_masks_and_compound_assignment():
    src = """precision highp float;
    varying vec2 v_texCoord;
    void main() {
        float x = v_texCoord.x;
        vec4 c = vec4(0.0);
        if (x > 0.5 || x < 0.2) { c = vec4(1.0, 2.0, 3.0, 4.0); c.y += 10.0; } else { c.zw = vec2(7.0, 8.0); }
        gl_FragColor = (x > 0.5) ? c : -c;
    }"""
    tc = np.array([[0.1, 0], [0.3, 0], [0.6, 0], [0.9, 0]], float)
    out = run(src, v_texCoord=tc)
    assert np.array_equal(out[0], [-1, -12, -3, -4]) and np.array_equal(out[2], [1, 12, 3, 4])
    assert np.array_equal(out[1], [0, 0, -7, -8])


def test_for_loop_with_float_counter_and_builtins():
    src = """precision highp float;
    varying vec2 v_texCoord;
    void main() {
        float acc = 0.0;
        for(float i = 0.0; i < 10.0; i++){ acc += i * v_texCoord.x; }
        vec3 a = vec3(1.0, 2.0, 3.0);
        vec3 b = vec3(-2.0, 0.5, 4.0);
        gl_FragColor = vec4(acc, dot(a, b), length(vec2(3.0, 4.0)), cross(a, b).x + sign(-2.0) + abs(-1.5) + max(1.0, 2.0) + min(1.0, 2.0) + mod(7.5, 2.0) + floor(2.7));
    }"""
    out = run(src, v_texCoord=np.array([[2.0, 0]] * 4))
    assert np.array_equal(out[0], [90.0, ((1 * -2.0 + 2 * 0.5) + 3 * 4.0), 5.0, (2 * 4.0 - 3 * 0.5) - 1 + 1.5 + 2 + 1 + 1.5 + 2])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_texture_sampling_nearest_clamp_and_nan(dtype):
    tex = Texture(np.arange(2 * 3 * 4, dtype=dtype).reshape(2, 3, 4))  # h = 2, w = 3; texel (i, j) = data[j, i]
    src = """uniform sampler2D u_t; varying vec2 v_texCoord;
    void main() { gl_FragColor = texture2D(u_t, v_texCoord); }"""
    tc = np.array([[0.0, 0.0], [0.34, 0.6], [1.0, 1.0], [5.0, -3.0], [np.nan, 0.9]], dtype)
    out = Shader(src, dtype).run(5, {"u_t": tex, "v_texCoord": tc})["gl_FragColor"]
    assert out.dtype == dtype
    assert out[0, 0] == 0 and out[1, 0] == tex.data[1, 1, 0] and out[2, 0] == tex.data[1, 2, 0]
    assert out[3, 0] == tex.data[0, 2, 0]          # clamped to the edge texels
    assert out[4, 0] == tex.data[1, 0, 0]          # NaN coordinate -> texel 0 along that axis


def test_float32_mode_rounds_every_operation():
    src = """varying vec2 v_texCoord; void main() { gl_FragColor = vec4(v_texCoord.x * 3.0 / 7.0 + 0.1); }"""
    x = np.float32(0.123456789)
    out = Shader(src, np.float32).run(1, {"v_texCoord": np.array([[x, 0]], np.float32)})["gl_FragColor"]
    assert out[0, 0] == np.float32(np.float32(np.float32(x * np.float32(3.0)) / np.float32(7.0)) + np.float32(0.1))


def test_js_expression_evaluator_and_array_reader(tmp_path):
    import js_shader_source as js
    f = tmp_path / "x.js"
    f.write_text('''var N = function(n) { return n.toFixed(20); };
        var programA = webgl.linkProgram({ fragmentShaderSource : (function() {
            var src_arr = [
                "precision highp float;",   // a comment, with "quotes" inside
                "void main() {",
                    "gl_FragColor = " + ((omega !== 1.0) ? N(omega) + " * " : "") + "vec4(" + N(2 * k) + ");",
                "}"
            ];
            return src_arr.join('\\n'); })()
        }).set({ "u_x" : buffer_A, "u_s" : spec.dt * c });
        out.step = function () { programA.draw({ triangles : 6, target : buffer_B }); /* programB.draw({target : nope}); */ };''')
    src = js.shader_sources(str(f), {"omega": 0.8, "k": 0.25})
    assert src["programA"][0].splitlines()[2] == "gl_FragColor = 0.80000000000000004441 * vec4(0.50000000000000000000);"
    assert js.shader_sources(str(f), {"omega": 1.0, "k": 0.25})["programA"][0].splitlines()[2].startswith("gl_FragColor = vec4(")
    b = js.program_bindings(str(f))["programA"]
    assert b["uniforms"] == {"u_x": "buffer_A", "u_s": "spec.dt * c"} and b["fragment"] == "programA"
    assert js.draw_sequence(str(f), "step") == [("programA", {"triangles": "6", "target": "buffer_B"}, {})]
    assert js.JsExpr("spec.dt * c", {"spec.dt": 2.0, "c": 3.0}).value() == 6.0


def test_js_transliterator_semantics():
    import js_transliterate as jt
    js = """
        var a = [];
        var u = [];
        var total = 0;
        a[0] = 0.25;
        for(i = 1; i < n; i++) {
            a[i] = i * 0.5;
            total += a[i];
        }
        var f = function(x) {
            var k = 0;
            while(a[k] < x) {
                k++;
            }
            if (k === 0) {
                return x / a[0];
            }
            return Math.min(k, Math.floor(7.9));
        };
        var r0 = f(0.0);
        var r1 = f(1.2);
        var beyond = a[99];
        var poisoned = 1.0 / u[0];
    """
    with np.errstate(all="ignore"):
        s = jt.run(js, {"n": 5, "i": 0})
    assert s["total"] == 0.5 + 1.0 + 1.5 + 2.0
    assert s["r0"] == 0.0 and s["r1"] == 3       # the while loop stops at a[3] = 1.5; Math.min(3, Math.floor(7.9))
    assert s["beyond"] != s["beyond"]            # reading past the end gives undefined (NaN here), not an exception
    assert s["poisoned"] != s["poisoned"]        # ... and undefined poisons arithmetic like NaN
