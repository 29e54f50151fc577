"""The C-ABI boundary (no GPU): the shared library loads, exports every symbol include/fusionsim.h
declares, the header is valid C whose struct layout matches the ctypes binding, and without a
CUDA device the product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fusionsim.h")


@pytest.fixture(scope="module")
def built():
    from fusion_sim_b200.build import build
    return build()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsim_[A-Za-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built):
    lib = C.CDLL(built)
    names = declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    from fusion_sim_b200._lib import _SIGS
    assert sorted(_SIGS) == names  # the ctypes binding covers the whole header, nothing else


def test_header_is_plain_c_and_struct_layout_matches(built, tmp_path):
    from fusion_sim_b200._lib import FsimSpec
    fields = [f[0] for f in FsimSpec._fields_]
    prog = "#include <stdio.h>\n#include <stddef.h>\n#include \"fusionsim.h\"\n#include \"fsim_constants.h\"\n" \
           "int main(void){printf(\"%zu\", sizeof(fsim_spec));\n" + \
           "".join(f'printf(" %zu", offsetof(fsim_spec, {f}));\n' for f in fields) + "return 0;}\n"
    c = tmp_path / "t.c"
    c.write_text(prog)
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(c), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(FsimSpec)
    assert out[1:] == [getattr(FsimSpec, f).offset for f in fields]


def test_abi_version_and_error_string(built):
    from fusion_sim_b200._lib import lib
    assert lib().fsim_abi_version() == 1
    assert isinstance(lib().fsim_last_error(), bytes)


def test_no_cpu_fallback(built):
    """On a machine without a GPU the constructor must throw, as the reference throws when WebGL
    float textures are missing (utilities.js:493-495)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fusion_sim_b200 import Error, makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import C1_SPEC
    with pytest.raises(Error, match="no CPU fallback"):
        makeCylindricalParticlePusher(C1_SPEC)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "fusion_sim_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".js", ".cc")) or fn == "Makefile":
                txt = open(os.path.join(dp, fn), errors="replace").read()
                for pat in ("import oracle", "from oracle", "oracle/", "oracle.oracle", "oracle.numpy_ref", "libfsim_oracle", "orc_"):
                    assert pat not in txt, (os.path.join(dp, fn), pat)
    code = "import sys; import fusion_sim_b200, fusion_sim_b200.scenes, fusion_sim_b200.pusher; " \
           "assert not [m for m in sys.modules if m.startswith('oracle')]"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def _uniform01(k):
    """splitmix64 of the draw index -> [0,1), as examples/headless_demo.c."""
    with np.errstate(over="ignore"):
        x = k.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def test_c_host_is_built(built):
    """The plain C host of the ABI (examples/headless_demo.c) links against libfusionsim.so alone."""
    exe = os.path.join(ROOT, "fusion_sim_b200", "csrc", "headless_demo")
    assert os.path.exists(exe)
    out = subprocess.run(["ldd", exe], stdout=subprocess.PIPE, text=True).stdout
    assert "libfusionsim.so" in out and "python" not in out.lower() and "torch" not in out.lower()


@pytest.mark.gpu
def test_c_host_matches_the_python_mirror(built, tmp_path):
    """The same demo scene driven from C and from the Python mirror: identical canvas bytes."""
    import json
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import C1_SPEC, c1_sink_source
    exe = os.path.join(ROOT, "fusion_sim_b200", "csrc", "headless_demo")
    ppm = str(tmp_path / "canvas.ppm")
    res = json.loads(subprocess.run([exe, "6", ppm], stdout=subprocess.PIPE, check=True, text=True).stdout)
    n = 160000
    u = _uniform01(np.arange(6 * n, dtype=np.uint64)).reshape(n, 6)
    pos = 0.2 * (u[:, :3] - 0.5)
    pos[:, 2] += 1
    vel = 0.002 * (u[:, 3:] - 0.5)
    sink, source = c1_sink_source(400, 800)
    g = makeCylindricalParticlePusher(C1_SPEC)
    g.set({"position": pos, "velocity": vel, "sink_mask": sink, "source_pdf": source})
    g.addCurrentLoop(0.8, 2.0, -10000000)
    g.addCurrentLoop(0.8, 0.0, 10000000)
    g.precalc()
    for _ in range(6):
        g.step(); g.density()
    canvas = g.canvas
    h = 1469598103934665603
    for b in canvas.reshape(-1).tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert res["canvas_fnv1a"] == "%016x" % h and res["particles"] == n and res["launches"] > 0
    raw = open(ppm, "rb").read()
    assert raw.startswith(b"P6\n400 800\n255\n")
    assert np.array_equal(np.frombuffer(raw[len(b"P6\n400 800\n255\n"):], np.uint8).reshape(800, 400, 3), canvas[..., :3])


def test_node_addon_source_type_checks_and_matches_the_js_shim():
    """No Node.js in this image: the addon cannot be built or run, but its source is type-checked against a
    declarations-only stand-in for napi.h (tests/stubs/napi.h, written from the documented node-addon-api surface),
    every C entry point it calls is declared in include/fusionsim.h, and every member the JS shim calls on the native
    object is one the addon registers."""
    js_dir = os.path.join(ROOT, "fusion_sim_b200", "js")
    cc = os.path.join(js_dir, "fusionsim_napi.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "stubs"),
                        "-I" + os.path.join(ROOT, "include"), cc], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(cc).read()
    registered = set(re.findall(r'InstanceMethod\("(\w+)"', src))
    shim = open(os.path.join(js_dir, "empic_b200.js")).read()
    called = set(re.findall(r"\bsim\.(\w+)\(", shim))
    assert called and called <= registered, sorted(called - registered)
    header = open(os.path.join(ROOT, "include", "fusionsim.h")).read()
    declared = set(re.findall(r"\b(fsim_\w+)\s*\(", header))
    used = set(re.findall(r"\b(fsim_\w+)\s*\(", src))
    assert used and used <= declared, sorted(used - declared)
