"""The C-ABI boundary (no GPU): the shared library loads, exports every symbol include/fusionsim.h
declares, the header is valid C whose struct layout matches the ctypes binding, and without a
CUDA device the product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess
import sys

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
