"""Host-side logic (no GPU): spec validation messages of utilities.validate_object, the demo
scene of fusionsim.js:72-148, and the bench's JSON contract on the CPU reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_validate_object_messages():
    from fusion_sim_b200 import Error
    from fusion_sim_b200.pusher import validate_object
    from fusion_sim_b200.scenes import C1_SPEC
    validate_object(C1_SPEC, {k: "number" for k in C1_SPEC})
    with pytest.raises(Error) as e:
        validate_object({k: v for k, v in C1_SPEC.items() if k != "dt"}, {k: "number" for k in C1_SPEC})
    assert str(e.value) == ".dt <- Non-optional property is undefined!"  # utilities.js:18,124
    with pytest.raises(Error) as e:
        validate_object(dict(C1_SPEC, nr="400"), {k: "number" for k in C1_SPEC})
    assert str(e.value) == ".nr <- Property does not match any given possible types!"  # utilities.js:68


def test_c1_scene_restates_fusionsim_js():
    from fusion_sim_b200.scenes import c1_scene
    sc = c1_scene(12345)
    assert sc["spec"]["nr"] == 400 and sc["spec"]["nz"] == 800 and sc["spec"]["nparticles"] == 400
    p, v = sc["position"], sc["velocity"]
    assert p.shape == (160000, 3) and v.shape == (160000, 3)
    assert np.abs(p[:, :2]).max() <= 0.1 and np.abs(p[:, 2] - 1).max() <= 0.1  # fusionsim.js:126
    assert np.abs(v).max() <= 0.001  # :127
    sink, src = sc["sink_mask"], sc["source_pdf"]
    assert sink[399].sum() == 0 and sink[0, 0] == 1 and sink[0, 799] == 1  # :105-112, corners stay 1
    assert sink[1:399, 0].sum() == 0 and sink[1:399, 799].sum() == 0 and sink[0, 1:799].all()
    assert src.sum() == 50 * 100 and src[:50, 350:450].all()  # :116-122
    assert sc["loops"] == [(0.8, 2.0, -1e7), (0.8, 0.0, 1e7)]  # :137-138
    again = c1_scene(12345)
    assert np.array_equal(again["position"], p) and np.array_equal(again["rand"], sc["rand"])


def test_scaled_scene_keeps_cell_size_and_occupancy():
    from fusion_sim_b200.scenes import plasma_particles, scaled_spec
    spec = scaled_spec(512, 256, 100000)
    assert spec["radius"] / spec["nr"] == 1 / 400 and spec["height"] / spec["nz"] == 1 / 400
    assert spec["nparticles_total"] == 100000
    pos, vel = plasma_particles(spec, 100000, 1, z_lo=0.25, z_hi=0.5)
    r = np.hypot(pos[:, 0], pos[:, 1]) / spec["radius"]
    z = pos[:, 2] / spec["height"]
    assert 0.02 <= r.min() and r.max() <= 0.98 and 0.25 <= z.min() and z.max() <= 0.5
    assert np.abs(vel).max() <= 0.001


def test_bench_reference_arm_contract():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                                   "--workload", "c2", "--steps", "2", "--warmup", "1"], cwd=ROOT, timeout=300)
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle_pushes_per_s_full_pic_step"
    assert line["unit"] == "pushes/s" and line["higher_is_better"] is True and line["steps"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["value"] > 1e5


def test_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    # SURVEY.md section 8d worked example, C3 fp64: 2.72 GB particle state + cell table + tables;
    # this build's cell record is 64 B + 1 sink BIT instead of the survey's 97 B => 3.02 GB instead of 3.17 GB
    b = bench.push_algorithmic_bytes(1 << 24, 2048 * 2048, "f64")
    assert abs(b / 1e9 - 3.02) < 0.02
    # tables count at most what the sweep can touch: 1000 particles x 2 half-steps x one texel each
    assert bench.push_algorithmic_bytes(1000, 0, "f32") == 82 * 1000 + 4 * (4 * 2000 + 2 * 2000)
    # the demo scene (C1): 160 000 particles touch 320 000 of 1 Mi entropy texels and at most 320 000 cells
    c1 = bench.push_algorithmic_bytes(160000, 320000, "f64")
    assert c1 == 162 * 160000 + 64 * 320000 + 320000 / 8 + 8 * (4 * 320000 + 2 * 262144)
    assert bench.xor_upto(6) == 0 ^ 1 ^ 2 ^ 3 ^ 4 ^ 5 ^ 6 and bench.xor_upto(-1) == 0
