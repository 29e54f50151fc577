"""Generates tests/golden/oracle_small_{f64,f32}.npz from the CPU oracle (oracle/).

These are NOT outputs of the reference (it is WebGL and cannot run headless, SURVEY.md
section 8c); they pin the oracle itself against silent drift.  Regenerate only when the
restatement is deliberately changed:   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def run(precision):
    from conftest import small_scene
    from fusion_sim_b200.scenes import apply_scene
    from oracle.oracle import OraclePusher
    sc = small_scene(seed=11, nr=24, nz=40, n=256, precision=precision, speed=0.3, with_E=True, blob=(0.5, 0.9))
    # a deterministic entropy table that needs no 32 MB fixture
    k = np.arange(1024 * 1024 * 4, dtype=np.float64)
    sc["entropy"] = np.mod(k * 0.6180339887498949, 1.0).reshape(-1, 4)
    o = OraclePusher(sc["spec"])
    apply_scene(o, sc)
    for _ in range(5):
        o.step()
        o.density()
    return dict(position=o.getPosition(), velocity=o.getVelocity(), rand=o.getRand(),
                cell_count=o.cell_count.copy(), cell_sums=o.cell_sums.astype(np.float64),
                moments01_avg=o.moments01_avg.astype(np.float64), R1=o.getField("R1"), A=o.getField("A"),
                B=o.getField("B"), canvas=o.canvas)


if __name__ == "__main__":
    for p in ("f64", "f32"):
        np.savez_compressed(os.path.join(HERE, f"oracle_small_{p}.npz"), **run(p))
        print("wrote", p)
