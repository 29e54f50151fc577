"""Where the end-to-end leg of bench.py spends its time (one B200): python tools/e2e_phases.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from fusion_sim_b200 import makeCylindricalParticlePusher  # noqa: E402
from fusion_sim_b200.scenes import apply_scene  # noqa: E402

sc = bench.build_scene("c5", 0, 1)
pos_h, k1 = bench.pinned_copy(sc["position"])
vel_h, k2 = bench.pinned_copy(sc["velocity"])
sc["position"], sc["velocity"] = pos_h, vel_h
sim = makeCylindricalParticlePusher(dict(sc["spec"], precision="f64"))
apply_scene(sim, sc)
nr, nz = int(sc["spec"]["nr"]), int(sc["spec"]["nz"])
keep = [torch.empty((nz, nr, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
canv = [t.numpy() for t in keep]
for _ in range(3):
    sim.step(); sim.density()
sim.sync()
for rep in range(2):
    t = [time.perf_counter()]
    sim.set({"position": pos_h}); sim.sync(); t.append(time.perf_counter())
    sim.set({"velocity": vel_h}); sim.sync(); t.append(time.perf_counter())
    sim.step(); sim.density(); sim.sync(); t.append(time.perf_counter())
    for k in range(19):
        sim.step(); sim.density()
    sim.sync(); t.append(time.perf_counter())
    for k in range(20):
        sim.step(); sim.density(); sim.render_async(canv[k & 1])
    sim.sync(); t.append(time.perf_counter())
    sim.render(canv[0]); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("rep %d: set(position) %.1f ms (%.1f GB/s), set(velocity) %.1f ms, first frame %.1f ms, 19 frames %.1f ms, "
          "20 frames + async canvas %.1f ms, one synchronous canvas %.1f ms" % (
              rep, d[0], pos_h.nbytes / d[0] / 1e6, d[1], d[2], d[3], d[4], d[5]), flush=True)
