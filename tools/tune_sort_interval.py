"""Frame time vs physical re-sort interval on the bench scene (one B200)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fusion_sim_b200 import makeCylindricalParticlePusher
from fusion_sim_b200.scenes import apply_scene
sc = bench.build_scene(sys.argv[1] if len(sys.argv) > 1 else "c5", 0, 1)
for interval in (4, 8, 16, 32, 64):
    sim = makeCylindricalParticlePusher(dict(sc["spec"], sort_interval=interval))
    apply_scene(sim, sc)
    for _ in range(4):
        sim.step(); sim.density()
    sim.sync(); sim.timing(True); sim.timing_reset()
    K = 2 * max(interval, 8)
    sim.mark(0)
    for _ in range(K):
        sim.step(); sim.density()
    sim.mark(1)
    ms = sim.elapsed_ms(0, 1) / K
    row = {"interval": interval, "frame_ms": round(ms, 3)}
    for nm in ("push2", "permute", "cellsum", "index_scatter"):
        t, c = sim.timing_get(nm)
        if c:
            row[nm] = round(t / c, 3)
    print(json.dumps(row), flush=True)
    sim.destroy()
