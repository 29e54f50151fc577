"""Frame time vs physical re-sort interval on the bench scene (one B200), and A/B over the engine flags
(8 = stencil and canvas draws on a second stream, 16 = re-sort as a pass of its own, 64 = second stream without priority).
    python tools/tune_sort_interval.py [workload] [precision] [intervals, comma separated] [flags]
The re-sort is fused into the sweep (push2_resort), so a sort costs the difference between push2_resort and
push2; what an interval buys is a more coherent cell-table gather (push2), longer runs of equal keys
(index_scatter) and nearer gathers (cellsum).  One JSON line per interval."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fusion_sim_b200 import makeCylindricalParticlePusher  # noqa: E402
from fusion_sim_b200.scenes import apply_scene  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c5"
precision = sys.argv[2] if len(sys.argv) > 2 else "f64"
intervals = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3, 4, 6, 8, 12, 16]
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sc = bench.build_scene(workload, 0, 1)
names = ("push2", "push2_resort", "scan", "index_scatter", "cellsum", "cellsum_warp", "conv", "render")
for interval in intervals:
    sim = makeCylindricalParticlePusher(dict(sc["spec"], sort_interval=interval, precision=precision, flags=flags))
    apply_scene(sim, sc)
    for _ in range(max(4, interval + 1)):
        sim.step(); sim.density(); sim.draw_canvas()
    sim.sync()
    K = max(24, 3 * interval)
    sim.mark(0)
    for _ in range(K):
        sim.step(); sim.density(); sim.draw_canvas()
    sim.mark(1)
    ms = sim.elapsed_ms(0, 1) / K
    sim.timing(True); sim.timing_reset()
    for _ in range(K):
        sim.step(); sim.density(); sim.draw_canvas()
    row = {"interval": interval, "flags": flags, "precision": precision, "frame_ms": round(ms, 4)}
    for nm in names:
        t, c = sim.timing_get(nm)
        if c:
            row[nm] = {"per_launch": round(t / c, 4), "per_frame": round(t / K, 4)}
    sim.timing(False)
    print(json.dumps(row), flush=True)
    sim.destroy()
