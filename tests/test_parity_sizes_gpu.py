"""GPU parity AT THE SIZES THAT ARE BENCHMARKED (BASELINE.json configs; bench.py WORKLOADS):
libfusionsim.so through the host driver against the CPU oracle, bit for bit.

  C1  the reference's demo scene, 1000 frames of step()+density() (fusionsim.js:170-178), compared
      with the oracle every 50 frames (both sides are IEEE arithmetic in a fixed order, so the
      trajectories are IDENTICAL for the whole run -- chaos does not matter when no bit differs),
      plus a bound on the energy drift of the particles that were never respawned;
  C2  1 Mi particles, 512 x 512 cells: every frame of 10, everything (particles, cells, counts,
      per-cell sums, running average, canvas);
  C3  16 Mi particles, 2048 x 2048 cells: every frame of 3, everything;
  C5  64 Mi particles, 8192 x 2048 cells (the bench default): fields and Boris records of the whole
      grid; then, on the frame after a plain binning and on a frame that re-sorts the storage, the
      particle state of a 1 Mi-id random subset (particles are independent: the oracle steps the
      subset from the GPU's pre-step state) and the per-cell counts, sums and running average on
      64-row windows (the oracle is fed the particles that deposit within 5 rows of the window).

The synthetic plasma of the C2-C5 tests reaches right up to the walls (the bench's stays 2 % away),
so absorption, inverse-cdf respawn (NaN texels included) and clipped sprites happen at these sizes.
"""
import os

import numpy as np
import pytest

from conftest import assert_same
from window_oracle import oracle_window

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 4


def sized_scene(n, nr, nz, seed, precision="f64"):
    from fusion_sim_b200.scenes import c1_sink_source, entropy_table, plasma_particles, scaled_loops, scaled_spec
    spec = scaled_spec(nr, nz, n, precision=precision)
    pos, vel = plasma_particles(spec, n, seed, z_lo=0.0005, z_hi=0.9995, r_lo=0.0005, r_hi=0.9995)
    sink, source = c1_sink_source(nr, nz)
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    rand = rng.random((n, 4))
    entropy = entropy_table(rng)
    return dict(spec=spec, position=pos, velocity=vel, sink_mask=sink, source_pdf=source, rand=rand,
                entropy=entropy, loops=scaled_loops(spec))


def make_pair(sc):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene
    from oracle.oracle import OraclePusher
    g = makeCylindricalParticlePusher(sc["spec"])
    o = OraclePusher(sc["spec"], nthreads=THREADS)
    apply_scene(g, sc)
    apply_scene(o, sc)
    return g, o


def oracle_cells(o):
    from oracle.numpy_ref import tex
    p = o.position
    r = np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1])
    return tex(r, o.nr) + o.nr * tex(p[:, 2], o.nz)


def compare_everything(g, o, what, canvas=False):
    gp, op = g.getPosition(), o.getPosition()
    assert_same(gp[:, 3], op[:, 3], what + " alive")
    assert_same(gp, op, what + " position")
    assert_same(g.getVelocity(), o.getVelocity(), what + " velocity")
    assert_same(g.getRand(), o.getRand(), what + " rand")
    assert_same(g.getCells(), oracle_cells(o), what + " cells")
    assert_same(g.getField("cell_count"), o.getField("cell_count"), what + " per-cell counts")
    assert_same(g.getField("cell_sums"), o.getField("cell_sums"), what + " per-cell sums")
    assert_same(g.getField("moments01_avg"), o.getField("moments01_avg"), what + " running average")
    if canvas:
        assert_same(g.canvas, o.canvas, what + " canvas")
    return int((op[:, 3] == 0).sum())


def run_full(sc, frames, what):
    g, o = make_pair(sc)
    for nm in ("B", "R1", "R2", "R3", "A"):
        assert_same(g.getField(nm), o.getField(nm), f"{what} {nm}")
    respawned = 0
    for f in range(frames):
        g.step(); o.step()
        g.density(); o.density()
        respawned += compare_everything(g, o, f"{what} frame {f}", canvas=(f == frames - 1))
    g.sync()
    assert respawned > 0, "the test plasma must reach the walls"
    assert int(g.getField("cell_count").sum()) > 0.9 * g.n


def test_c2_every_frame_of_ten():
    run_full(sized_scene(1 << 20, 512, 512, seed=11), 10, "C2")


def test_c2_fp32_every_frame_of_ten():
    """The reference's own storage precision (RGBA32F) at C2 size."""
    run_full(sized_scene(1 << 20, 512, 512, seed=12, precision="f32"), 10, "C2 fp32")


def test_c3_every_frame_of_three():
    run_full(sized_scene(1 << 24, 2048, 2048, seed=13), 3, "C3")


def test_c1_thousand_frames():
    """BASELINE config 1: the demo scene (fusionsim.js:72-148) for 1000 frames of the page's loop
    (fusionsim.js:170-178)."""
    from fusion_sim_b200.scenes import c1_scene
    sc = c1_scene(12345)
    g, o = make_pair(sc)
    fac = np.array([1.0, 1.0, 0.5])  # factor_r, factor_r, factor_z of the demo (radius 1, height 2)
    e0 = ((o.getVelocity() / fac) ** 2).sum(1)
    never = np.ones(o.n, bool)
    respawns = 0
    for f in range(1000):
        g.step()
        for _ in range(2):  # out.step = two half-steps (empic.js:1436-1469); a respawn shows in the alive flag for ONE of them
            o.half_step()
            alive = o.position[:, 3] == 1
            respawns += int((~alive).sum())
            never &= alive
        g.density(); o.density()
        if f % 50 == 49:
            last = f == 999
            gp, op = g.getPosition(), o.getPosition()
            assert_same(gp, op, f"C1 frame {f} position")
            assert_same(g.getVelocity(), o.getVelocity(), f"C1 frame {f} velocity")
            assert_same(g.getRand(), o.getRand(), f"C1 frame {f} rand")
            assert_same(g.getField("cell_count"), o.getField("cell_count"), f"C1 frame {f} counts")
            assert_same(g.getField("moments01_avg"), o.getField("moments01_avg"), f"C1 frame {f} running average")
            if last:
                assert_same(g.canvas, o.canvas, "C1 canvas after 1000 frames")
    g.sync()
    assert respawns > 1000  # the blob reached the wall (after ~1500 half-steps)
    # Boris rotation with E = 0 conserves |v|: 2000 rotations x ~1e-16 each
    e1 = ((g.getVelocity() / fac) ** 2).sum(1)
    drift = np.abs(e1[never] / e0[never] - 1)
    assert never.sum() > 1000
    assert drift.max() < 1e-11, drift.max()
    assert_same(np.sort(g.getIds()), np.arange(160000, dtype=np.uint64), "ids after 1000 frames")


# ---------------------------------------------------------------------------------------------------
# C5: the bench default
# ---------------------------------------------------------------------------------------------------
def test_c5_bench_default_shape():
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene
    from oracle.oracle import OraclePusher
    n, nr, nz, M = 1 << 26, 8192, 2048, 1 << 20
    sc = sized_scene(n, nr, nz, seed=17)
    g = makeCylindricalParticlePusher(sc["spec"])
    apply_scene(g, sc)
    # the oracle carries the whole grid but only a subset of the particles
    rng = np.random.Generator(np.random.PCG64(5))
    sub = np.sort(rng.choice(n, size=M, replace=False))
    o = OraclePusher(dict(sc["spec"], nparticles=1, nparticles_total=M), nthreads=THREADS)
    o.set({k: sc[k] for k in ("sink_mask", "source_pdf", "entropy")})
    for (r, z, I) in sc["loops"]:
        o.addCurrentLoop(r, z, I)
    o.precalc()
    del sc
    for nm in ("B", "R1", "R2", "R3", "A", "sink_mask", "inv_cdf"):
        assert_same(g.getField(nm), o.getField(nm), f"C5 {nm} (whole 8192 x 2048 grid)")

    windows = [(0, 64), (nz // 2 - 32, nz // 2 + 32), (nz - 64, nz)]  # both edges and the middle
    respawned = deposited = 0

    def checked_frame(what):
        nonlocal respawned, deposited
        gp0, gv0, gr0 = g.getPosition(), g.getVelocity(), g.getRand()
        avg0 = g.getField("moments01_avg").reshape(nz, nr, 4)
        prev = [avg0[w0:w1].reshape(-1, 4).copy() for (w0, w1) in windows]
        del avg0
        o.position[:] = gp0[sub]
        o.velocity[:, :3] = gv0[sub]
        o.rand[:] = gr0[sub]
        del gp0, gv0, gr0
        g.step(); o.step()
        g.density()
        gp, gv = g.getPosition(), g.getVelocity()
        assert_same(gp[sub], o.getPosition(), what + " position of the subset")
        assert_same(gv[sub], o.getVelocity(), what + " velocity of the subset")
        assert_same(g.getRand()[sub], o.getRand(), what + " rand of the subset")
        assert_same(g.getCells()[sub], oracle_cells(o), what + " cells of the subset")
        respawned += int((o.position[:, 3] == 0).sum())
        cnt = g.getField("cell_count").reshape(nz, nr)
        sums = g.getField("cell_sums").reshape(nz, nr, 4)
        avg = g.getField("moments01_avg").reshape(nz, nr, 4)
        for (w0, w1), pv in zip(windows, prev):
            oc, oS, oavg, nsel = oracle_window(gp, gv, pv, nr, nz, w0, w1, THREADS)
            assert nsel > 1000
            assert_same(cnt[w0:w1].reshape(-1), oc, f"{what} counts of rows {w0}..{w1}")
            assert_same(sums[w0:w1].reshape(-1, 4), oS, f"{what} per-cell sums of rows {w0}..{w1}")
            assert_same(avg[w0:w1].reshape(-1, 4), oavg, f"{what} running average of rows {w0}..{w1}")
            deposited += int(oc.sum())
        # size-independent property on the WHOLE grid: every in-range particle is counted exactly once
        r = np.sqrt(gp[:, 0] ** 2 + gp[:, 1] ** 2)
        inside = (r * nr >= 0) & (r * nr < nr) & (gp[:, 2] * nz >= 0) & (gp[:, 2] * nz < nz)
        assert int(cnt.sum(dtype=np.int64)) == int(inside.sum())

    # set({position}) makes the first step() sort the storage; density() re-sorts it on every 8th frame
    g.step(); g.density()          # frame 1
    checked_frame("C5 frame 2 (index binning)")
    for _ in range(5):             # frames 3..7
        g.step(); g.density()
    checked_frame("C5 frame 8 (re-sorts the storage)")
    checked_frame("C5 frame 9 (first frame on the re-sorted storage)")
    g.sync()
    assert respawned > 0 and deposited > 100000
