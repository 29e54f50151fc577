"""GPU parity: libfusionsim.so (through the C ABI / host driver) against the CPU oracle on the
same seeded inputs.  Bar: bit-exact -- the path is IEEE arithmetic in a fixed order on both
sides (-fmad=false / -ffp-contract=off), so positions, velocities, RNG state, alive flags, cell
indices, per-cell counts and all grid fields must be identical, in fp64 and in fp32 mode."""
import numpy as np
import pytest

from conftest import assert_same, small_scene

pytestmark = pytest.mark.gpu

PRECISIONS = ["f64", "f32"]


def make_pair(sc, precalc=True):
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene
    from oracle.oracle import OraclePusher
    g = makeCylindricalParticlePusher(sc["spec"])
    o = OraclePusher(sc["spec"], nthreads=4)
    apply_scene(g, sc, precalc)
    apply_scene(o, sc, precalc)
    return g, o


def compare_particles(g, o, what):
    gp, op = g.getPosition(), o.getPosition()
    assert_same(gp[:, 3], op[:, 3], what + " alive")
    assert_same(gp, op, what + " position")
    assert_same(g.getVelocity(), o.getVelocity(), what + " velocity")
    assert_same(g.getRand(), o.getRand(), what + " rand")


def compare_grid(g, o, names, what):
    for nm in names:
        assert_same(g.getField(nm), o.getField(nm), f"{what} {nm}")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_set_conversions(precision):
    sc = small_scene(precision=precision, with_E=True)
    g, o = make_pair(sc, precalc=False)
    compare_particles(g, o, "after set")
    compare_grid(g, o, ["E", "sink_mask", "inv_cdf", "entropy"], "after set")
    assert np.isnan(g.getField("inv_cdf")).sum() > 0  # the demo pdf yields NaN texels


@pytest.mark.parametrize("precision", PRECISIONS)
def test_static_fields_and_precalc(precision):
    sc = small_scene(precision=precision, with_E=True)
    sc["current_z"], sc["bz"], sc["btheta"] = 3.0e5, 0.02, -0.01
    g, o = make_pair(sc)
    compare_grid(g, o, ["B", "R1", "R2", "R3", "A"], "precalc")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_corrected_preA(precision):
    sc = small_scene(precision=precision, with_E=True)
    sc["spec"]["corrected_preA"] = True
    g, o = make_pair(sc)
    compare_grid(g, o, ["A"], "corrected preA")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_half_steps_bit_exact(precision):
    sc = small_scene(precision=precision, n=4099)  # not a multiple of the vector width
    g, o = make_pair(sc)
    for k in range(30):
        g.half_step(); o.half_step()
        if k % 3 == 0 or k > 25:
            compare_particles(g, o, f"half-step {k}")
            assert_same(g.getCells(), _cells(o), f"half-step {k} cells")
    g.sync()


def _cells(o):
    from oracle.numpy_ref import tex
    p = o.position
    r = np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1])
    return tex(r, o.nr) + o.nr * tex(p[:, 2], o.nz)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_absorb_respawn_and_E(precision):
    """Fast particles near the wall: absorption, inverse-cdf respawn (NaN texels and r = 0
    included), fresh random velocity on the next half-step; E != 0 exercises A."""
    sc = small_scene(precision=precision, n=8192, speed=0.2, with_E=True, blob=(0.6, 0.9))
    g, o = make_pair(sc)
    respawned = 0
    for k in range(24):
        g.half_step(); o.half_step()
        respawned += int((o.position[:, 3] == 0).sum())
        compare_particles(g, o, f"half-step {k}")
    assert respawned > 100
    assert np.isnan(o.position).any()  # NaN respawn texels were hit and handled alike


@pytest.mark.parametrize("speed", [0.004, 0.2])
def test_float_entropy_and_sink_update(speed):
    """A reference-style entropy table (uint32 / 0xFFFFFFFF stored as floats, empic.js:143-155) and a
    set({sink_mask}) between steps: absorption must follow the new mask at once."""
    from fusion_sim_b200.scenes import entropy_table
    sc = small_scene(precision="f64", n=8192, speed=speed, with_E=True, blob=(0.6, 0.9))
    sc["entropy"] = entropy_table(np.random.Generator(np.random.PCG64(3)))
    g, o = make_pair(sc)
    for k in range(8):
        g.half_step(); o.half_step()
    compare_particles(g, o, "half-step 7")
    rng = np.random.Generator(np.random.PCG64(4))
    mask = (rng.random((o.nr, o.nz)) > 0.02).astype(np.float64)  # scattered absorbing cells
    g.set({"sink_mask": mask}); o.set({"sink_mask": mask})
    respawned = 0
    for k in range(8):
        g.half_step(); o.half_step()
        respawned += int((o.position[:, 3] == 0).sum())
        compare_particles(g, o, f"new mask, half-step {k}")
    assert respawned > 50
    g.step(); o.step()
    compare_particles(g, o, "fused step()")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_density_bit_exact(precision):
    sc = small_scene(precision=precision, n=20000, speed=0.02, blob=(0.5, 0.8))
    g, o = make_pair(sc)
    for frame in range(4):
        g.step(); o.step()
        g.density(); o.density()
        assert_same(g.getField("cell_count"), o.getField("cell_count"), f"frame {frame} counts")
        compare_grid(g, o, ["cell_sums", "moments01", "moments01_norm", "moments01_avg"], f"frame {frame}")
        compare_particles(g, o, f"frame {frame}")  # sorting must not disturb identities
    assert_same(g.canvas, o.canvas, "canvas")
    assert int(g.getField("cell_count").sum()) > 0


@pytest.mark.parametrize("precision", PRECISIONS)
def test_crowded_cells(precision):
    """All particles in a few cells: exercises the block-per-cell (bitonic by id) deposit path."""
    sc = small_scene(precision=precision, n=6000, speed=0.0005, blob=(0.01, 0.01))
    g, o = make_pair(sc)
    for frame in range(2):
        g.step(); o.step()
        g.density(); o.density()
        assert g.getField("cell_count").max() > 64
        assert_same(g.getField("cell_count"), o.getField("cell_count"), "counts")
        compare_grid(g, o, ["cell_sums", "moments01_avg"], f"crowded frame {frame}")


def test_atomic_deposit_alternative_within_tolerance():
    """FSIM_FLAG_ATOMIC_DEPOSIT: same terms, arrival-order sums -> 1e-12 relative, counts exact."""
    sc = small_scene(n=20000, speed=0.02, blob=(0.5, 0.8))
    sc["spec"]["flags"] = 4
    g, o = make_pair(sc)
    g.step(); o.step()
    g.density(); o.density()
    assert_same(g.getField("cell_count"), o.getField("cell_count"), "counts")
    a, b = g.getField("cell_sums"), o.getField("cell_sums")
    scale = np.abs(b).max()
    assert np.nanmax(np.abs(a - b)) <= 1e-12 * scale
    m, n = g.getField("moments01_avg"), o.getField("moments01_avg")
    ok = ~np.isnan(n)
    assert np.abs(m[ok] - n[ok]).max() <= 1e-12 * np.abs(n[ok]).max()


def test_sort_is_invisible():
    sc = small_scene(n=5000, speed=0.05, blob=(0.5, 0.8))
    g, o = make_pair(sc)
    ids0 = g.getIds()  # storage order: set({position}) on a fresh handle already puts the storage into cell order
    assert_same(np.sort(ids0), np.arange(5000, dtype=np.uint64), "initial ids")
    for k in range(6):
        g.step(); o.step()
        g.sort()
        compare_particles(g, o, f"step {k} after sort")
    ids = g.getIds()
    assert sorted(ids.tolist()) == list(range(5000))
    assert not np.array_equal(ids, ids0)
    # set() after a sort addresses particles by id, not by storage slot
    g.set({"velocity": sc["velocity"]}); o.set({"velocity": sc["velocity"]})
    compare_particles(g, o, "set after sort")


def test_c1_demo_scene_frames():
    """Config C1 (fusionsim.js:72-148): 160 000 particles, 400x800 grid, frames of step+density."""
    from fusion_sim_b200.scenes import c1_scene
    sc = c1_scene(12345)
    sc["spec"]["keep_moments"] = True
    g, o = make_pair(sc)
    compare_grid(g, o, ["B", "R1", "R2", "R3", "A"], "C1 precalc")
    for frame in range(5):
        g.step(); o.step()
        g.density(); o.density()
    compare_particles(g, o, "C1 frame 5")
    assert_same(g.getField("cell_count"), o.getField("cell_count"), "C1 counts")
    compare_grid(g, o, ["moments01", "moments01_avg"], "C1 frame 5")
    assert_same(g.canvas, o.canvas, "C1 canvas")


def test_async_canvas_matches_sync():
    sc = small_scene(n=3000, speed=0.02, blob=(0.5, 0.8))
    g, o = make_pair(sc)
    imgs = [np.zeros((g.nz, g.nr, 4), np.uint8) for _ in range(3)]
    want = []
    for k in range(3):
        g.step(); o.step()
        g.density(); o.density()
        g.render_async(imgs[k])
        want.append(o.canvas)
    g.sync()
    for k in range(3):
        assert_same(imgs[k], want[k], f"async canvas {k}")


def test_ids_outside_the_block_are_refused():
    from fusion_sim_b200 import Error, makeCylindricalParticlePusher
    import ctypes as C
    sc = small_scene(n=256)
    g = makeCylindricalParticlePusher(sc["spec"])
    ids = np.arange(256, dtype=np.uint64) * 2  # not a permutation of [0, N)
    from fusion_sim_b200._lib import check, lib
    with pytest.raises(Error):
        check(lib().fsim_set_ids(g.handle, ids.ctypes.data_as(C.c_void_p)))


def test_errors_are_thrown():
    from fusion_sim_b200 import Error, makeCylindricalParticlePusher
    sc = small_scene(n=64)
    with pytest.raises(Error, match=r"\.nr <- Non-optional property is undefined!"):
        makeCylindricalParticlePusher({k: v for k, v in sc["spec"].items() if k != "nr"})
    g = makeCylindricalParticlePusher(sc["spec"])
    with pytest.raises(Error, match="beta_c"):
        g.addSpindleCuspPlasmaField(1.0, 0.5, 2.0)
    with pytest.raises(Error):
        g.set({"position": np.zeros((3, 3))})


def test_full_size_properties():
    """Config C3 shape (16.7 M particles, 2048^2 grid): size-independent invariants.
    Physical kinetic energy is conserved by the Boris rotation (E = 0) to rounding; RNG state
    stays in [0,1]; the deposit accounts for every in-range particle exactly once."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene, c1_sink_source, plasma_particles, scaled_loops, scaled_spec
    n, nr, nz = 1 << 24, 2048, 2048
    spec = scaled_spec(nr, nz, n)
    pos, vel = plasma_particles(spec, n, seed=3)
    sink, source = c1_sink_source(nr, nz)
    g = makeCylindricalParticlePusher(spec)
    apply_scene(g, dict(position=pos, velocity=vel, sink_mask=sink, source_pdf=source,
                        loops=scaled_loops(spec)))
    fac = np.array([1 / spec["radius"], 1 / spec["radius"], 1 / spec["height"]])
    e0 = ((g.getVelocity() / fac) ** 2).sum(1)
    for _ in range(3):
        g.step()
    g.density()
    g.sync()
    p = g.getPosition()
    stayed = p[:, 3] == 1
    e1 = ((g.getVelocity() / fac) ** 2).sum(1)
    rel = np.abs(e1[stayed] / e0[stayed] - 1)
    assert stayed.mean() > 0.99
    assert rel.max() < 1e-12, rel.max()  # tolerance: 6 rotations x a few ulp
    q = g.getRand()
    assert q.min() >= 0.0 and q.max() <= 1.0
    r = np.sqrt(p[:, 0] ** 2 + p[:, 1] ** 2)
    inside = (r * nr >= 0) & (r * nr < nr) & (p[:, 2] * nz >= 0) & (p[:, 2] * nz < nz)
    counts = g.getField("cell_count")
    assert int(counts.sum()) == int(inside.sum())
    cells = (r[inside] * nr).astype(np.int64) + nr * (p[inside, 2] * nz).astype(np.int64)
    assert_same(counts, np.bincount(cells, minlength=nr * nz).astype(np.uint32), "C3 per-cell counts")
    sums = g.getField("cell_sums")
    np.testing.assert_allclose(sums[:, 3].sum(), 0.001 * inside.sum(), rtol=1e-12)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_edge_shapes(precision):
    """Empty particle set, a single particle, a one-column grid and a non-square particle count."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene
    from oracle.oracle import OraclePusher
    # no particles at all: every call must still work
    sc = small_scene(precision=precision, n=16)
    spec0 = dict(sc["spec"], nparticles=0)
    spec0.pop("nparticles_total", None)
    g = makeCylindricalParticlePusher(spec0)
    g.set({"sink_mask": sc["sink_mask"], "source_pdf": sc["source_pdf"]})
    g.addCurrentLoop(0.8, 2.0, -1e7)
    g.precalc()
    g.step(); g.density(); g.sort(); g.sync()
    assert g.getPosition().shape == (0, 4) and int(g.getField("cell_count").sum()) == 0
    assert g.canvas.shape == (g.nz, g.nr, 4)
    # one particle; odd counts; tiny grids
    for n, nr, nz in ((1, 8, 8), (7, 1, 16), (33, 16, 1), (1000, 5, 7)):
        sc = small_scene(precision=precision, n=n, nr=nr, nz=nz, speed=0.05, blob=(0.7, 0.9))
        g, o = make_pair(sc)
        for _ in range(3):
            g.step(); o.step()
            g.density(); o.density()
        compare_particles(g, o, f"n={n} grid {nr}x{nz}")
        assert_same(g.getField("cell_count"), o.getField("cell_count"), "counts")
        compare_grid(g, o, ["moments01_avg"], f"n={n} grid {nr}x{nz}")
        assert_same(g.canvas, o.canvas, "canvas")


def test_long_run_c1_energy_and_population():
    """Config C1 for 200 frames (400 half-steps): bounded energy drift of the particles that were
    never respawned (Boris rotation, E = 0), RNG state in range, particle count conserved."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene, c1_scene
    sc = c1_scene(7)
    g = makeCylindricalParticlePusher(sc["spec"])
    apply_scene(g, sc)
    fac = np.array([1.0, 1.0, 0.5])
    e0 = ((g.getVelocity() / fac) ** 2).sum(1)
    never = np.ones(g.n, bool)
    for k in range(200):
        g.step(); g.density()
        if k % 20 == 19:
            never &= g.getPosition()[:, 3] == 1
    g.sync()
    e1 = ((g.getVelocity() / fac) ** 2).sum(1)
    # a respawned particle gets |v| <= 0.001*sqrt(3) again; the untouched ones keep their energy
    drift = np.abs(e1[never] / e0[never] - 1)
    assert never.sum() > 1000
    assert drift.max() < 1e-11, drift.max()  # 400 rotations x ~1e-16
    q = g.getRand()
    assert q.min() >= 0 and q.max() <= 1 and g.n == 160000
    ids = np.sort(g.getIds())
    assert_same(ids, np.arange(160000, dtype=np.uint64), "ids after 200 frames")


@pytest.mark.parametrize("precision", PRECISIONS)
def test_checkpoint_restore_resumes_bit_for_bit(precision):
    """checkpoint() on a running simulation (self-consistent fields included), restore() into a fresh
    handle with the same static tables: both continue identically."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene
    sc = small_scene(precision=precision, n=20000, speed=0.02, blob=(0.5, 0.8))
    a = makeCylindricalParticlePusher(sc["spec"])
    apply_scene(a, sc)
    v = {"macro_weight": 5e11, "sweeps": 6, "omega": 0.9}
    for _ in range(3):
        a.step(); a.density(); a.solveFields(v)
    ck = a.checkpoint()
    b = makeCylindricalParticlePusher(sc["spec"])
    b.set({k: sc[k] for k in ("sink_mask", "source_pdf", "entropy")})
    b.restore(ck)
    for sim in (a, b):
        for _ in range(3):
            sim.step(); sim.density(); sim.solveFields(v)
    for nm, fa, fb in (("position", a.getPosition(), b.getPosition()), ("velocity", a.getVelocity(), b.getVelocity()),
                       ("rand", a.getRand(), b.getRand())):
        assert_same(fa, fb, "resumed " + nm)
    for nm in ("moments01_avg", "phi", "E", "A", "cell_count"):
        assert_same(a.getField(nm), b.getField(nm), "resumed " + nm)
    assert_same(a.canvas, b.canvas, "resumed canvas")


def test_check_digest_and_fused_resort_and_post_stream_flags():
    """fsim_check_digest against NumPy on the accessor values; the re-sort fused into step()'s sweep (every
    sort_interval-th frame), the un-fused re-sort pass and the stencil / canvas draws on a second stream change
    no bit: runs with sort_interval 2, FSIM_FLAG_POST_STREAM and FSIM_FLAG_UNFUSED_SORT equal the default run."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene
    sc = small_scene(n=20000, speed=0.05, blob=(0.7, 0.9))
    runs = []
    for extra in ({}, {"sort_interval": 2}, {"flags": 8}, {"flags": 16, "sort_interval": 3}):
        g = makeCylindricalParticlePusher(dict(sc["spec"], **extra))
        apply_scene(g, sc)
        imgs = []
        for k in range(9):
            g.step(); g.density(); g.draw_canvas()
            if k % 4 == 0:
                imgs.append(g.canvas.copy())
        runs.append((g, imgs))
    g0 = runs[0][0]
    for g, imgs in runs[1:]:
        assert_same(g.getPosition(), g0.getPosition(), "position")
        assert_same(g.getRand(), g0.getRand(), "rand")
        assert_same(g.getField("moments01_avg"), g0.getField("moments01_avg"), "running average")
        assert_same(g.getField("cell_count"), g0.getField("cell_count"), "counts")
        for a, b in zip(imgs, runs[0][1]):
            assert_same(a, b, "canvas")
    assert not np.array_equal(runs[1][0].getIds(), g0.getIds())  # the storage orders DO differ
    d = g0.check_digest()
    ids = g0.getIds().astype(np.uint64)
    assert d["particles"] == 20000 and d["id_xor"] == int(np.bitwise_xor.reduce(ids)) and d["id_sum"] == int(ids.sum())
    cnt = g0.getField("cell_count")
    assert d["deposited"] == int(cnt.sum())
    np.testing.assert_allclose(d["sum_alpha"], g0.getField("cell_sums")[:, 3].sum(), rtol=1e-12)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_run_frames_graph_replay_changes_no_bit(precision):
    """fsim_run_frames(n) = n x (step, density, canvas draws); after one cycle launched one by one it captures a CUDA
    graph of 2 x sort_interval frames and replays it.  Replayed frames, the frames around the capture, a tail shorter
    than a cycle, frames after a setter dropped the graph, and frames after a hand-made step() shifted the cycle all
    equal the same number of frames launched one by one -- particles, running average, counts, canvas."""
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene
    sc = small_scene(n=20000, speed=0.05, blob=(0.7, 0.9))
    spec = dict(sc["spec"], precision=precision, sort_interval=3)  # cycle = 6 frames
    a, b = makeCylindricalParticlePusher(spec), makeCylindricalParticlePusher(spec)
    apply_scene(a, sc); apply_scene(b, sc)

    def by_hand(sim, n):
        for _ in range(n):
            sim.step(); sim.density(); sim.draw_canvas()

    def same(what):
        assert_same(a.getPosition(), b.getPosition(), what + " position")
        assert_same(a.getVelocity(), b.getVelocity(), what + " velocity")
        assert_same(a.getRand(), b.getRand(), what + " rand")
        # (the storage order inside a cell is decided by cursor atomics: it differs from run to run and changes no result)
        assert np.array_equal(np.sort(a.getIds()), np.sort(b.getIds())), what + " ids"
        assert_same(a.getField("moments01_avg"), b.getField("moments01_avg"), what + " running average")
        assert_same(a.getField("cell_count"), b.getField("cell_count"), what + " counts")
        assert_same(a.canvas, b.canvas, what + " canvas")

    by_hand(a, 47); b.run_frames(47)
    info = b.frame_graph_info()
    assert info["frames_per_cycle"] == 6 and info["replays"] >= 5 and info["launches_per_cycle"] > 6 * 8, info
    assert a.launch_count == b.launch_count  # replayed launches are counted
    same("47 frames")
    # a setter drops the graph; the next call launches a cycle one by one, captures again
    vel = a.getVelocity() * 0.5
    a.set({"velocity": vel * 2.998e8}); b.set({"velocity": vel * 2.998e8})
    by_hand(a, 31); b.run_frames(31)
    assert b.frame_graph_info()["replays"] > info["replays"]
    same("after set(velocity)")
    # frames by hand shift the cycle: the captured phase is never met again, the graph is dropped and re-captured later
    by_hand(a, 2); by_hand(b, 2)
    by_hand(a, 40); b.run_frames(40)
    same("after frames by hand")
    # reading does not drop it
    r0 = b.frame_graph_info()["replays"]
    by_hand(a, 12); b.run_frames(6); b.getPosition(); b.check_digest(); b.sync(); b.run_frames(6)
    same("after getters")
    assert b.frame_graph_info()["frames_per_cycle"] == 6 and b.frame_graph_info()["replays"] >= r0


def test_run_frames_on_the_demo_scene_and_refusals():
    import torch
    from fusion_sim_b200 import Error, makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene, c1_scene
    sc = c1_scene(7)
    a, b = makeCylindricalParticlePusher(sc["spec"]), makeCylindricalParticlePusher(sc["spec"])
    apply_scene(a, sc); apply_scene(b, sc)

    def by_hand(sim, n):
        for _ in range(n):
            sim.step(); sim.density(); sim.draw_canvas()

    by_hand(a, 48)
    b.run_frames(48)  # 16 one by one, capture, 2 replays: the handle is back at the captured phase
    assert b.frame_graph_info()["frames_per_cycle"] == 16 and b.frame_graph_info()["replays"] == 2
    assert_same(a.getPosition(), b.getPosition(), "position")
    assert_same(a.getField("moments01_avg"), b.getField("moments01_avg"), "running average")
    assert_same(a.canvas, b.canvas, "canvas")
    # two canvas read-backs in flight on the copy stream (one per device image) when the next replay wants to draw
    # into those images: the replay waits for them
    want = a.canvas.copy()
    pinned = [torch.empty((800, 400, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    b.render_async(pinned[0].numpy()); b.render_async(pinned[1].numpy())
    b.run_frames(16); by_hand(a, 16)
    b.sync()
    assert b.frame_graph_info()["replays"] == 3
    assert_same(pinned[0].numpy(), want, "first read-back"); assert_same(pinned[1].numpy(), want, "second read-back")
    assert_same(a.getPosition(), b.getPosition(), "position after the read-backs")
    assert_same(a.canvas, b.canvas, "canvas after the read-backs")
    with pytest.raises(Error, match="negative"):
        b.run_frames(-1)
    # per-launch timing needs its events: frames go one by one while it is on
    b.timing(True)
    b.run_frames(40)
    assert b.frame_graph_info()["frames_per_cycle"] == 0
    b.timing(False)
