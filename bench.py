#!/usr/bin/env python
"""bench.py -- particle-pushes/s of the full PIC frame (step() + density()) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c5|c3|c2|c1] [--precision f64|f32]

One "step" = one frame of the reference's loop (fusionsim.js:172-174): simulation.step() (two
leap-frog half-steps, empic.js:1436-1469) + simulation.density() (empic.js:1471-1505: deposit,
normalise, running average AND the two canvas draws) over the whole synthetic plasma; 1 push = one
half-step of one particle, so a frame is 2*N pushes.  Default workload:
BASELINE.json configs[4], the weak-scaling shape the metric is quoted on -- 64 Mi particles and an
8192 x 2048 slab of cells per GPU.  `--workload c4` is BASELINE configs[3]: 256 Mi particles on
8192 x 8192 cells, STRONG scaling over 2/4/8 GPUs.

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events on the engine's stream)
with the state resident in HBM; `e2e` is the same metric through the public host API with host
buffers: set(position, velocity) from pinned host memory, K frames, and a canvas read-back per
frame, all inside the timed region.  `check` holds run invariants reduced over all ranks (no particle
lost or duplicated; every in-range particle deposited once) and, at N > 1, a reduced scene on which
the slab run must equal a single-GPU run bit for bit.  `--impl reference` times the CPU restatement
of the reference's shader arithmetic (oracle/, OpenMP, all host threads) on the SAME per-GPU
configuration -- the reference itself is WebGL and cannot run headless.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_pushes_per_s_full_pic_step"
UNIT = "pushes/s"

WORKLOADS = {
    # name: (particles per GPU, nr, nz per GPU, description)
    "c5": (1 << 26, 8192, 2048, "C5 weak scaling: 64Mi particles + 8192x2048-cell slab per GPU"),
    "c4": (1 << 28, 8192, 8192, "C4 strong scaling: 256Mi particles, 8192x8192 grid, slab-decomposed"),  # TOTALS
    "c3": (1 << 24, 2048, 2048, "C3: 16Mi particles, 2048x2048 grid"),
    "c2": (1 << 20, 512, 512, "C2: 1Mi particles, 512x512 grid"),
    "c1": (160000, 400, 800, "C1: default demo scene, 160000 particles, 400x800 grid"),
}
DEFAULT_STEPS = {"c5": 20, "c4": 20, "c3": 200, "c2": 2000, "c1": 3000}  # timed region >= 0.1 s: the clock sampler sees it


def per_gpu_shape(workload, world):
    """(particles per GPU, nr, grid rows per GPU): C4 is strong scaling (totals divided), the rest weak."""
    n, nr, nz, _ = WORKLOADS[workload]
    if workload == "c4":
        return n // world, nr, nz // world
    return n, nr, nz


def xor_upto(m: int) -> int:
    """XOR of 0..m."""
    return [m, 1, m + 1, 0][m % 4] if m >= 0 else 0


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


_SAMPLER_SRC = r"""
import sys, time
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
    print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
    while True:
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        print("s", repr(time.time()), sm, int(r), flush=True)
        time.sleep(0.003)
except Exception as e:
    print("err", type(e).__name__, flush=True)
"""


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML) -- from a process of its own, so that
    neither the GIL nor the launch loop of the bench can starve it; samples are time-stamped and the ones
    inside [begin(), end()] are kept."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        import subprocess
        import tempfile
        self.t0 = self.t1 = None
        self.sm_max = None
        self.first = []
        try:
            self.log = tempfile.NamedTemporaryFile("w+", suffix=".clocks", delete=False)  # a file, not a pipe: never blocks the sampler
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index)], stdout=self.log,
                                         stderr=subprocess.DEVNULL)
            for _ in range(2000):  # "max <MHz>" once NVML is up: the timed region may start
                with open(self.log.name) as f:
                    line = f.readline()
                if line.endswith("\n"):
                    self.first = line.split()
                    break
                if self.proc.poll() is not None:
                    break
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def result(self):
        sm, reasons = [], set()
        if self.proc is None or not self.first or self.first[0] != "max":
            try:
                if self.proc is not None:
                    self.proc.kill()
                os.unlink(self.log.name)
            except Exception:
                pass
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"], "samples": 0}
        self.sm_max = int(self.first[1])
        time.sleep(0.01)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        with open(self.log.name) as f:
            out = f.read()
        os.unlink(self.log.name)
        for line in out.splitlines():
            f = line.split()
            if len(f) == 4 and f[0] == "s" and self.t0 <= float(f[1]) <= self.t1:
                sm.append(int(f[2]))
                for bit, nm in self.REASONS.items():
                    if int(f[3]) & bit:
                        reasons.add(nm)
        sm.sort()
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.sm_max,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def push_algorithmic_bytes(n, ncell, precision, halves=2):
    """Bytes one sweep of the step kernel must move (per GPU): SURVEY.md section 8d with this
    build's cell record (8 reals per cell instead of the survey's 12, sink mask 1 bit per cell):
    particle state read + written once (10 reals + 1 flag byte each way); the cell table, the
    entropy table (4 reals x 1024^2) and the inverse-cdf table (2 reals x 512^2) once -- but never
    more than the sweep can TOUCH (one texel / record per particle and half-step): the demo scene's
    160 000 particles touch a third of the entropy table and 4 % of the cells."""
    rs = 8 if precision == "f64" else 4
    touches = float(halves) * float(n)
    cells = min(float(ncell), touches)
    return ((2 * (10 * rs + 1)) * float(n) + 8 * rs * cells + float(ncell) / 8
            + rs * (4 * min(1024.0 * 1024, touches) + 2 * min(512.0 * 512, touches)))


def build_scene(workload, rank, world, seed=2026):
    """Synthetic plasma of the named shape; with world > 1 the grid is nz*world rows and this
    rank owns rows [rank*nz, (rank+1)*nz) with the particles inside them."""
    from fusion_sim_b200.scenes import (c1_scene, c1_sink_source, plasma_particles, scaled_loops,
                                        scaled_spec)
    n, nr, nz_local = per_gpu_shape(workload, world)
    if workload == "c1":
        assert world == 1, "C1 is a single-GPU scene"
        return c1_scene(seed)
    nz = nz_local * world
    spec = scaled_spec(nr, nz, n * world)
    # The plasma keeps the SAME margin in grid rows from the two end walls whatever the number of slabs (2 % of one
    # slab = 41 rows at C5, as on one GPU), so every rank sees the particle density of the single-GPU case to within
    # 2 %.  (Round 1 kept 2 % of the WHOLE height free: at 8 slabs the two end ranks then held their 64 Mi particles
    # in 84 % of their rows, 19 % denser than the middle ranks -- their per-cell pass ran 0.08 ms slower and every
    # frame waited for them.)
    margin = 0.02 / world
    lo = max(margin, rank / world)
    hi = min(1.0 - margin, (rank + 1) / world)
    pos, vel = plasma_particles(spec, n, seed + rank, z_lo=lo, z_hi=hi)
    sink, source = c1_sink_source(nr, nz)
    return dict(spec=spec, position=pos, velocity=vel, sink_mask=sink, source_pdf=source,
                loops=scaled_loops(spec), n_local=n, nz_local=nz_local)


def pinned_copy(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    out = t.numpy()
    out[...] = a
    return out, t


def make_config(workload, world, precision, n_total, nr, nz, field_sweeps=0, decomposition="slab"):
    n_local = n_total // world
    return {"workload": WORKLOADS[workload][3], "particles_total": n_total, "grid": [nr, nz],
            "precision": precision,
            "frame": "step()+density() = 2 half-steps + deposit + normalise + running average + the two canvas draws "
                     "(empic.js:1436-1505)" + (" + solveFields(%d sweeps) [EXTENSION]" % field_sweeps if field_sweeps else ""),
            "l2": "inputs larger than L2 (particle state %.1f GB per GPU)" % (n_local * (81 if precision == "f64" else 41) / 1e9),
            "parallelism": ("replicated%d" if decomposition == "replicated" and world > 1 else "slab%d") % world}


def cpu_baseline(workload, precision, steps, target_s=12.0, threads=None, scene=None, one_thread=True):
    """CPU restatement of the reference's shader arithmetic (oracle/, kind "port"), timed on the host
    cores on the WHOLE per-GPU configuration of the workload (C5: 64 Mi particles on 8192 x 2048 cells;
    C4: the 8-GPU share), a bounded number of frames."""
    from fusion_sim_b200.scenes import apply_scene, c1_scene, c1_sink_source, entropy_table, plasma_particles, scaled_loops, scaled_spec
    from oracle.oracle import OraclePusher
    threads = threads or (os.cpu_count() or 1)
    n, nr, nz = per_gpu_shape(workload, 8 if workload == "c4" else 1)
    if workload == "c1":
        sc = c1_scene(2026)
        sc["spec"]["precision"] = precision
        sample = "whole C1 scene (160000 particles, 400x800 grid)"
    else:
        spec = scaled_spec(nr, nz, n, precision=precision)
        if scene is not None and len(scene["position"]) == n and scene["spec"]["nz"] == nz:
            pos, vel = scene["position"], scene["velocity"]  # the very arrays the GPU arm uploads
        else:
            pos, vel = plasma_particles(spec, n, 2026, z_lo=0.02, z_hi=0.98)
        sink, source = c1_sink_source(nr, nz)
        sc = dict(spec=spec, position=pos, velocity=vel, sink_mask=sink, source_pdf=source, loops=scaled_loops(spec))
        sample = f"whole per-GPU workload: {nr}x{nz} cells, {n} particles"
        if workload == "c4":
            sample += " (the 8-GPU share of C4)"
    o = OraclePusher(sc["spec"], nthreads=threads)
    rng = np.random.Generator(np.random.PCG64(5))
    sc["rand"] = rng.random((o.n, 4))
    sc["entropy"] = entropy_table(rng)
    apply_scene(o, sc)
    del sc

    def frame():
        o.step()
        o.density(timing_mt=True)
        o.canvas  # noqa: B018 -- the two canvas draws are part of the reference's density()

    frame()  # warm-up (page faults, OpenMP pool)
    t0 = time.perf_counter()
    frame()
    t1 = time.perf_counter() - t0
    k = steps if steps else max(3, min(2000, int(target_s / max(t1, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(k):
        frame()
    dt = time.perf_counter() - t0
    # the same configuration on ONE host thread (SURVEY.md section 8d asks for both): one or a few frames,
    # only when that stays within ~25 s
    one = None
    est1 = t1 * threads * 0.7
    if one_thread and threads > 1 and est1 < 25.0:
        o.nthreads = 1
        k1 = max(1, min(k, int(6.0 / max(est1, 1e-6))))
        t0 = time.perf_counter()
        for _ in range(k1):
            frame()
        one = 2.0 * o.n * k1 / (time.perf_counter() - t0)
        o.nthreads = threads
    return {"value": 2.0 * o.n * k / dt, "unit": UNIT, "cores": threads, "kind": "port", "value_1_thread": one,
            "sample": sample + f", {k} frames in {dt:.1f} s", "sample_fraction": 1.0,
            "what": "CPU restatement of the reference's shader arithmetic (C, OpenMP); the reference "
                    "itself is WebGL and has no CPU path"}, dt / k * 1e3, k


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # the driver passes the steps of the GPU arm; a CPU frame of the full configuration costs about a
    # second, so the count is capped to keep the run within a few minutes
    want = min(args.steps, 12) if args.steps_given else 0
    cb, ms, k = cpu_baseline(args.workload, args.precision, want, threads=threads, one_thread=False)
    world = max(1, args.gpus)
    n, nr, nz = per_gpu_shape(args.workload, world)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": k, "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None, "dtype": args.precision,
        "data": "synthetic",
        "config": make_config(args.workload, world, args.precision, n * world, nr, nz * world),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "one CPU host whatever N is: the sample is the per-GPU configuration (at N = 1 that IS the "
                "GPU arm's configuration); a rate, so comparable per GPU",
    }
    print(json.dumps(line), flush=True)


def reduced_slab_check(rank, world, local, precision):
    """At N > 1: a reduced scene (2^17 particles and a 256 x 64 slab per rank, plasma right up to the
    walls so that particles are absorbed and respawn into other slabs) run twice on this GPU's rank --
    as one slab of the N-rank run and, whole, as a single-GPU run.  After 4 frames this rank's particles
    (by global id) and its rows of per-cell counts and running average must be equal bit for bit."""
    import ctypes as C
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200._lib import check, lib
    from fusion_sim_b200.dist import SlabPusher, slab_bounds
    from fusion_sim_b200.scenes import apply_scene, c1_sink_source, entropy_table, plasma_particles, scaled_loops, scaled_spec
    nr, nz, n = 256, 64 * world, (1 << 17) * world
    spec = scaled_spec(nr, nz, n, precision=precision, device=local)
    pos, vel = plasma_particles(spec, n, 4242, z_lo=0.0005, z_hi=0.9995, r_lo=0.0005, r_hi=0.9995)
    vel *= 8.0  # several rows per frame: drift across the slab boundaries, not only respawns
    sink, source = c1_sink_source(nr, nz)
    rng = np.random.Generator(np.random.PCG64(77))
    sc = dict(spec=spec, position=pos, velocity=vel, sink_mask=sink, source_pdf=source, rand=rng.random((n, 4)),
              entropy=entropy_table(rng), loops=scaled_loops(spec))
    single = makeCylindricalParticlePusher(spec)
    apply_scene(single, sc)
    b = slab_bounds(nz, world)
    row = np.clip(np.floor(pos[:, 2] / spec["height"] * nz).astype(np.int64), 0, nz - 1)
    sel = np.nonzero((row >= b[rank]) & (row < b[rank + 1]))[0]
    loc = dict(sc)
    for k in ("position", "velocity", "rand"):
        loc[k] = sc[k][sel]
    slab = SlabPusher(dict(spec), loc, rank, world, halo_rows=16, cap_neighbour=16384, cap_far=8192)
    gid = np.ascontiguousarray(sel.astype(np.uint64))
    check(lib().fsim_set_ids(slab.sim.handle, gid.ctypes.data_as(C.c_void_p)))
    for _ in range(4):
        single.step(); single.density()
        slab.step(); slab.density()
    slab.sync(); single.sync()
    ids = slab.sim.getIds().astype(np.int64)
    same = lambda x, y: bool(((x == y) | (np.isnan(x) & np.isnan(y))).all())
    ok = same(single.getPosition()[ids], slab.sim.getPosition()) and same(single.getVelocity()[ids], slab.sim.getVelocity()) \
        and same(single.getRand()[ids], slab.sim.getRand())
    row0 = max(0, b[rank] - 16)
    lo, hi = (b[rank] - row0) * nr, (b[rank + 1] - row0) * nr
    glo, ghi = b[rank] * nr, b[rank + 1] * nr
    ok = ok and same(single.getField("moments01_avg")[glo:ghi], slab.sim.getField("moments01_avg")[lo:hi])
    ok = ok and bool((single.getField("cell_count")[glo:ghi] == slab.sim.getField("cell_count")[lo:hi]).all())
    moved = slab.migrated
    slab.sim.destroy(); single.destroy()
    return ok, moved


def run_ours(args):
    import torch
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if world > 1:
        import torch.distributed as dist
        # stdout carries ONE JSON line: everything else a library writes to file descriptor 1 (NCCL prints its
        # version banner there) is sent to stderr; the line itself goes to the saved descriptor
        sys.stdout.flush()
        args.real_stdout = os.dup(1)
        os.dup2(2, 1)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fusion_sim_b200 import makeCylindricalParticlePusher
    from fusion_sim_b200.scenes import apply_scene

    def all_ranks(values, op="sum"):
        """Reduce a list of Python ints over the ranks (exact: gathered, combined on the host)."""
        if world == 1:
            return list(values)
        import torch.distributed as dist
        t = torch.tensor([int(v) & 0x7fffffffffffffff for v in values] + [int(v) >> 63 for v in values],
                         dtype=torch.int64, device="cuda")
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        k = len(values)
        res = [0] * k
        for o in out:
            o = o.tolist()
            for i in range(k):
                v = o[i] | (o[k + i] << 63)
                res[i] = (res[i] ^ v) if op == "xor" else ((res[i] + v) & 0xffffffffffffffff)
        return res

    reduced = None
    if world > 1 and args.decomposition == "slab" and not args.no_reduced_check:
        ok, moved = reduced_slab_check(rank, world, local, args.precision)
        okall, movedall = all_ranks([0 if ok else 1, moved])
        reduced = {"slab_equals_single_gpu": okall == 0, "records_migrated": movedall,
                   "what": "reduced scene (2^17 particles + 256x64 cells per rank, walls reached), 4 frames: every rank's "
                           "particles by global id, per-cell counts and running average bit-equal to a single-GPU run"}

    sc = build_scene(args.workload, rank, world)
    spec = dict(sc["spec"])
    spec.update(precision=args.precision, device=local)
    n_local = sc.get("n_local", len(sc["position"]))
    pos_h, _keep1 = pinned_copy(sc["position"])
    vel_h, _keep2 = pinned_copy(sc["velocity"])
    sc["position"], sc["velocity"] = pos_h, vel_h
    if world > 1 and args.decomposition == "replicated":
        from fusion_sim_b200.dist import ReplicatedPusher
        sim = ReplicatedPusher(spec, sc, rank, world)  # measured alternative: no migration, all-reduce of the sums
    elif world > 1:
        from fusion_sim_b200.dist import SlabPusher
        sim = SlabPusher(spec, sc, rank, world, exchange=args.exchange)
    else:
        sim = makeCylindricalParticlePusher(spec)
        apply_scene(sim, sc)
    nr, nz = int(spec["nr"]), int(spec["nz"])
    ncell_local = sim.ncell_local
    slab_mode = world > 1 and args.decomposition == "slab"
    own_rows = nz // world if slab_mode else nz  # replicated: every rank renders the whole canvas
    # two pinned host images for the e2e leg: frame k's read-back overlaps frame k+1.  A slab rank reads
    # back its own rows only (its share of the canvas).
    _keep3 = [torch.empty((own_rows, nr, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    canvases = [t.numpy() for t in _keep3]

    solve = {"macro_weight": 1.0e6, "sweeps": args.field_sweeps, "omega": 1.0} if args.field_sweeps > 0 else None

    def frame():
        sim.step()
        sim.density()
        if solve:  # EXTENSION: the self-consistent frame; off by default (the reference has none)
            sim.solveFields(solve)
        sim.draw_canvas()  # the two canvas draws of the reference's density() (empic.js:1497-1504), device-resident

    def barrier():
        sim.sync()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # launch-bound scenes (C1, the reference's demo scene: ~14 launches of 3-30 us per frame; C2): the K frames go
    # through fsim_run_frames, which replays a captured CUDA graph of 16 frames -- same kernels, same bits
    use_graph = world == 1 and not solve and (args.graph == "on" or (args.graph == "auto" and args.workload in ("c1", "c2")))

    def run(nframes):
        if use_graph:
            sim.run_frames(nframes)
        else:
            for _ in range(nframes):
                frame()

    # ---- device-resident timing: W warm-up frames, then exactly K frames ----
    sampler = ClockSampler(local)  # returns once its process has NVML up
    if use_graph:
        args.warmup = max(args.warmup, 48)  # one cycle launched one by one, then the capture: both outside the timed region
    run(args.warmup)
    barrier()
    sampler.begin()
    l0 = sim.launch_count
    t_host0 = time.perf_counter()
    sim.mark(0)
    run(args.steps)
    sim.mark(1)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3
    graph_info = sim.frame_graph_info() if use_graph else None
    ms_total = sim.elapsed_ms(0, 1)
    sampler.end()
    barrier()
    launches = sim.launch_count - l0
    ms_total = reduce_max(ms_total)
    clocks = sampler.result()
    n_total = n_local * world
    value = 2.0 * n_total * args.steps / (ms_total * 1e-3)

    # ---- run invariants, reduced over the ranks: nothing lost, nothing duplicated, everything deposited ----
    dg = sim.check_digest()
    n_local_now = dg["particles"]
    tot = all_ranks([dg["particles"], dg["id_sum"], dg["deposited"]])
    (xr,) = all_ranks([dg["id_xor"]], op="xor")
    want_sum = (n_total * (n_total - 1) // 2) & 0xffffffffffffffff
    check_line = {"particles_total": tot[0], "id_xor": xr, "id_sum": tot[1], "deposited": tot[2],
                  "sum_alpha": reduce_max(dg["sum_alpha"]) if world == 1 else None,
                  "expected": {"particles_total": n_total, "id_xor": xor_upto(n_total - 1), "id_sum": want_sum},
                  "ok": tot[0] == n_total and xr == xor_upto(n_total - 1) and tot[1] == want_sum and tot[2] <= n_total
                        and tot[2] > 0.98 * n_total}
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dg["sum_alpha"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        check_line["sum_alpha"] = float(t.item())
    if args.decomposition == "replicated" and world > 1:  # every rank deposits every particle there
        check_line["ok"] = tot[0] == n_total and xr == xor_upto(n_total - 1) and tot[1] == want_sum
    if reduced is not None:
        check_line["reduced_scene"] = reduced
        check_line["ok"] = check_line["ok"] and reduced["slab_equals_single_gpu"]

    # ---- per-kernel device times over K more frames (CUDA events around every launch), and the exchanges ----
    sim.timing(True)
    sim.timing_reset()
    if slab_mode:
        sim.comm_timing = True
        sim.comm_ms()
    for _ in range(args.steps):
        frame()
    kern = {}
    for nm in ("push", "push2", "push2_resort", "prepass", "scan", "index_scatter", "permute", "cellsum", "cellsum_warp", "cellsum_heavy", "conv",
               "render", "migrate_pack", "migrate_unpack", "charge_source", "relax4", "relax2", "relax1", "efield", "precalc"):
        ms, cnt = sim.timing_get(nm)
        if cnt:
            kern[nm] = {"ms_per_launch": ms / cnt, "launches_per_step": cnt / args.steps,
                        "ms_per_step": ms / args.steps}
    sim.timing(False)
    comm = None
    if slab_mode:
        c = sim.comm_ms()
        sim.comm_timing = False
        comm = {k: reduce_max(v / args.steps) for k, v in c.items()}
        comm["migrate_exchange_min_over_ranks"] = -reduce_max(-c["migrate_exchange"] / args.steps)  # the slowest rank waits least
        comm["migrate_exchange_alone"] = reduce_max(sim.exchange_alone_ms())  # the collective by itself, all ranks entering together
        comm["what"] = ("max over ranks, ms per frame, CUDA events on the frame's stream: migrate_exchange = the all-to-all of "
                        "the fixed-size record regions (includes waiting for the slowest peer: compare migrate_exchange_alone, "
                        "the same collective timed with all ranks entering together, and the minimum over ranks); halo_exchange_exposed = what "
                        "is left of the 5-row halo exchange after the interior stencil ran under it; host_wait = host time "
                        "blocked in synchronisations inside the frame (none with the fixed-region exchange)")
        comm["exchange"] = args.exchange
        comm["host_enqueue_ms_per_step"] = reduce_max(host_enqueue_ms / args.steps)
    per_rank = None
    if world > 1:  # who sets the pace: every rank's own kernel time per frame, its sweep, its SM clock under load
        import torch.distributed as dist
        mine = [sum(v["ms_per_step"] for v in kern.values()), kern.get("push2", {}).get("ms_per_launch", 0.0),
                kern.get("cellsum", {}).get("ms_per_launch", 0.0), float(clocks["sm_mhz"] or 0),
                1.0 if "sw_power_cap" in clocks["reasons"] else 0.0, float(n_local_now)]
        t = torch.tensor(mine, dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        cols = list(zip(*[a.tolist() for a in allr]))
        per_rank = {"kernels_ms_per_step": [round(v, 4) for v in cols[0]], "push2_ms_per_launch": [round(v, 4) for v in cols[1]],
                    "cellsum_ms_per_launch": [round(v, 4) for v in cols[2]], "sm_mhz_median": cols[3],
                    "sw_power_cap_seen": [bool(v) for v in cols[4]], "particles": [int(v) for v in cols[5]]}
    peak, peak_kind = measured_peak()
    # dominant kernel: the fused step sweep (both half-steps of step() in one pass over HBM,
    # push.cu NH=2).  Its algorithmic bytes are those of ONE sweep: particle state read + written
    # once, tables once (SURVEY.md section 8d per-half-step figure).
    pk = "push2" if "push2" in kern else "push"
    halves = 2 if pk == "push2" else 1
    push_ms = kern[pk]["ms_per_launch"]
    alg = push_algorithmic_bytes(n_local, ncell_local, args.precision, halves)
    achieved = alg / (push_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "push_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{args.workload}_{args.precision}")
        except Exception:
            traffic = None
    roofline = {"kernel": "push_kernel<NH=%d> (rand + gather/Boris + push/absorb/respawn, %d half-step(s) per sweep"
                          " + deposit prepass)" % (halves, halves),
                "half_steps_per_launch": halves,
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_kind": peak_kind + " HBM copy bandwidth",
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": push_ms, "traffic": traffic,
                "achieved_per_push_formula": halves * alg / (push_ms * 1e-3) / 1e9,
                "frac_of_nominal_8000": achieved / 8000.0, "kernels_ms_per_step": kern}
    # whole frame against the roofline: algorithmic bytes of the sweep + the deposit chain (SURVEY 8d:
    # 128 bytes per cell in fp64, the particle part being fused into the sweep)
    rs = 8 if args.precision == "f64" else 4
    frame_alg = alg + 16.0 * rs * ncell_local
    roofline["frame"] = {"algorithmic_bytes": frame_alg, "achieved": frame_alg / (ms_total / args.steps * 1e-3) / 1e9,
                         "frac": frame_alg / (ms_total / args.steps * 1e-3) / 1e9 / peak}

    # ---- end to end through the public host API with host buffers ----
    # set(position, velocity) from pinned host arrays, then K frames each followed by the read-back
    # of the canvas (this rank's rows) into pinned host memory; all inside the timed region.
    from fusion_sim_b200._lib import check, lib
    base = sim.sim if world > 1 else sim
    h2d = 2 * pos_h.nbytes
    d2h = 4 * nr * own_rows
    def e2e_pass(frames):
        check(lib().fsim_set_particle_count(base.handle, n_local))
        base.set({"position": pos_h, "velocity": vel_h})
        for k in range(frames):
            sim.step()
            sim.density()
            if solve:
                sim.solveFields(solve)
            if slab_mode:
                sim.render_rows_async(canvases[k & 1])
            else:
                sim.render_async(canvases[k & 1])

    e2e_pass(2)  # untimed warm-up of exactly this path: first use of the copy stream, canvas buffers, pinned pages
    barrier()
    t0 = time.perf_counter()
    sim.mark(2)
    e2e_pass(args.steps)
    sim.mark(3)
    ms_e2e = sim.elapsed_ms(2, 3)
    sim.sync()  # also waits for the last canvas copy
    ms_e2e = reduce_max(max(ms_e2e, (time.perf_counter() - t0) * 1e3))
    e2e = {"value": 2.0 * n_total * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h, "ms_total": ms_e2e,
           "what": "per rank: set(position,velocity) from pinned host arrays once + per frame step(), density(), "
                   "canvas read-back (own rows) to pinned host memory on a copy stream, overlapping the next frame; upload "
                   "amortised over the K frames; wall clock to the last byte on the host"}

    # ---- EXTENSION row N4, reported beside the headline (never inside it unless --field-sweeps is given):
    # device time of one solveFields() of 8 sweeps and of its kernels on this workload's grid
    ext = None
    if world == 1 and not solve and not args.no_extension_probe:
        probe = {"macro_weight": 1.0e6, "sweeps": 8, "omega": 1.0}
        sim.solveFields(probe)  # allocation + warm-up
        sim.timing(True)
        sim.timing_reset()
        sim.mark(4)
        for _ in range(5):
            sim.solveFields(probe)
        sim.mark(5)
        ext = {"what": "solveFields(8 weighted-Jacobi sweeps) = charge source + 2 x relax4 (TMA-staged, 4 sweeps per "
                       "launch) + E = -grad(phi) + precalc; EXTENSION, no reference counterpart",
               "ms_per_solve": sim.elapsed_ms(4, 5) / 5, "kernels_ms_per_launch": {}}
        for nm, nbytes in (("charge_source", 2 * rs), ("relax4", 3 * rs), ("efield", 4 * rs), ("precalc", 14 * rs)):
            ms, cnt = sim.timing_get(nm)
            if cnt:
                ext["kernels_ms_per_launch"][nm] = {"ms": ms / cnt, "algorithmic_GBps": nbytes * ncell_local / (ms / cnt * 1e-3) / 1e9}
        sim.timing(False)

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if hasattr(sim, "destroy"):
            sim.destroy()  # the CPU leg needs the host memory bandwidth to itself
        cb, _, _ = cpu_baseline(args.workload, args.precision, 0, scene=sc)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic",
            "config": make_config(args.workload, world, args.precision, n_total, nr, nz, args.field_sweeps, args.decomposition),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cb, "check": check_line, "comm_ms_per_step": comm, "per_rank": per_rank,
            "push_only_pushes_per_s": halves * n_local * world / (push_ms * 1e-3),
            "extension_field_solve": ext,
            "frame_graph": graph_info,
        }
        text = json.dumps(line) + "\n"
        if getattr(args, "real_stdout", None) is not None:
            os.write(args.real_stdout, text.encode())
        else:
            sys.stdout.write(text)
            sys.stdout.flush()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if not check_line["ok"]:
        sys.stderr.write("bench.py: run invariants violated: %s\n" % json.dumps(check_line))
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--exchange", default="fixed", choices=["fixed", "exact"],
                    help="slab runs: fixed-capacity regions with device-side counts (no host round trip) or the exact "
                         "all-to-all-v with counts read back to the host")
    ap.add_argument("--no-reduced-check", action="store_true", help="skip the reduced-scene slab == single-GPU check")
    ap.add_argument("--no-extension-probe", action="store_true")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="launch the timed frames through fsim_run_frames (a replayed CUDA graph of 16 frames); auto = "
                         "the launch-bound workloads c1 and c2 on one GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decomposition", default="slab", choices=["slab", "replicated"],
                    help="multi-GPU: slab decomposition (default) or the measured alternative (replicated tables + all-reduce)")
    ap.add_argument("--field-sweeps", type=int, default=0,
                    help="EXTENSION: add solveFields(N sweeps) to every frame (self-consistent fields); 0 = the reference's frame")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = DEFAULT_STEPS[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
