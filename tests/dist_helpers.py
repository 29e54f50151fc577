"""Helpers for the multi-rank tests: an oracle-backed slab rank (CPU, numpy) that runs through the
SAME exchange functions as the CUDA SlabPusher (fusion_sim_b200/dist.py), and the worker
functions spawned by the tests.  TEST CODE: may use oracle/."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

REC = np.dtype([("state", np.float64, (10,)), ("id", np.uint32), ("alive", np.uint32)])


def scene_for_dist(n=6000, nr=32, nz=64):
    from conftest import small_scene
    sc = small_scene(seed=21, nr=nr, nz=nz, n=n, speed=0.3, blob=(0.6, 0.95))
    sc["spec"].pop("nparticles_total", None)
    sc["spec"]["nparticles"] = 1
    sc["spec"]["nparticles_total"] = n
    return sc


def split_scene(sc, rank, world):
    """Particles of `sc` whose initial row belongs to `rank`; returns (local scene, global ids)."""
    from fusion_sim_b200.dist import slab_bounds
    from oracle.numpy_ref import tex
    spec = sc["spec"]
    z = sc["position"][:, 2] / spec["height"]
    row = tex(z, spec["nz"])
    b = slab_bounds(spec["nz"], world)
    sel = np.nonzero((row >= b[rank]) & (row < b[rank + 1]))[0]
    loc = dict(sc)
    for k in ("position", "velocity", "rand"):
        loc[k] = sc[k][sel]
    return loc, sel.astype(np.uint32)


class OracleSlab:
    """Numpy/oracle emulation of one slab rank, driven through dist.exchange_records/halo."""

    def __init__(self, sc, rank, world, exchange="fixed"):
        from fusion_sim_b200.dist import slab_bounds
        self.exchange = exchange
        from fusion_sim_b200.scenes import apply_scene
        from oracle.oracle import OraclePusher
        self.rank, self.world = rank, world
        loc, ids = split_scene(sc, rank, world)
        spec = dict(sc["spec"], nparticles_total=len(ids))
        self.o = OraclePusher(spec)
        apply_scene(self.o, loc)
        self.ids = ids
        self.bounds = slab_bounds(self.o.nz, world)

    def _resize(self, pos, vel, rnd, ids):
        o = self.o
        o.n = len(ids)
        o.position, o.velocity, o.rand = pos, vel, rnd
        o.last_cell = np.zeros(o.n, np.int64)
        o.deposit_cell = np.zeros(o.n, np.int64)
        self.ids = ids

    def step(self):
        self.o.step()

    def migrate(self):
        from fusion_sim_b200.dist import (HEADER_BYTES, exchange_records, exchange_regions, region_bytes,
                                          region_capacities)
        from oracle.numpy_ref import tex
        o = self.o
        row = tex(o.position[:, 2].astype(np.float64), o.nz)
        dest = np.searchsorted(np.asarray(self.bounds[1:]), row, side="right")
        dest = np.minimum(dest, self.world - 1)
        leave = dest != self.rank
        order = np.argsort(dest[leave], kind="stable")
        idx = np.nonzero(leave)[0][order]
        rec = np.zeros(len(idx), REC)
        rec["state"][:, 0:3] = o.position[idx, 0:3]
        rec["state"][:, 3:6] = o.velocity[idx, 0:3]
        rec["state"][:, 6:10] = o.rand[idx]
        rec["id"] = self.ids[idx]
        rec["alive"] = o.position[idx, 3].astype(np.uint32)
        counts = [int((dest[leave] == k).sum()) for k in range(self.world)]
        if self.exchange == "exact":  # all-to-all-v sized by exchanged counts
            send = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy())
            recv, nrecv = exchange_records(send, counts, REC.itemsize)
            got = recv.numpy().view(REC)
        else:  # the wire format of the CUDA SlabPusher: fixed-capacity regions, counts in the headers
            caps = region_capacities(self.rank, self.world, 4096, 1024)
            nbytes = [region_bytes(c, REC.itemsize) for c in caps]
            send = np.zeros(sum(nbytes), np.uint8)
            off = first = 0
            for k in range(self.world):
                assert counts[k] <= caps[k], "test scene overflows the exchange region"
                send[off:off + 4].view(np.uint32)[0] = counts[k]
                body = rec[first:first + counts[k]].view(np.uint8).reshape(-1)
                send[off + HEADER_BYTES:off + HEADER_BYTES + body.size] = body
                off += nbytes[k]
                first += counts[k]
            recv = torch.zeros(sum(nbytes), dtype=torch.uint8)
            exchange_regions(torch.from_numpy(send), nbytes, recv, nbytes)
            buf, parts, off = recv.numpy(), [], 0
            for k in range(self.world):
                c = int(buf[off:off + 4].view(np.uint32)[0])
                parts.append(buf[off + HEADER_BYTES:off + HEADER_BYTES + c * REC.itemsize].view(REC))
                off += nbytes[k]
            got = np.concatenate(parts)
            nrecv = len(got)
        keep = ~leave
        pos = np.concatenate([o.position[keep], np.c_[got["state"][:, 0:3], got["alive"].astype(np.float64)]])
        vel = np.concatenate([o.velocity[keep], np.c_[got["state"][:, 3:6], np.ones(nrecv)]])
        rnd = np.concatenate([o.rand[keep], got["state"][:, 6:10]])
        ids = np.concatenate([self.ids[keep], got["id"]])
        self._resize(np.ascontiguousarray(pos), np.ascontiguousarray(vel), np.ascontiguousarray(rnd), ids)
        self.sent = getattr(self, "sent", 0) + sum(counts)
        self.sent_far = getattr(self, "sent_far", 0) + sum(c for k, c in enumerate(counts) if abs(k - self.rank) > 1)
        return sum(counts)

    def density(self):
        from fusion_sim_b200.dist import HALO_DEPOSIT, exchange_halo
        import ctypes as C
        from oracle import oracle as orc
        o = self.o
        self.migrate()
        order = np.argsort(self.ids, kind="stable")  # GL primitive order = particle id order
        self._resize(np.ascontiguousarray(o.position[order]), np.ascontiguousarray(o.velocity[order]),
                     np.ascontiguousarray(o.rand[order]), self.ids[order])
        f = lambda name: getattr(orc.lib(), name + "_f64")
        f("orc_cell_sums")(C.c_int64(o.n), orc._p(o.position), orc._p(o.velocity), C.c_int64(o.nr),
                           C.c_int64(o.nz), orc._p(o.cell_sums), orc._p(o.cell_count), orc._p(o.deposit_cell))
        S = o.cell_sums.reshape(o.nz, o.nr, 4)
        lo, hi, H = self.bounds[self.rank], self.bounds[self.rank + 1], HALO_DEPOSIT
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        send_lo, send_hi = t(S[lo:lo + H]), t(S[hi - H:hi])
        recv_lo = t(np.zeros((H, o.nr, 4))) if self.rank > 0 else torch.empty(0)
        recv_hi = t(np.zeros((H, o.nr, 4))) if self.rank < self.world - 1 else torch.empty(0)
        exchange_halo(send_lo, send_hi, recv_lo, recv_hi, self.rank, self.world)
        if self.rank > 0:
            S[lo - H:lo] = recv_lo.numpy()
        if self.rank < self.world - 1:
            S[hi:hi + H] = recv_hi.numpy()
        f("orc_convolve")(C.c_int64(o.nr), C.c_int64(o.nz), orc._p(o.cell_sums), orc._p(o.shape),
                          orc._p(o.moments01), C.c_int(1))
        f("orc_normalize_ema")(C.c_int64(o.nr), C.c_int64(o.nz), orc._p(o.moments01), orc._p(o.moments01_norm),
                               orc._p(o.moments01_avg), C.c_int(1))

    # EXTENSION: the staged field solve, through the same driver as the CUDA SlabPusher
    def solveFields(self, value):
        from fusion_sim_b200.dist import solve_fields_slab
        solve_fields_slab(self, value)

    def fs_stage(self, stage, value, sweeps):
        import ctypes as C
        from oracle import oracle as orc
        o = self.o
        sp = o.spec
        dr, dz = sp["radius"] / o.nr, sp["height"] / o.nz
        f = lambda name: getattr(orc.lib(), name + "_f64")
        if not hasattr(o, "phi"):
            o.phi = np.zeros(o.ncell, o.dt)
            o.rho_src = np.zeros(o.ncell, o.dt)
        if stage == 0:
            dens = o.moments01_avg if value.get("source", "avg") == "avg" else o.moments01_norm
            scale = sp["particle_charge"] * float(value["macro_weight"]) / (orc.FSIM_PI * sp["radius"] * dr * dz * orc.FSIM_EPS0)
            f("orc_charge_source")(C.c_int64(o.ncell), orc._p(dens), C.c_double(scale), orc._p(o.rho_src))
        elif stage == 1:
            coef = np.empty((o.nr, 4), np.float64)
            orc.lib().orc_relax_coeffs(C.c_int64(o.nr), C.c_double(dr), C.c_double(dz), orc._p(coef))
            tmp = np.empty_like(o.phi)
            f("orc_relax")(C.c_int64(o.nr), C.c_int64(o.nz), orc._p(o.phi), orc._p(tmp), orc._p(o.rho_src), orc._p(coef),
                           C.c_double(float(value.get("omega", 1.0))), C.c_int(int(sweeps)), C.c_int(1))
        elif stage == 2:
            f("orc_efield")(C.c_int64(o.nr), C.c_int64(o.nz), orc._p(o.phi), C.c_double(1 / (2 * dr)),
                            C.c_double(1 / (2 * dz)), orc._p(o.E))
        else:
            o.precalc()

    def fs_exchange(self, name, nrows):
        """This emulation keeps the WHOLE grid on every rank but only trusts its owned rows and the
        rows it received: rows further away hold stale values, like the far halo rows of a GPU slab."""
        from fusion_sim_b200.dist import exchange_halo
        o = self.o
        a = getattr(o, name).reshape(o.nz, o.nr, -1)
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        h = 8 if nrows is None else int(nrows)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x))
        up, down = self.rank < self.world - 1, self.rank > 0
        recv_lo = t(np.zeros_like(a[lo - h:lo])) if down else torch.empty(0)
        recv_hi = t(np.zeros_like(a[hi:hi + h])) if up else torch.empty(0)
        exchange_halo(t(a[lo:lo + h]) if down else torch.empty(0), t(a[hi - h:hi]) if up else torch.empty(0),
                      recv_lo, recv_hi, self.rank, self.world)
        if down:
            a[lo - h:lo] = recv_lo.numpy()
        if up:
            a[hi:hi + h] = recv_hi.numpy()

    def owned(self, a):
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        return a.reshape(self.o.nz, self.o.nr, -1)[lo:hi].reshape((hi - lo) * self.o.nr, -1)


def _gather(rank, world, parts):
    out = [None] * world if rank == 0 else None
    dist.gather_object(parts, out, dst=0)
    return out


SOLVE = {"macro_weight": 5e11, "sweeps": 7, "omega": 0.9}  # per-frame field solve of the self-consistent tests


def cpu_worker(rank, world, port, frames, result_path, solve=False, exchange="fixed"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = scene_for_dist()
        s = OracleSlab(sc, rank, world, exchange)
        for _ in range(frames):
            s.step()
            s.density()
            if solve:
                s.solveFields(SOLVE)
        parts = dict(ids=s.ids, pos=s.o.position, vel=s.o.velocity, rnd=s.o.rand,
                     avg=s.owned(s.o.moments01_avg), cnt=s.owned(s.o.cell_count), moved=s.sent, far=s.sent_far)
        if solve:
            parts.update(phi=s.owned(s.o.phi).reshape(-1), E=s.owned(s.o.E)[:, :3])
        out = _gather(rank, world, parts)
        if rank == 0:
            np.savez(result_path, **assemble(out))
    finally:
        dist.destroy_process_group()


def assemble(out):
    ids = np.concatenate([o["ids"] for o in out])
    order = np.argsort(ids, kind="stable")
    extra = {k: np.concatenate([o[k] for o in out]) for k in ("phi", "E") if k in out[0]}
    return dict(extra, ids=ids[order], pos=np.concatenate([o["pos"] for o in out])[order],
                vel=np.concatenate([o["vel"] for o in out])[order][:, :3],
                rnd=np.concatenate([o["rnd"] for o in out])[order],
                avg=np.concatenate([o["avg"] for o in out]), cnt=np.concatenate([o["cnt"] for o in out]).reshape(-1),
                moved=sum(o["moved"] for o in out), far=sum(o.get("far", 0) for o in out))


def gpu_worker(rank, world, port, frames, result_path, solve=False, exchange="fixed"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fusion_sim_b200.dist import SlabPusher
        sc = scene_for_dist()
        loc, ids = split_scene(sc, rank, world)
        spec = dict(sc["spec"], device=rank)
        s = SlabPusher(spec, loc, rank, world, halo_rows=8, exchange=exchange)
        import ctypes as C
        gid = np.ascontiguousarray(ids.astype(np.uint64))
        s._lib.check(s._lib.lib().fsim_set_ids(s.sim.handle, gid.ctypes.data_as(C.c_void_p)))
        for _ in range(frames):
            s.step()
            s.density()
            if solve:
                s.solveFields(SOLVE)
        s.sync()
        pos = s.sim.getPosition()
        lo = (s.bounds[rank] - max(0, s.bounds[rank] - 8)) * s.nr
        hi = lo + (s.bounds[rank + 1] - s.bounds[rank]) * s.nr
        parts = dict(ids=s.sim.getIds().astype(np.uint32), pos=pos, vel=np.c_[s.sim.getVelocity(), np.ones(len(pos))],
                     rnd=s.sim.getRand(), avg=s.sim.getField("moments01_avg")[lo:hi],
                     cnt=s.sim.getField("cell_count")[lo:hi], moved=s.migrated)
        if solve:
            parts.update(phi=s.sim.getField("phi")[lo:hi], E=s.sim.getField("E")[lo:hi])
        out = _gather(rank, world, parts)
        if rank == 0:
            np.savez(result_path, **assemble(out))
    finally:
        dist.destroy_process_group()


def single_oracle(frames, solve=False):
    from fusion_sim_b200.scenes import apply_scene
    from oracle.oracle import OraclePusher
    sc = scene_for_dist()
    o = OraclePusher(sc["spec"])
    apply_scene(o, sc)
    for _ in range(frames):
        o.step()
        o.density()
        if solve:
            o.solveFields(SOLVE)
    return o


def gpu_worker_replicated(rank, world, port, frames, result_path):
    """Index-sharded particles + replicated tables + all-reduce of the per-cell sums."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fusion_sim_b200.dist import ReplicatedPusher
        sc = scene_for_dist()
        n = len(sc["position"])
        per = n // world  # contiguous id blocks: rank r holds ids [r per, (r+1) per) -- nothing depends on WHERE they are
        sel = np.arange(rank * per, (rank + 1) * per if rank < world - 1 else n)
        assert n % world == 0
        loc = dict(sc)
        for k in ("position", "velocity", "rand"):
            loc[k] = sc[k][sel]
        s = ReplicatedPusher(dict(sc["spec"], device=rank), loc, rank, world)
        for _ in range(frames):
            s.step()
            s.density()
        s.sync()
        pos = s.sim.getPosition()
        # outside slab mode the accessors return particles in id order: row k is id id_base + k = sel[k]
        parts = dict(ids=sel.astype(np.uint32), pos=pos, vel=np.c_[s.sim.getVelocity(), np.ones(len(pos))],
                     rnd=s.sim.getRand(), avg=s.sim.getField("moments01_avg"), cnt=s.sim.getField("cell_count"), moved=1)
        out = _gather(rank, world, parts)
        if rank == 0:
            res = assemble(out)
            res["avg"], res["cnt"] = out[0]["avg"], out[0]["cnt"]  # replicated: every rank holds the whole grid
            res["avg_other"] = out[1]["avg"]
            np.savez(result_path, **res)
    finally:
        dist.destroy_process_group()


def gpu_worker_overflow(rank, world, port, result_path):
    """Exchange regions far too small for the scene: sync() must say so (never a silent loss of particles)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fusion_sim_b200 import Error
        from fusion_sim_b200.dist import SlabPusher
        sc = scene_for_dist()
        loc, ids = split_scene(sc, rank, world)
        s = SlabPusher(dict(sc["spec"], device=rank), loc, rank, world, halo_rows=8, cap_neighbour=2, cap_far=1)
        msg = ""
        try:
            for _ in range(3):
                s.step()
                s.density()
            s.sync()
        except Error as e:
            msg = str(e)
        out = _gather(rank, world, msg)
        if rank == 0:
            with open(result_path, "w") as f:
                f.write("\n".join(out))
    finally:
        dist.destroy_process_group()
