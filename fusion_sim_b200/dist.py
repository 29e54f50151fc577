"""Multi-GPU driver: slab decomposition of the grid along z, one process per GPU
(SURVEY.md section 8e).  torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in
the CPU tests); the data path is libfusionsim.so.

Per frame:
  step()     every rank pushes its own particles (no communication: the gather is nearest-grid-
             point, `halo_rows` rows of cell table beyond the slab cover the drift of a frame);
  density()  1. particles whose row left the slab (drift, or respawn anywhere in the domain) are
                packed by destination rank into FIXED-CAPACITY regions whose headers carry the record
                counts, and moved with one all-to-all of host-known sizes; records carry the global id.
                No count is read back: the frame has no host round trip (exchange="exact" keeps the
                host-synchronising all-to-all-v for arbitrary volumes);
             2. sort + per-cell sums of the owned rows (id order => same bits as on one GPU);
             3. the 5 boundary rows of per-cell sums go to each neighbour (send/recv) WHILE the stencil
                runs on the rows that need no halo row;
             4. 11x11 stencil + normalise + running average on the remaining (boundary) rows.

The exchange functions work on any backend object with the small interface used below, so the
host logic is tested on CPU (gloo, world_size 2) against the single-process oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

HALO_DEPOSIT = 5  # footprint radius, empic.js:949-952


def slab_bounds(nz: int, world: int):
    """Rows [b[k], b[k+1]) belong to rank k."""
    return [(k * nz) // world for k in range(world + 1)]


def exchange_records(send: torch.Tensor, send_counts, record_bytes: int, group=None):
    """All-to-all-v of packed particle records.  `send` is a uint8 tensor holding the records
    grouped by destination rank, `send_counts[k]` records for rank k.  Returns (recv, nrecv)."""
    world = dist.get_world_size(group)
    dev = send.device
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    rc_l = [int(v) for v in rc.tolist()]
    sc_l = [int(v) for v in send_counts]
    recv = torch.empty(sum(rc_l) * record_bytes, dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv, send[: sum(sc_l) * record_bytes],
                           output_split_sizes=[c * record_bytes for c in rc_l],
                           input_split_sizes=[c * record_bytes for c in sc_l], group=group)
    assert world == len(sc_l)
    return recv, sum(rc_l)


HEADER_BYTES = 16  # every exchange region: record count (u32) + padding, then the records


def region_capacities(rank: int, world: int, cap_neighbour: int, cap_far: int):
    """Records the region between `rank` and every other rank holds: drift only reaches the two
    neighbours, a respawn can land anywhere (the source pdf is global).  Symmetric in the two ranks,
    so the sender's and the receiver's host agree on every size without talking."""
    return [0 if k == rank else (int(cap_neighbour) if abs(k - rank) == 1 else int(cap_far)) for k in range(world)]


def region_bytes(cap: int, record_bytes: int) -> int:
    return (HEADER_BYTES + cap * record_bytes + 15) // 16 * 16


def exchange_regions(send: torch.Tensor, send_bytes, recv: torch.Tensor, recv_bytes, group=None):
    """All-to-all of fixed-size regions: region k of `send` goes to rank k, region k of `recv` comes
    from rank k.  Sizes are known to both hosts in advance; the record counts travel in the headers."""
    dist.all_to_all_single(recv, send, output_split_sizes=[int(b) for b in recv_bytes],
                           input_split_sizes=[int(b) for b in send_bytes], group=group)


def exchange_halo_begin(send_lo, send_hi, recv_lo, recv_hi, rank: int, world: int, group=None):
    """Boundary rows of the per-cell sums: send_lo -> rank-1 (its recv_hi), send_hi -> rank+1.
    Returns the pending work objects (wait with exchange_halo_finish)."""
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, send_lo, rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_lo, rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, send_hi, rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_hi, rank + 1, group))
    return dist.batch_isend_irecv(ops) if ops else []


def exchange_halo_finish(works):
    for w in works:
        w.wait()  # NCCL: the current stream waits, the host does not


def exchange_halo(send_lo, send_hi, recv_lo, recv_hi, rank: int, world: int, group=None):
    exchange_halo_finish(exchange_halo_begin(send_lo, send_hi, recv_lo, recv_hi, rank, world, group))


HALO_RELAX = 4  # rows of potential a relaxation launch may consume (most sweeps per launch)


def solve_fields_slab(backend, value: dict):
    """EXTENSION on a slab: the stages of the field solve with the halo exchanges
    between them.  `backend` offers fs_stage(stage, value, sweeps) and fs_exchange(name, nrows);
    the CUDA SlabPusher and the oracle-backed test rank both run through this function.  A launch
    of T <= 4 sweeps consumes T halo rows, so 4 boundary rows of the potential are refreshed
    before every launch; every rank recomputes the halo cells it needs with the same arithmetic,
    hence the result equals the single-GPU solve bit for bit."""
    sweeps = int(value["sweeps"])
    backend.fs_stage(0, value, 0)                 # charge source on the owned rows
    backend.fs_exchange("rho_src", HALO_RELAX)
    left = sweeps
    while left > 0:
        t = 4 if left >= 4 else (2 if left >= 2 else 1)
        backend.fs_exchange("phi", HALO_RELAX)
        backend.fs_stage(1, value, t)
        left -= t
    backend.fs_exchange("phi", HALO_RELAX)
    backend.fs_stage(2, value, 0)                 # E = -grad(phi)
    backend.fs_exchange("E", None)                # all halo rows of the cell table: the push gathers there
    backend.fs_stage(3, value, 0)                 # precalc()


class _DevPtr:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3}


def _dev_tensor(ptr: int, nbytes: int, device) -> torch.Tensor:
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)


class SlabPusher:
    """One rank of a slab-decomposed simulation; same member names as the single-GPU object.

    exchange="fixed" (default): fixed-capacity exchange regions, every count stays on the device, no host
    round trip in the frame; `cap_neighbour` / `cap_far` are the records a region to a neighbouring / any
    other rank holds per frame (an overflow is reported by sync()).  exchange="exact": all-to-all-v sized
    by counts read back to the host (two host waits per frame, any volume)."""

    def __init__(self, spec: dict, scene: dict, rank: int, world: int, halo_rows: int = 16, slack: float = 0.25,
                 exchange: str = "fixed", cap_neighbour: int | None = None, cap_far: int | None = None):
        from . import _lib
        from .pusher import CylindricalParticlePusher
        from .scenes import apply_scene
        self.rank, self.world = rank, world
        self.nz, self.nr = int(spec["nz"]), int(spec["nr"])
        self.bounds = slab_bounds(self.nz, world)
        n_local = len(scene["position"])
        lspec = dict(spec)
        lspec.update(slab_row0=self.bounds[rank], slab_rows=self.bounds[rank + 1] - self.bounds[rank],
                     halo_rows=halo_rows, nparticles_total=n_local,
                     capacity=int(n_local * (1 + slack)) + 4096, id_base=rank * n_local)
        self.device = torch.device("cuda", int(lspec.get("device", 0)))
        self.sim = CylindricalParticlePusher(lspec)
        apply_scene(self.sim, scene)
        self._lib = _lib
        L, h = _lib.lib(), self.sim.handle
        # one stream for the engine's kernels AND the NCCL collectives: torch orders a collective after
        # the current stream's work and makes the current stream wait for it, so no host
        # synchronisation is needed between a kernel and the exchange that consumes its output
        self.stream = torch.cuda.Stream(self.device)
        self.sim.sync()
        _lib.check(L.fsim_set_stream(h, C.c_void_p(self.stream.cuda_stream)))
        self.record_bytes = int(L.fsim_migrate_record_bytes(h))
        self.ncell_local = self.sim.ncell_local
        self._bounds_c = (C.c_int64 * (world + 1))(*self.bounds)
        self.exchange = exchange
        self._sent_exact = 0
        self.host_wait_s = 0.0
        self.comm_timing = False
        self._ev = {"migrate_exchange": [], "halo_exchange_exposed": []}
        if exchange == "fixed":
            nb = int(cap_neighbour) if cap_neighbour else max(4096, n_local // 1024)
            far = int(cap_far) if cap_far else max(1024, n_local // 16384)
            caps = region_capacities(rank, world, nb, far)
            caps_c = (C.c_int64 * world)(*caps)
            sb, rb = (C.c_int64 * world)(), (C.c_int64 * world)()
            sp, rp = C.c_void_p(), C.c_void_p()
            _lib.check(L.fsim_migrate_setup(h, self._bounds_c, world, rank, caps_c, caps_c, C.byref(sp), C.byref(rp), sb, rb))
            self._send_bytes, self._recv_bytes = list(sb), list(rb)
            assert self._send_bytes == [region_bytes(c, self.record_bytes) for c in caps]
            with torch.cuda.stream(self.stream):
                self._send = _dev_tensor(sp.value, sum(self._send_bytes), self.device)
                self._recv = _dev_tensor(rp.value, sum(self._recv_bytes), self.device)
            self.capacities = caps
        elif exchange != "exact":
            raise ValueError("exchange: 'fixed' or 'exact'")
        ptrs = [C.c_void_p() for _ in range(4)]
        nbytes = C.c_int64()
        _lib.check(L.fsim_halo_ptrs(h, *[C.byref(p) for p in ptrs], C.byref(nbytes)))
        with torch.cuda.stream(self.stream):
            self._halo = [_dev_tensor(p.value or 0, nbytes.value if p.value else 0, self.device) for p in ptrs]

    # -- frame --------------------------------------------------------------------------------
    def step(self):
        self.sim.step()

    def _mark(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record(self.stream)
        return e

    def migrate(self):
        L, h = self._lib.lib(), self.sim.handle
        if self.exchange == "fixed":
            self._lib.check(L.fsim_migrate_begin(h))
            with torch.cuda.stream(self.stream):
                e0 = self._mark() if self.comm_timing else None
                exchange_regions(self._send, self._send_bytes, self._recv, self._recv_bytes)
                if e0 is not None:
                    self._ev["migrate_exchange"].append((e0, self._mark()))
            self._lib.check(L.fsim_migrate_end(h))
            return
        import time
        counts = (C.c_int64 * self.world)()
        buf = C.c_void_p()
        t0 = time.perf_counter()
        self._lib.check(L.fsim_migrate_pack(h, self._bounds_c, self.world, self.rank, counts, C.byref(buf)))  # reads counts back
        self.host_wait_s += time.perf_counter() - t0
        sc = list(counts)
        with torch.cuda.stream(self.stream):
            send = _dev_tensor(buf.value or 0, sum(sc) * self.record_bytes, self.device)
            e0 = self._mark() if self.comm_timing else None
            t0 = time.perf_counter()
            recv, nrecv = exchange_records(send, sc, self.record_bytes)  # the count exchange is a host wait
            self.host_wait_s += time.perf_counter() - t0
            if e0 is not None:
                self._ev["migrate_exchange"].append((e0, self._mark()))
            self._lib.check(L.fsim_migrate_unpack(h, C.c_void_p(recv.data_ptr() if nrecv else 0), nrecv))
            del recv  # allocated on self.stream: the caching allocator reuses it in stream order
        self._sent_exact += sum(sc)

    @property
    def migrated(self) -> int:
        """Records this rank has sent so far (synchronises)."""
        if self.exchange != "fixed":
            return self._sent_exact
        v = C.c_int64()
        self._lib.check(self._lib.lib().fsim_migrate_stats(self.sim.handle, C.byref(v)))
        return int(v.value)

    def density(self):
        L, h = self._lib.lib(), self.sim.handle
        self.migrate()
        self._lib.check(L.fsim_density_begin(h))     # ... per-cell sums, boundary rows -> send buffers
        t = self._halo
        with torch.cuda.stream(self.stream):
            works = exchange_halo_begin(t[0], t[1], t[2], t[3], self.rank, self.world)
        self._lib.check(L.fsim_density_interior(h))  # stencil on the rows that read no halo row, under the exchange
        with torch.cuda.stream(self.stream):
            e0 = self._mark() if self.comm_timing else None
            exchange_halo_finish(works)
            if e0 is not None:
                self._ev["halo_exchange_exposed"].append((e0, self._mark()))
        self._lib.check(L.fsim_density_end(h))       # halo rows in, stencil on the boundary tiles

    def exchange_alone_ms(self, reps: int = 50) -> float:
        """Device time of ONE all-to-all of the (fixed-size) migration regions with nothing else going on: every
        rank enters together, so this is the collective's own cost -- what migrate_exchange exceeds it by inside a
        frame is time spent waiting for the slowest rank."""
        if self.exchange != "fixed":
            return float("nan")
        with torch.cuda.stream(self.stream):
            for _ in range(5):
                exchange_regions(self._send, self._send_bytes, self._recv, self._recv_bytes)
            self.stream.synchronize()
            dist.barrier()
            e0 = self._mark()
            for _ in range(reps):
                exchange_regions(self._send, self._send_bytes, self._recv, self._recv_bytes)
            e1 = self._mark()
            e1.synchronize()
        return float(e0.elapsed_time(e1)) / reps

    def comm_ms(self, reset: bool = True) -> dict:
        """Device time (ms, CUDA events on the frame's stream) spent in the exchanges since the last
        call: the migration all-to-all (it waits for the slowest peer) and the part of the halo
        exchange the interior stencil did not hide; plus the host time blocked in synchronisations."""
        self.sim.sync()
        out = {k: float(sum(a.elapsed_time(b) for a, b in v)) for k, v in self._ev.items()}
        out["host_wait"] = self.host_wait_s * 1e3
        if reset:
            for v in self._ev.values():
                v.clear()
            self.host_wait_s = 0.0
        return out

    # -- EXTENSION: self-consistent field solve on the slab -------------------------------------------
    def solveFields(self, value: dict):
        solve_fields_slab(self, value)

    def fs_stage(self, stage: int, value: dict, sweeps: int):
        src = 0 if value.get("source", "avg") == "avg" else 1
        self._lib.check(self._lib.lib().fsim_solve_fields_stage(
            self.sim.handle, stage, float(value["macro_weight"]), int(sweeps), float(value.get("omega", 1.0)), src))

    def _rows(self, name: str, first: int, nrows: int) -> torch.Tensor:
        ptr, nb = C.c_void_p(), C.c_int64()
        self._lib.check(self._lib.lib().fsim_field_rows(self.sim.handle, name.encode(), first, nrows,
                                                        C.byref(ptr), C.byref(nb)))
        return _dev_tensor(ptr.value or 0, nb.value, self.device)

    def fs_exchange(self, name: str, nrows):
        """Owned boundary rows -> the neighbours' halo rows, in place in device memory."""
        halo = int(self.sim.spec["halo_rows"])
        own = self.bounds[self.rank + 1] - self.bounds[self.rank]
        lo = self.bounds[self.rank] - max(0, self.bounds[self.rank] - halo)  # local index of the first owned row
        hi = lo + own
        h = halo if nrows is None else int(nrows)
        assert h <= halo and h <= own, "slab thinner than the halo"
        empty = torch.empty(0, dtype=torch.uint8, device=self.device)
        up, down = self.rank < self.world - 1, self.rank > 0
        with torch.cuda.stream(self.stream):
            exchange_halo(self._rows(name, lo, h) if down else empty, self._rows(name, hi - h, h) if up else empty,
                          self._rows(name, lo - h, h) if down else empty, self._rows(name, hi, h) if up else empty,
                          self.rank, self.world)

    # -- pass-throughs ----------------------------------------------------------------------------
    def sync(self):
        self.sim.sync()

    def mark(self, slot):
        self.sim.mark(slot)

    def elapsed_ms(self, a, b):
        return self.sim.elapsed_ms(a, b)

    def timing(self, on):
        self.sim.timing(on)

    def timing_reset(self):
        self.sim.timing_reset()

    def timing_get(self, name):
        return self.sim.timing_get(name)

    def render(self, out):
        return self.sim.render(out)

    def check_digest(self):
        return self.sim.check_digest()

    def render_async(self, out):
        return self.sim.render_async(out)

    def render_rows_async(self, out):
        return self.sim.render_rows_async(out)

    def draw_canvas(self):
        self.sim.draw_canvas()

    def set(self, value):
        self.sim.set(value)

    @property
    def launch_count(self):
        return self.sim.launch_count

    @property
    def n(self):
        return self.sim.n

    def gather_particles(self):
        """(ids, position[N][4], velocity[N][3], rand[N][4]) of all ranks, sorted by id, on rank 0."""
        ids = self.sim.getIds().astype(np.int64)
        parts = [ids, self.sim.getPosition(), self.sim.getVelocity(), self.sim.getRand()]
        out = [None] * self.world if self.rank == 0 else None
        dist.gather_object(parts, out, dst=0)
        if self.rank != 0:
            return None
        cat = [np.concatenate([o[k] for o in out]) for k in range(4)]
        order = np.argsort(cat[0], kind="stable")
        return tuple(c[order] for c in cat)

    def gather_field(self, name):
        """Owned rows of a cell field from every rank, assembled in global row order on rank 0."""
        a = self.sim.getField(name)
        row0 = max(0, self.bounds[self.rank] - int(self.sim.spec["halo_rows"]))
        lo = (self.bounds[self.rank] - row0) * self.nr
        hi = (self.bounds[self.rank + 1] - row0) * self.nr
        out = [None] * self.world if self.rank == 0 else None
        dist.gather_object(a[lo:hi], out, dst=0)
        return np.concatenate(out) if self.rank == 0 else None


class ReplicatedPusher:
    """MEASURED ALTERNATIVE to the slab decomposition: particles sharded by index,
    every table replicated on every rank.  No migration and no halo -- a particle never leaves its
    rank -- but every rank bins over the WHOLE grid, the per-cell sums and counts of all ranks are
    added with an all-reduce (32 + 4 bytes per cell of the whole grid per frame), and every rank runs
    the stencil over the whole grid.  The floating-point sums are added across ranks in the order the
    collective chooses, not in particle-id order: equal to the single-GPU result to rounding only
    (tests/test_dist.py holds it to 1e-12), where the slab decomposition is bit-identical."""

    def __init__(self, spec: dict, scene: dict, rank: int, world: int):
        from . import _lib
        from .pusher import CylindricalParticlePusher
        from .scenes import apply_scene
        self.rank, self.world = rank, world
        n_local = len(scene["position"])
        lspec = dict(spec, nparticles_total=n_local, id_base=rank * n_local)
        self.device = torch.device("cuda", int(lspec.get("device", 0)))
        self.sim = CylindricalParticlePusher(lspec)
        apply_scene(self.sim, scene)
        self._lib = _lib
        self.ncell_local = self.sim.ncell_local
        self.stream = torch.cuda.Stream(self.device)
        self.sim.sync()
        _lib.check(_lib.lib().fsim_set_stream(self.sim.handle, C.c_void_p(self.stream.cuda_stream)))
        sums, nb_s, cnt, nb_c = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64()
        _lib.check(_lib.lib().fsim_cellsum_ptrs(self.sim.handle, C.byref(sums), C.byref(nb_s), C.byref(cnt), C.byref(nb_c)))
        real = torch.float64 if self.sim.spec.get("precision", "f64") in ("f64", 0) else torch.float32
        self._sums = _dev_tensor(sums.value, nb_s.value, self.device).view(real)
        self._counts = _dev_tensor(cnt.value, nb_c.value, self.device).view(torch.int32)
        self.allreduce_bytes = nb_s.value + nb_c.value

    def step(self):
        self.sim.step()

    def density(self):
        L, h = self._lib.lib(), self.sim.handle
        self._lib.check(L.fsim_density_begin(h))
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self._sums)
            dist.all_reduce(self._counts)
        self._lib.check(L.fsim_density_end(h))

    def __getattr__(self, name):  # sync, mark, elapsed_ms, timing*, render*, set, launch_count, n, getField ...
        return getattr(self.sim, name)
