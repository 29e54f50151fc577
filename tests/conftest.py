import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def small_scene(seed=7, nr=48, nz=96, n=4096, precision="f64", speed=0.002, with_E=False,
                blob=None):
    """A C1-like scene on a small grid: same proportions, same two opposing loops."""
    from fusion_sim_b200.scenes import c1_sink_source
    rng = np.random.Generator(np.random.PCG64(seed))
    side = int(round(n ** 0.5))
    spec = dict(radius=1.0, height=2.0, nr=nr, nz=nz, dt=2e-9, nparticles=side,
                particle_mass=1.67e-27, particle_charge=1.602e-19, precision=precision,
                keep_moments=True)
    if side * side != n:
        spec["nparticles_total"] = n
    lo, hi = blob if blob is not None else (0.2, 0.2)
    position = np.stack([lo * (rng.random(n) - 0.5) * 2, lo * (rng.random(n) - 0.5) * 2,
                         1.0 + hi * (rng.random(n) - 0.5) * 2], 1)
    velocity = speed * (rng.random((n, 3)) - 0.5)
    rand = rng.random((n, 4))
    entropy = rng.random((1024 * 1024, 4))
    sink, source = c1_sink_source(nr, nz)
    sc = dict(spec=spec, position=position, velocity=velocity, sink_mask=sink, source_pdf=source,
              rand=rand, entropy=entropy, loops=[(0.8, 2.0, -1.0e7), (0.8, 0.0, 1.0e7)])
    if with_E:
        sc["E"] = 1.0e5 * (rng.random((nr, nz, 3)) - 0.5)
    return sc


def assert_same(a, b, what=""):
    """Bit-for-bit equality of values; NaN == NaN (CPU and GPU differ in NaN payload bits)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        ok = (a == b) | (np.isnan(a) & np.isnan(b))
    else:
        ok = a == b
    if not ok.all():
        bad = np.argwhere(~ok)
        k = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {a.size} differ; first at {k}: {a[k]!r} vs {b[k]!r}")


def deposit_reorder_bound(pos, vel, nr, nz, shape, dtype):
    """Per cell and channel, how far two floating-point sums of the SAME deposit terms in DIFFERENT orders can
    lie apart: the literal sprite raster adds a pixel's terms in particle order (GL blend order, empic.js:1473-1478),
    the canonical form adds per-cell sums times weights.  With n terms t_k each sum is within gamma_n * sum|t_k| of
    the exact value (gamma_n ~ n*eps), the per-cell partial sums of the canonical form add one more rounding per
    term: |literal - canonical| <= 2 (n + 2) eps sum|t_k|.  Returns that bound, shape [nr*nz][4], in fp64."""
    pos = np.asarray(pos, np.float64)
    vel = np.asarray(vel, np.float64)
    with np.errstate(all="ignore"):
        r = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2)
        dx, dy = pos[:, 0] / r, pos[:, 1] / r
        col = 0.001 * np.abs(np.stack([vel[:, 0] * dx + vel[:, 1] * dy, vel[:, 1] * dx - vel[:, 0] * dy,
                                       vel[:, 2], np.ones(len(pos))], 1))
        xw, yw = r * nr, pos[:, 2] * nz
        ok = (xw >= 0) & (xw < nr) & (yw >= 0) & (yw < nz) & ~np.isnan(col).any(1)
    cell = xw[ok].astype(np.int64) + nr * yw[ok].astype(np.int64)
    A = np.zeros((nz + 10, nr + 10, 4))
    N = np.zeros((nz + 10, nr + 10))
    np.add.at(A.reshape(-1, 4), (cell // nr + 5) * (nr + 10) + cell % nr + 5, col[ok])
    np.add.at(N.reshape(-1), (cell // nr + 5) * (nr + 10) + cell % nr + 5, 1.0)
    w = np.asarray(shape, np.float64).reshape(11, 11)
    T = np.zeros((nz, nr, 4))
    K = np.zeros((nz, nr))
    for tj in range(11):
        for ti in range(11):
            if w[tj, ti] == 0:
                continue
            T += A[10 - tj:10 - tj + nz, 10 - ti:10 - ti + nr] * w[tj, ti]
            K += N[10 - tj:10 - tj + nz, 10 - ti:10 - ti + nr]
    eps = np.finfo(dtype).eps
    return (2.0 * (K[..., None] + 2.0) * eps * T).reshape(nr * nz, 4) + float(np.finfo(dtype).tiny)  # + one denormal step
