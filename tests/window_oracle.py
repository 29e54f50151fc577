"""TEST CODE (may use oracle/): the oracle's deposit on a WINDOW of grid rows, for parity checks at sizes
where only part of the grid is compared (tests/test_parity_sizes_gpu.py, C5).  Held to the whole-grid
oracle by tests/test_oracle.py::test_window_oracle_equals_the_whole_grid_oracle."""
import ctypes as C

import numpy as np


def oracle_window(gp, gv, prev_avg_rows, nr, nz, w0, w1, threads=4):
    """Per-cell counts / sums / running average of rows [w0, w1) as the oracle computes them from the
    particles (all of them, id order) whose sprite can reach those rows: the oracle's per-cell pass and
    stencil run on the sub-grid of rows [lo, hi) = [w0-5, w1+5) clipped to the grid -- the grid edge of
    the sub-grid coincides with the true edge exactly where the window touches it, elsewhere the 5
    outer rows are discarded."""
    from oracle.numpy_ref import tex
    from oracle import oracle as orc
    lo, hi = max(0, w0 - 5), min(nz, w1 + 5)
    rows = hi - lo
    yw = gp[:, 2] * nz
    inside = (yw >= 0) & (yw < nz)  # NaN z: clipped (the r test is the oracle's own)
    row = tex(gp[:, 2], nz)
    sel = np.nonzero(inside & (row >= lo) & (row < hi))[0]  # ascending id = GL primitive order
    pos = np.ascontiguousarray(gp[sel])
    pos[:, 2] = ((row[sel] - lo) + 0.5) / rows  # same texel row on the sub-grid
    pos[:, 3] = 1.0
    vel = np.ones((len(sel), 4))
    vel[:, :3] = gv[sel]
    S = np.zeros((nr * rows, 4))
    cnt = np.zeros(nr * rows, np.uint32)
    f = lambda name: getattr(orc.lib(), name + "_f64")
    f("orc_cell_sums")(C.c_int64(len(sel)), orc._p(pos), orc._p(vel), C.c_int64(nr), C.c_int64(rows),
                       orc._p(S), orc._p(cnt), None)
    mom = np.zeros_like(S)
    shape = orc.shape_table(False)
    f("orc_convolve")(C.c_int64(nr), C.c_int64(rows), orc._p(S), orc._p(shape), orc._p(mom), C.c_int(threads))
    norm = np.zeros_like(S)
    avg = np.zeros_like(S)
    a, b = (w0 - lo) * nr, (w1 - lo) * nr
    avg[a:b] = prev_avg_rows
    f("orc_normalize_ema")(C.c_int64(nr), C.c_int64(rows), orc._p(mom), orc._p(norm), orc._p(avg), C.c_int(threads))
    return cnt[a:b], S[a:b], avg[a:b], len(sel)
