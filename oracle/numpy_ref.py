"""TEST INFRASTRUCTURE ONLY -- second, independent restatement of the reference shaders in
vectorised NumPy, used to cross-check the C oracle (oracle/fsim_oracle.c) on small cases.

PARITY UNPINNED (no reference tests / golden vectors exist): two
restatements written separately from public/javascripts/empic.js agreeing bit for bit is the
strongest check available.  NumPy element-wise arithmetic is IEEE round-to-nearest without
fused multiply-add, the same rules the C oracle is compiled under.
"""
from __future__ import annotations

import numpy as np


def tex(u, n):
    """NEAREST + CLAMP_TO_EDGE texel index (utilities.js:528-531); NaN -> 0."""
    with np.errstate(invalid="ignore"):
        t = u * u.dtype.type(n)
        idx = np.zeros(t.shape, np.int64)
        pos = t > 0
        big = t >= n
        idx[pos & ~big] = t[pos & ~big].astype(np.int64)
        idx[big] = n - 1
    return idx


def half_step(pos, vel, rnd, ent, R1, R2, R3, A, sink, invcdf, nr, nz, step_factor):
    """One leap-frog half-step on vec4 arrays (N,4); returns new (pos, vel, rnd).
    empic.js:800-807 (rand), :750-772 (velocity), :714-719 (position)."""
    T = pos.dtype.type
    with np.errstate(all="ignore"):
        # rand
        s = ent[tex(rnd[:, 2], 1024) + 1024 * tex(rnd[:, 3], 1024)]
        x = T(0.999) * rnd[:, 2:4] + T(0.001) * s[:, 2:4]
        m = rnd[:, 0:2] + s[:, 0:2]
        new_rnd = np.empty_like(rnd)
        new_rnd[:, 0:2] = np.where(m > T(1.0), m - T(1.0), m)
        new_rnd[:, 2:4] = T(4.0) * x * (T(1.0) - x)
        # velocity
        px, py, pz, alive = pos[:, 0], pos[:, 1], pos[:, 2], pos[:, 3]
        r = np.sqrt(px * px + py * py)
        dx, dy = px / r, py / r
        vr = vel[:, 0] * dx + vel[:, 1] * dy
        va = vel[:, 1] * dx - vel[:, 0] * dy
        vz = vel[:, 2]
        c = tex(r, nr) + nr * tex(pz, nz)
        r1, r2, r3, a = R1[c], R2[c], R3[c], A[c]
        c0 = (r1[:, 0] * vr + r1[:, 1] * va + r1[:, 2] * vz) + a[:, 0]
        c1 = (r2[:, 0] * vr + r2[:, 1] * va + r2[:, 2] * vz) + a[:, 1]
        c2 = (r3[:, 0] * vr + r3[:, 1] * va + r3[:, 2] * vz) + a[:, 2]
        nv = np.stack([c0 * dx - c1 * dy, c0 * dy + c1 * dx, c2], axis=1)
        fresh = T(0.001) * (T(2.0) * rnd[:, 0:3] - T(1.0))
        new_vel = np.ones_like(vel)
        new_vel[:, 0:3] = np.where((alive > T(0.5))[:, None], nv, fresh)
        new_vel[:, 3] = np.where(alive > T(0.5), T(1.0), T(0.001) * T(1.0))  # w: written, never read (:772)
        # position
        nxt = pos[:, 0:3] + T(step_factor) * new_vel[:, 0:3]
        rn = np.sqrt(nxt[:, 0] * nxt[:, 0] + nxt[:, 1] * nxt[:, 1])
        ok = ~(np.isnan(rn) | np.isnan(nxt[:, 2]))
        keep = np.zeros(len(pos), bool)
        cc = tex(rn[ok], nr) + nr * tex(nxt[ok, 2], nz)
        keep[ok] = sink[cc, 0] > T(0.5)
        t = invcdf[tex(rnd[:, 0], 512) + 512 * tex(rnd[:, 1], 512)]
        new_pos = np.empty_like(pos)
        new_pos[:, 0:3] = nxt
        new_pos[:, 3] = 1.0
        resp = np.stack([t[:, 0], np.zeros(len(pos), pos.dtype), t[:, 1], np.zeros(len(pos), pos.dtype)], 1)
        new_pos[~keep] = resp[~keep]
    return new_pos, new_vel, new_rnd


def precalc(E, B, h, k13, k31, kr, kz, corrected=False):
    """programPre1/2/3/A, empic.js:519-527, 558-566, 598-606, 640-647.  (ncell,4) in/out."""
    T = B.dtype.type
    h, k13, k31, kr, kz = T(h), T(k13), T(k31), T(kr), T(kz)
    Bx, By, Bz = B[:, 0], B[:, 1], B[:, 2]
    Ex, Ey, Ez = E[:, 0], E[:, 1], E[:, 2]
    Bmag = np.sqrt(Bx * Bx + By * By + Bz * Bz)
    hB2 = h * h * Bmag * Bmag
    f = T(2.0) / (T(1.0) + hB2)
    one_m = T(1.0) - hB2 * f
    one = np.ones_like(Bx)
    R1 = np.stack([one_m + f * h * h * Bx * Bx, f * h * (Bz + h * Bx * By),
                   (f * h * (-By + h * Bx * Bz)) * k13, one], 1)
    R2 = np.stack([f * h * (-Bz + h * By * Bx), one_m + f * h * h * By * By,
                   (f * h * (Bx + h * By * Bz)) * k13, one], 1)
    R3 = np.stack([(f * h * (By + h * Bz * Bx)) * k31, (f * h * (-Bx + h * Bz * By)) * k31,
                   one_m + f * h * h * Bz * Bz, one], 1)
    cx, cy, cz = Ey * Bz - Ez * By, Ez * Bx - Ex * Bz, Ex * By - Ey * Bx
    d = Ex * Bx + Ey * By + Ez * Bz
    t1 = h * (T(2.0) - hB2 * f)
    t2 = h * h * f
    c = T(2.998e8)
    if corrected:
        ax = (t1 * Ex + t2 * (cx + h * d * Bx)) / c
        ay = (t1 * Ey + t2 * (cy + h * d * By)) / c
        az = (t1 * Ez + t2 * (cz + h * d * Bz)) / c
    else:
        hd = h * d
        ax = (t1 * Ex + t2 * (cx + hd)) / c
        ay = (t1 * Ey + t2 * (cy + hd)) / c
        az = (t1 * Ez + t2 * (cz + hd)) / c
    A = np.stack([ax * kr, ay * kr, az * kz, one], 1)
    return R1, R2, R3, A


def boris_rotation_textbook(B, h):
    """Textbook Boris rotation matrix for t = h*B (double precision, unit factors 1):
    v' = v + (2/(1+t^2)) (v + v x t) x t, written for the reference's sign convention
    (force q v x B with the (r, theta, z) components of B)."""
    B = np.asarray(B, np.float64)
    t = h * B
    t2 = np.sum(t * t, axis=-1)
    f = 2.0 / (1.0 + t2)
    I = np.eye(3)

    def cross_matrix(v):  # [v]x such that [v]x w = v x w
        return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])

    out = np.empty(B.shape[:-1] + (3, 3))
    for idx in np.ndindex(B.shape[:-1]):
        tx = cross_matrix(t[idx])
        # v x t = -[t]x v ; (v + v x t) x t
        M = I - tx
        out[idx] = I + f[idx] * (-(tx @ M))
    return out


def cell_sums(pos, vel, nr, nz):
    """Per-cell NGP sums of the sprite colours in particle order (empic.js:994-1006)."""
    T = pos.dtype.type
    S = np.zeros((nr * nz, 4), pos.dtype)
    count = np.zeros(nr * nz, np.uint32)
    with np.errstate(all="ignore"):
        r = np.sqrt(pos[:, 0] * pos[:, 0] + pos[:, 1] * pos[:, 1])
        dx, dy = pos[:, 0] / r, pos[:, 1] / r
        vr = vel[:, 0] * dx + vel[:, 1] * dy
        va = vel[:, 1] * dx - vel[:, 0] * dy
        col = np.stack([T(0.001) * vr, T(0.001) * va, T(0.001) * vel[:, 2],
                        np.full(len(pos), T(0.001) * T(1.0))], 1)
        xw, yw = r * T(nr), pos[:, 2] * T(nz)
        ok = (xw >= 0) & (xw < nr) & (yw >= 0) & (yw < nz)
    for p in np.nonzero(ok)[0]:
        c = int(xw[p]) + nr * int(yw[p])
        S[c] += col[p]
        count[c] += 1
    return S, count


def convolve(S, shape, nr, nz):
    """moments01 = S (*) shape with the mirror sources of each weight added first (the footprint is
    mirror-symmetric): classes di = 0..5 outer, dj = 0..5 inner; inside a class the two sources of a row
    first, then the two rows: (S[-di,-dj] + S[+di,-dj]) + (S[-di,+dj] + S[+di,+dj]), duplicates once; zero
    weights skipped; sources outside the grid are exact zeros."""
    S2 = np.zeros((nz + 10, nr + 10, 4), S.dtype)
    S2[5:5 + nz, 5:5 + nr] = S.reshape(nz, nr, 4)
    out = np.zeros((nz, nr, 4), S.dtype)

    def src(di, dj):
        return S2[5 + dj:5 + dj + nz, 5 + di:5 + di + nr]

    def row_pair(di, dj):
        return src(-di, dj) + src(di, dj) if di else src(0, dj)

    for di in range(6):
        for dj in range(6):
            w = shape[(5 + di) + 11 * (5 + dj)]
            if w == 0:
                continue
            tot = row_pair(di, -dj) + row_pair(di, dj) if dj else row_pair(di, 0)
            out = out + tot * w
    return out.reshape(nr * nz, 4)


def normalize_ema(mom, avg, nr, nz):
    """empic.js:1053-1056 then avg_frag :274-277 with u_ratio 0.01."""
    T = mom.dtype.type
    i = np.arange(nr * nz) % nr
    u = (i.astype(mom.dtype) + T(0.5)) / T(nr)
    a = mom[:, 3]
    with np.errstate(all="ignore"):
        M = np.where((a > 0)[:, None],
                     np.stack([mom[:, 0] / a, mom[:, 1] / a, mom[:, 2] / a, a], 1), T(0.0)).astype(mom.dtype)
        norm = T(1000.0) * M * T(0.5) / u[:, None]
        new_avg = T(0.01) * norm + (T(1.0) - T(0.01)) * avg
    return norm, new_avg


# ---- EXTENSION: self-consistent electrostatic field solve ----------------------
def relax_coeffs(nr, dr, dz):
    """[nr][4] = cE cW cZ cB in float64 (oracle/fsim_oracle_fields_impl.h)."""
    i = np.arange(nr, dtype=np.float64)
    rc = (i + 0.5) * dr * dr
    aE, aW, aZ = (i + 1.0) / rc, i / rc, np.full(nr, 1.0 / (dz * dz))
    aC = aE + aW + 2.0 * aZ
    return np.stack([aE / aC, aW / aC, aZ / aC, 1.0 / aC], 1)


def relax(phi, src, coef64, omega, sweeps, nr, nz):
    """`sweeps` weighted-Jacobi sweeps of the 5-point cylindrical operator, grounded ghost cells."""
    T = phi.dtype.type
    k = coef64.astype(phi.dtype)
    cE, cW, cZ, cB = (k[:, q][None, :] for q in range(4))
    om = T(omega)
    one_m = T(1.0) - om
    p = phi.reshape(nz, nr).copy()
    s = src.reshape(nz, nr)
    for _ in range(sweeps):
        g = np.zeros((nz + 2, nr + 2), phi.dtype)
        g[1:-1, 1:-1] = p
        pE, pW, pN, pS = g[1:-1, 2:], g[1:-1, :-2], g[2:, 1:-1], g[:-2, 1:-1]
        t = ((cE * pE + cW * pW) + cZ * (pN + pS)) + cB * s
        p = om * t + one_m * p
    return p.reshape(-1)


def efield(phi, nr, nz, inv2dr, inv2dz):
    T = phi.dtype.type
    p = phi.reshape(nz, nr)
    g = np.zeros((nz + 2, nr + 2), phi.dtype)
    g[1:-1, 1:-1] = p
    g[1:-1, 0] = p[:, 0]  # axis: phi_W := phi
    E = np.zeros((nz, nr, 4), phi.dtype)
    E[..., 0] = -((g[1:-1, 2:] - g[1:-1, :-2]) * T(inv2dr))
    E[..., 2] = -((g[2:, 1:-1] - g[:-2, 1:-1]) * T(inv2dz))
    E[..., 3] = 1
    return E.reshape(-1, 4)
