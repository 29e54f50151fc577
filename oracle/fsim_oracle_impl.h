/* fsim_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, precision-generic body).
 *
 * A CPU restatement of the GLSL shaders of kcdodd/fusion-sim's particle pusher
 * (public/javascripts/empic.js).  Included twice by fsim_oracle.c, once with
 * REAL=double / SFX=f64 and once with REAL=float / SFX=f32.
 *
 * PINNING: the reference ships no tests, golden vectors or fixtures and cannot run as a
 * whole here (WebGL + DOM, no JS engine).  Its SHADERS can: their
 * GLSL text is read out of /root/reference/public/javascripts/empic.js and executed by
 * oracle/glsl_interp.py (tests/golden/make_reference_vectors.py); this file must
 * reproduce those outputs bit for bit (tests/test_reference_glsl.py: static fields,
 * precalc, half-steps, sprite deposit, normalise, running average, canvas; fp64 and
 * fp32).  Still unpinned: what GLSL ES 1.00 / GL leave open (rounding of / and sqrt,
 * evaluation order, NaN texture coordinates, fixed-function blending and rasterisation)
 * -- there the documented IEEE / left-to-right choices below apply.  The host-side
 * JavaScript of set() that builds the inverse-cdf table is pinned by executing a mechanical
 * transliteration of it (tests/golden/js_transliterate.py).  Further cross-checks: an
 * independent NumPy restatement (oracle/numpy_ref.py) and analytic invariants (tests/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call into this file.  The product never does.
 *
 * Conventions restated from the reference:
 *   - every texture is RGBA (4 reals per texel), texel (i,j) at 4*(i + j*W)
 *     (empic.js:1162), sampled NEAREST + CLAMP_TO_EDGE (utilities.js:528-531);
 *   - arithmetic is evaluated left to right exactly as the GLSL is written,
 *     IEEE round-to-nearest, no fused multiply-add (compile -ffp-contract=off);
 *   - documented choices where GLSL ES 1.00 leaves behaviour undefined:
 *       * a NaN texture coordinate samples texel 0;
 *       * a particle whose pushed position has NaN r or z is absorbed and
 *         respawned (SURVEY.md section 7 "edge semantics");
 *       * a point sprite whose centre lies outside the render target, or is
 *         NaN, is discarded (GLES2 point clipping).
 */

#define ORC_CAT2(a, b) a##_##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC(name) ORC_CAT(name, SFX)
#define RC(x) ((REAL)(x))

/* NEAREST + CLAMP_TO_EDGE texel index along one axis (utilities.js:528-531). */
static inline int64_t ORC(tex)(REAL u, int64_t n)
{
    REAL t = u * (REAL)n;
    if (!(t > RC(0))) return 0; /* <=0 clamps to edge; NaN: documented rule */
    if (t >= (REAL)n) return n - 1;
    return (int64_t)t;
}

/* ------------------------------------------------------------------------- */
/* One leap-frog half-step = programStepRand + programStepVelocity +
 * programStepPosition for every particle (empic.js:1438-1451 or :1454-1467).
 * All three shaders of a half-step read the rand texture from BEFORE this
 * half-step's rand update (u_rand bindings empic.js:819,832,849 / 894,907,924),
 * the position shader reads the OLD position and the NEW velocity
 * (empic.js:847-848 / 922-923).                                              */
void ORC(orc_half_step)(int64_t n, REAL *pos, REAL *vel, REAL *rnd,
                        const REAL *ent, const REAL *R1, const REAL *R2,
                        const REAL *R3, const REAL *A, const REAL *sink,
                        const REAL *invcdf, int64_t nr, int64_t nz,
                        double step_factor_d, int64_t *cell_out, int nthreads)
{
    const REAL sf = (REAL)step_factor_d; /* u_step_factor = dt*c, empic.js:852 */
    int64_t p;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (p = 0; p < n; ++p) {
        REAL *P = pos + 4 * p, *V = vel + 4 * p, *Q = rnd + 4 * p;
        const REAL q0 = Q[0], q1 = Q[1], q2 = Q[2], q3 = Q[3];

        /* --- programStepRandA/B, empic.js:800-807 ------------------------- */
        {
            const REAL *s = ent + 4 * (ORC(tex)(q2, FSIM_N_ENTROPY) +
                                       FSIM_N_ENTROPY * ORC(tex)(q3, FSIM_N_ENTROPY));
            REAL x0 = RC(FSIM_RNG_KEEP) * q2 + RC(FSIM_RNG_MIX) * s[2];
            REAL x1 = RC(FSIM_RNG_KEEP) * q3 + RC(FSIM_RNG_MIX) * s[3];
            REAL m0 = q0 + s[0];
            REAL m1 = q1 + s[1];
            Q[0] = (m0 > RC(1.0)) ? m0 - RC(1.0) : m0;
            Q[1] = (m1 > RC(1.0)) ? m1 - RC(1.0) : m1;
            Q[2] = RC(4.0) * x0 * (RC(1.0) - x0);
            Q[3] = RC(4.0) * x1 * (RC(1.0) - x1);
        }

        /* --- step_velocity_frag, empic.js:750-772 ------------------------- */
        REAL nvx, nvy, nvz;
        {
            REAL x = P[0], y = P[1], z = P[2], alive = P[3];
            REAL vx = V[0], vy = V[1], vz = V[2];
            REAL r = ORC(orc_sqrt)(x * x + y * y);
            REAL dx = x / r, dy = y / r;
            REAL vr = vx * dx + vy * dy;
            REAL va = vy * dx - vx * dy;
            int64_t c = ORC(tex)(r, nr) + nr * ORC(tex)(z, nz);
            const REAL *r1 = R1 + 4 * c, *r2 = R2 + 4 * c, *r3 = R3 + 4 * c,
                       *a = A + 4 * c;
            if (cell_out) cell_out[p] = c;
            REAL c0 = (r1[0] * vr + r1[1] * va + r1[2] * vz) + a[0];
            REAL c1 = (r2[0] * vr + r2[1] * va + r2[2] * vz) + a[1];
            REAL c2 = (r3[0] * vr + r3[1] * va + r3[2] * vz) + a[2];
            REAL nvw; /* the w texel channel: written, never read by any shader */
            if (alive > RC(0.5)) {
                nvx = c0 * dx - c1 * dy;
                nvy = c0 * dy + c1 * dx;
                nvz = c2;
                nvw = RC(1.0);
            } else { /* just respawned: fresh random velocity, empic.js:772 */
                nvx = RC(FSIM_RESPAWN_SPEED) * (RC(2.0) * q0 - RC(1.0));
                nvy = RC(FSIM_RESPAWN_SPEED) * (RC(2.0) * q1 - RC(1.0));
                nvz = RC(FSIM_RESPAWN_SPEED) * (RC(2.0) * q2 - RC(1.0));
                nvw = RC(FSIM_RESPAWN_SPEED) * RC(1.0); /* 0.001 * vec4(.., 1.0): found by running the shader text */
            }
            V[0] = nvx; V[1] = nvy; V[2] = nvz; V[3] = nvw;
        }

        /* --- step_position_frag, empic.js:714-719 ------------------------- */
        {
            REAL nx = P[0] + sf * nvx;
            REAL ny = P[1] + sf * nvy;
            REAL nzp = P[2] + sf * nvz;
            if (g_periodic_z) { /* EXTENSION: periodic in z */
                nzp = nzp - ORC(orc_floor)(nzp);
                if (nzp >= RC(1.0)) nzp = RC(0.0);
            }
            REAL r = ORC(orc_sqrt)(nx * nx + ny * ny);
            int keep = 0;
            if (r == r && nzp == nzp) { /* NaN => absorbed (documented rule) */
                int64_t c = ORC(tex)(r, nr) + nr * ORC(tex)(nzp, nz);
                keep = sink[4 * c] > RC(0.5);
            }
            if (keep) {
                P[0] = nx; P[1] = ny; P[2] = nzp; P[3] = RC(1.0);
            } else {
                const REAL *t = invcdf + 4 * (ORC(tex)(q0, FSIM_N_INVCDF) +
                                              FSIM_N_INVCDF * ORC(tex)(q1, FSIM_N_INVCDF));
                P[0] = t[0]; P[1] = RC(0.0); P[2] = t[1]; P[3] = RC(0.0);
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* programPre1/2/3 and programPreA, empic.js:519-527, 558-566, 598-606, 640-647.
 * k13 = N(factor_r/factor_z), k31 = N(factor_z/factor_r), kr = N(factor_r),
 * kz = N(factor_z): the toFixed(20) literals the reference bakes into GLSL,
 * parsed back by the caller.  corrected != 0 replaces the reference's
 * "cross(E,B) + u_h*dot(E,B)" (scalar added to a vec3, empic.js:645) by the
 * textbook h*(E.B)*B.                                                         */
void ORC(orc_precalc)(int64_t ncell, double h_d, double k13_d, double k31_d,
                      double kr_d, double kz_d, const REAL *E, const REAL *B,
                      REAL *R1, REAL *R2, REAL *R3, REAL *A, int corrected,
                      int nthreads)
{
    const REAL h = (REAL)h_d, k13 = (REAL)k13_d, k31 = (REAL)k31_d,
               kr = (REAL)kr_d, kz = (REAL)kz_d;
    int64_t c;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (c = 0; c < ncell; ++c) {
        REAL Bx = B[4 * c], By = B[4 * c + 1], Bz = B[4 * c + 2];
        REAL Ex = E[4 * c], Ey = E[4 * c + 1], Ez = E[4 * c + 2];
        REAL Bmag = ORC(orc_sqrt)(Bx * Bx + By * By + Bz * Bz);
        REAL hB2 = h * h * Bmag * Bmag;
        REAL f = RC(2.0) / (RC(1.0) + hB2);
        REAL one_m = RC(1.0) - hB2 * f;

        R1[4 * c + 0] = one_m + f * h * h * Bx * Bx;
        R1[4 * c + 1] = f * h * (Bz + h * Bx * By);
        R1[4 * c + 2] = (f * h * (-By + h * Bx * Bz)) * k13;
        R1[4 * c + 3] = RC(1.0);

        R2[4 * c + 0] = f * h * (-Bz + h * By * Bx);
        R2[4 * c + 1] = one_m + f * h * h * By * By;
        R2[4 * c + 2] = (f * h * (Bx + h * By * Bz)) * k13;
        R2[4 * c + 3] = RC(1.0);

        R3[4 * c + 0] = (f * h * (By + h * Bz * Bx)) * k31;
        R3[4 * c + 1] = (f * h * (-Bx + h * Bz * By)) * k31;
        R3[4 * c + 2] = one_m + f * h * h * Bz * Bz;
        R3[4 * c + 3] = RC(1.0);

        REAL cx = Ey * Bz - Ez * By;
        REAL cy = Ez * Bx - Ex * Bz;
        REAL cz = Ex * By - Ey * Bx;
        REAL d = Ex * Bx + Ey * By + Ez * Bz;
        REAL t1 = h * (RC(2.0) - hB2 * f);
        REAL t2 = h * h * f;
        REAL ax, ay, az;
        if (corrected) {
            ax = (t1 * Ex + t2 * (cx + h * d * Bx)) / RC(FSIM_C_LIGHT);
            ay = (t1 * Ey + t2 * (cy + h * d * By)) / RC(FSIM_C_LIGHT);
            az = (t1 * Ez + t2 * (cz + h * d * Bz)) / RC(FSIM_C_LIGHT);
        } else {
            REAL hd = h * d;
            ax = (t1 * Ex + t2 * (cx + hd)) / RC(FSIM_C_LIGHT);
            ay = (t1 * Ey + t2 * (cy + hd)) / RC(FSIM_C_LIGHT);
            az = (t1 * Ez + t2 * (cz + hd)) / RC(FSIM_C_LIGHT);
        }
        A[4 * c + 0] = ax * kr;
        A[4 * c + 1] = ay * kr;
        A[4 * c + 2] = az * kz;
        A[4 * c + 3] = RC(1.0);
    }
}

/* ------------------------------------------------------------------------- */
/* programCurrentLoopShape, empic.js:308-326: Biot-Savart table of a loop of
 * normalised radius u_R over the nr x nz texel centres.  costab[k] =
 * cos(PI*(k+0.5)/1000) is supplied by the caller (host libm, rounded to REAL)
 * so that CPU and GPU see the same cosines.                                   */
void ORC(orc_loop_shape)(int64_t nr, int64_t nz, double R_d, const REAL *costab,
                         REAL *out, int nthreads)
{
    const REAL R = (REAL)R_d;
    const REAL constant = R * RC(FSIM_QUAD_SCALE) * RC(FSIM_MU0) /
                          (RC(4.0) * RC(FSIM_PI_GLSL));
    int64_t c;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (c = 0; c < nr * nz; ++c) {
        int64_t i = c % nr, j = c / nr;
        REAL u = ((REAL)i + RC(0.5)) / (REAL)nr;
        REAL v = ((REAL)j + RC(0.5)) / (REAL)nz;
        REAL Bx = RC(0.0), Bz = RC(0.0);
        for (int k = 0; k < FSIM_NQUAD; ++k) {
            REAL cosine = costab[k];
            REAL r = ORC(orc_sqrt)(R * R + u * u + v * v - RC(2.0) * u * R * cosine);
            REAL factor = (r > RC(0.0)) ? constant * RC(1.0) / (r * r * r) : RC(0.0);
            Bx += v * factor * cosine;
            Bz += factor * (R - u * cosine);
        }
        out[4 * c + 0] = Bx; out[4 * c + 1] = RC(0.0);
        out[4 * c + 2] = Bz; out[4 * c + 3] = RC(1.0);
    }
}

/* programCurrentLoop, empic.js:367-377, blended ONE,ONE into B (:1358-1362). */
void ORC(orc_add_current_loop)(int64_t nr, int64_t nz, double R_d, double Z_d,
                               double I_d, const REAL *half, const REAL *tenth,
                               REAL *B, int nthreads)
{
    const REAL R = (REAL)R_d, Z = (REAL)Z_d, I = (REAL)I_d;
    int64_t c;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (c = 0; c < nr * nz; ++c) {
        int64_t i = c % nr, j = c / nr;
        REAL u = ((REAL)i + RC(0.5)) / (REAL)nr;
        REAL v = ((REAL)j + RC(0.5)) / (REAL)nz;
        REAL a = u / R;
        REAL b = (v - Z) / R;
        REAL sgn = (b > RC(0.0)) ? RC(1.0) : ((b < RC(0.0)) ? RC(-1.0) : RC(0.0));
        REAL ab = (b < RC(0.0)) ? -b : b;
        const REAL *t;
        if (a > RC(FSIM_LOOP_FAR) || b > RC(FSIM_LOOP_FAR))
            t = tenth + 4 * (ORC(tex)(a / RC(10.0), nr) + nr * ORC(tex)(ab / RC(10.0), nz));
        else
            t = half + 4 * (ORC(tex)(a / RC(2.0), nr) + nr * ORC(tex)(ab / RC(2.0), nz));
        B[4 * c + 0] = B[4 * c + 0] + (I * sgn) * t[0];
        B[4 * c + 1] = B[4 * c + 1] + (I * RC(1.0)) * t[1];
        B[4 * c + 2] = B[4 * c + 2] + (I * RC(1.0)) * t[2];
        B[4 * c + 3] = B[4 * c + 3] + (I * RC(1.0)) * t[3];
    }
}

/* programCurrentZ :404, programBZ :429, programBTheta :454 (gl_FragColor +=
 * on an undefined value is taken as "=", then blended ONE,ONE into B).      */
void ORC(orc_add_uniform)(int64_t nr, int64_t nz, int kind, double val_d, REAL *B)
{
    const REAL val = (REAL)val_d;
    for (int64_t c = 0; c < nr * nz; ++c) {
        int64_t i = c % nr;
        REAL u = ((REAL)i + RC(0.5)) / (REAL)nr;
        if (kind == 0) /* addCurrentZ */
            B[4 * c + 1] = B[4 * c + 1] +
                           val * RC(FSIM_MU0) / (RC(2.0) * RC(FSIM_PI_GLSL) * u);
        else if (kind == 1) /* addBZ */
            B[4 * c + 2] = B[4 * c + 2] + val;
        else /* addBTheta */
            B[4 * c + 1] = B[4 * c + 1] + val;
        B[4 * c + 3] = B[4 * c + 3] + RC(1.0);
    }
}

/* ------------------------------------------------------------------------- */
/* Vertex shader of programMoments01, empic.js:994-1006: sprite centre and colour. */
static inline int ORC(sprite)(const REAL *P, const REAL *V, int64_t nr, int64_t nz,
                              REAL col[4], REAL *xw, REAL *yw)
{
    REAL x = P[0], y = P[1], z = P[2];
    REAL r = ORC(orc_sqrt)(x * x + y * y);
    REAL dx = x / r, dy = y / r;
    REAL vr = V[0] * dx + V[1] * dy;
    REAL va = V[1] * dx - V[0] * dy;
    col[0] = RC(FSIM_DEPOSIT_WEIGHT) * vr;
    col[1] = RC(FSIM_DEPOSIT_WEIGHT) * va;
    col[2] = RC(FSIM_DEPOSIT_WEIGHT) * V[2];
    col[3] = RC(FSIM_DEPOSIT_WEIGHT) * RC(1.0);
    *xw = r * (REAL)nr;
    *yw = z * (REAL)nz;
    /* GLES2 clips a point whose centre is outside the clip volume; NaN too. */
    if (!(*xw >= RC(0.0)) || !(*xw < (REAL)nr)) return 0;
    if (!(*yw >= RC(0.0)) || !(*yw < (REAL)nz)) return 0;
    return 1;
}

/* LITERAL form: N point sprites of size 11, additive blend, in particle-index
 * order (empic.js:1473-1478, fragment shader :1022).  GLES 2.0 section 3.3: a
 * fragment is produced for each pixel whose centre lies in the size-11 square
 * about (xw,yw); gl_PointCoord.s = 1/2 + (xf + 1/2 - xw)/size.  Used only to
 * validate the convolution identity on small cases (O(121 N)).               */
void ORC(orc_deposit_sprites)(int64_t n, const REAL *pos, const REAL *vel,
                              const REAL *shape /* 11x11 reals */, int64_t nr,
                              int64_t nz, REAL *mom /* nr*nz*4, cleared here */)
{
    const REAL size = RC(FSIM_NSHAPE);
    for (int64_t k = 0; k < 4 * nr * nz; ++k) mom[k] = RC(0.0);
    for (int64_t p = 0; p < n; ++p) {
        REAL col[4], xw, yw;
        if (!ORC(sprite)(pos + 4 * p, vel + 4 * p, nr, nz, col, &xw, &yw)) continue;
        int64_t x0 = (int64_t)ORC(orc_floor)(xw - size * RC(0.5)) - 1;
        int64_t y0 = (int64_t)ORC(orc_floor)(yw - size * RC(0.5)) - 1;
        for (int64_t yf = y0; yf <= y0 + FSIM_NSHAPE + 2; ++yf) {
            REAL cy = (REAL)yf + RC(0.5);
            if (!(cy >= yw - size * RC(0.5) && cy < yw + size * RC(0.5))) continue;
            if (yf < 0 || yf >= nz) continue;
            REAL t = RC(0.5) + (cy - yw) / size;
            int64_t tj = ORC(tex)(t, FSIM_NSHAPE);
            for (int64_t xf = x0; xf <= x0 + FSIM_NSHAPE + 2; ++xf) {
                REAL cx = (REAL)xf + RC(0.5);
                if (!(cx >= xw - size * RC(0.5) && cx < xw + size * RC(0.5))) continue;
                if (xf < 0 || xf >= nr) continue;
                REAL s = RC(0.5) + (cx - xw) / size;
                int64_t ti = ORC(tex)(s, FSIM_NSHAPE);
                REAL w = shape[ti + FSIM_NSHAPE * tj];
                REAL *m = mom + 4 * (xf + yf * nr);
                m[0] += col[0] * w; m[1] += col[1] * w;
                m[2] += col[2] * w; m[3] += col[3] * w;
            }
        }
    }
}

/* CANONICAL form, step 1: per-cell nearest-grid-point sums of the sprite
 * colours, accumulated sequentially in ascending particle-index order (GL
 * primitive order).  count[] is the integer population of each cell.         */
void ORC(orc_cell_sums)(int64_t n, const REAL *pos, const REAL *vel, int64_t nr,
                        int64_t nz, REAL *S /* nr*nz*4 */, uint32_t *count,
                        int64_t *cell_out /* n, -1 = clipped, may be NULL */)
{
    for (int64_t k = 0; k < 4 * nr * nz; ++k) S[k] = RC(0.0);
    for (int64_t k = 0; k < nr * nz; ++k) count[k] = 0;
    for (int64_t p = 0; p < n; ++p) {
        REAL col[4], xw, yw;
        if (!ORC(sprite)(pos + 4 * p, vel + 4 * p, nr, nz, col, &xw, &yw)) {
            if (cell_out) cell_out[p] = -1;
            continue;
        }
        int64_t c = (int64_t)xw + nr * (int64_t)yw;
        if (cell_out) cell_out[p] = c;
        S[4 * c + 0] += col[0]; S[4 * c + 1] += col[1];
        S[4 * c + 2] += col[2]; S[4 * c + 3] += col[3];
        count[c] += 1;
    }
}

/* Timing-only variant of orc_cell_sums: particle-parallel with atomic adds
 * (sum order not fixed).  Used by the cpu_baseline / --impl reference legs.  */
void ORC(orc_cell_sums_mt)(int64_t n, const REAL *pos, const REAL *vel, int64_t nr,
                           int64_t nz, REAL *S, uint32_t *count, int nthreads)
{
    int64_t k, p;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (k = 0; k < nr * nz; ++k) {
        S[4 * k] = S[4 * k + 1] = S[4 * k + 2] = S[4 * k + 3] = RC(0.0);
        count[k] = 0;
    }
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (p = 0; p < n; ++p) {
        REAL col[4], xw, yw;
        if (!ORC(sprite)(pos + 4 * p, vel + 4 * p, nr, nz, col, &xw, &yw)) continue;
        int64_t c = (int64_t)xw + nr * (int64_t)yw;
        for (int q = 0; q < 4; ++q) {
#pragma omp atomic
            S[4 * c + q] += col[q];
        }
#pragma omp atomic
        count[c] += 1;
    }
}

/* CANONICAL form, step 2: moments01 = S (*) shape, gather form of the sprite scatter: pixel (i,j)
 * receives S[i+di, j+dj] * shape[5+di, 5+dj] for every offset with a non-zero weight (the 40 corner
 * texels with d > 5 are exactly 0 and are skipped); sources outside the grid do not exist (sprites
 * are clipped at the target edge).  The footprint is mirror-symmetric, shape[5+di,5+dj] =
 * shape[5-di,5+dj] = shape[5+di,5-dj] = shape[5-di,5-dj] bit for bit, so the (up to four) mirror
 * sources of a weight are added first and weighted once.  Fixed order: di = 0..5 outer, dj = 0..5
 * inner; inside a class the two sources of a ROW are added first, then the two rows:
 *     (S[-di,-dj] + S[+di,-dj]) + (S[-di,+dj] + S[+di,+dj]),   duplicates (di or dj = 0) once;
 * a source outside the grid counts as an exact zero (x + 0 = x).  The row pairs do not depend on the
 * output row, which is what lets the device kernel form them once per window and share them. */
static inline void ORC(conv_row_pair)(const REAL *S, int64_t nr, int64_t nz, int64_t i, int64_t jj, int di,
                                      REAL h[4])
{
    h[0] = h[1] = h[2] = h[3] = RC(0.0);
    if (g_periodic_z) jj = (jj % nz + nz) % nz; /* EXTENSION: the footprint wraps in z */
    if (jj < 0 || jj >= nz) return;
    if (i - di >= 0) {
        const REAL *s = S + 4 * ((i - di) + jj * nr);
        h[0] = s[0]; h[1] = s[1]; h[2] = s[2]; h[3] = s[3];
    }
    if (di) {
        REAL b[4] = {RC(0.0), RC(0.0), RC(0.0), RC(0.0)};
        if (i + di < nr) {
            const REAL *s = S + 4 * ((i + di) + jj * nr);
            b[0] = s[0]; b[1] = s[1]; b[2] = s[2]; b[3] = s[3];
        }
        h[0] = h[0] + b[0]; h[1] = h[1] + b[1]; h[2] = h[2] + b[2]; h[3] = h[3] + b[3];
    }
}

void ORC(orc_convolve)(int64_t nr, int64_t nz, const REAL *S, const REAL *shape,
                       REAL *mom, int nthreads)
{
    int64_t j;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (j = 0; j < nz; ++j)
        for (int64_t i = 0; i < nr; ++i) {
            REAL acc[4] = {RC(0.0), RC(0.0), RC(0.0), RC(0.0)};
            for (int di = 0; di <= FSIM_SHAPE_MID; ++di)
                for (int dj = 0; dj <= FSIM_SHAPE_MID; ++dj) {
                    REAL w = shape[(FSIM_SHAPE_MID + di) + FSIM_NSHAPE * (FSIM_SHAPE_MID + dj)];
                    if (w == RC(0.0)) continue;
                    REAL sum[4], hp[4];
                    ORC(conv_row_pair)(S, nr, nz, i, j - dj, di, sum);
                    if (dj) {
                        ORC(conv_row_pair)(S, nr, nz, i, j + dj, di, hp);
                        sum[0] = sum[0] + hp[0]; sum[1] = sum[1] + hp[1];
                        sum[2] = sum[2] + hp[2]; sum[3] = sum[3] + hp[3];
                    }
                    acc[0] = acc[0] + sum[0] * w; acc[1] = acc[1] + sum[1] * w;
                    acc[2] = acc[2] + sum[2] * w; acc[3] = acc[3] + sum[3] * w;
                }
            REAL *m = mom + 4 * (i + j * nr);
            m[0] = acc[0]; m[1] = acc[1]; m[2] = acc[2]; m[3] = acc[3];
        }
}

/* programNormalizeMoments01, empic.js:1053-1056, then programAvgMoments
 * (avg_frag empic.js:274-277, u_ratio :1083) and the avgA -> avgB copy
 * (:1490-1495).  avg holds avgB on entry and avgA (== new avgB) on exit.     */
void ORC(orc_normalize_ema)(int64_t nr, int64_t nz, const REAL *mom, REAL *norm,
                            REAL *avg, int nthreads)
{
    const REAL ratio = RC(FSIM_EMA_RATIO);
    int64_t c;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (c = 0; c < nr * nz; ++c) {
        int64_t i = c % nr;
        REAL u = ((REAL)i + RC(0.5)) / (REAL)nr;
        REAL M[4];
        REAL a = mom[4 * c + 3];
        if (a > RC(0.0)) {
            M[0] = mom[4 * c] / a; M[1] = mom[4 * c + 1] / a;
            M[2] = mom[4 * c + 2] / a; M[3] = a;
        } else {
            M[0] = M[1] = M[2] = M[3] = RC(0.0);
        }
        for (int q = 0; q < 4; ++q) {
            REAL v = RC(FSIM_NORM_SCALE) * M[q] * RC(FSIM_NORM_HALF) / u;
            norm[4 * c + q] = v;
            avg[4 * c + q] = ratio * v + (RC(1.0) - ratio) * avg[4 * c + q];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* programBMag (empic.js:479-482) drawn to the RGBA8 canvas, then
 * programDensity (:1101-1105) blended SRC_ALPHA,ONE (:1497-1504).  Fixed-point
 * target: each draw's colour is clamped to [0,1]; the stored value is
 * round(255*c).  NaN converts to 0.  Output rows are canvas rows (top row =
 * GL row nz-1).                                                              */
static inline REAL ORC(clamp01)(REAL v)
{
    if (!(v > RC(0.0))) return RC(0.0);
    if (v > RC(1.0)) return RC(1.0);
    return v;
}
static inline REAL ORC(quant8)(REAL v) /* v in [0,1] -> stored value k/255 */
{
    return ORC(orc_floor)(v * RC(255.0) + RC(0.5));
}
void ORC(orc_render_mt)(int64_t nr, int64_t nz, const REAL *B, const REAL *avg,
                        uint8_t *rgba, int nthreads)
{
    int64_t j;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (j = 0; j < nz; ++j)
        for (int64_t i = 0; i < nr; ++i) {
            int64_t c = i + j * nr;
            REAL Bx = B[4 * c], By = B[4 * c + 1], Bz = B[4 * c + 2];
            REAL mag = ORC(orc_sqrt)(Bx * Bx + By * By + Bz * Bz);
            REAL dx = Bx / mag, dz = Bz / mag;
            REAL mn = (dz < RC(0.0)) ? dz : RC(0.0);
            REAL mx = (dz > RC(0.0)) ? dz : RC(0.0);
            REAL c1[4];
            c1[0] = mag * ((mn < RC(0.0)) ? -mn : mn);
            c1[1] = mag * dx;
            c1[2] = mag * ((mx < RC(0.0)) ? -mx : mx);
            c1[3] = RC(1.0);
            REAL a = avg[4 * c + 3];
            REAL src[4] = {RC(FSIM_RENDER_DENSITY) * a, RC(FSIM_RENDER_DENSITY) * a,
                           RC(FSIM_RENDER_DENSITY) * a, RC(FSIM_RENDER_DENSITY) * RC(1.0)};
            REAL sa = ORC(clamp01)(src[3]);
            uint8_t *o = rgba + 4 * (i + (nz - 1 - j) * nr);
            for (int q = 0; q < 4; ++q) {
                REAL dst = ORC(quant8)(ORC(clamp01)(c1[q])) / RC(255.0);
                REAL out = ORC(clamp01)(src[q]) * sa + dst;
                o[q] = (uint8_t)ORC(quant8)(ORC(clamp01)(out));
            }
        }
}

void ORC(orc_render)(int64_t nr, int64_t nz, const REAL *B, const REAL *avg, uint8_t *rgba)
{
    ORC(orc_render_mt)(nr, nz, B, avg, rgba, 1);
}

#undef ORC_CAT2
#undef ORC_CAT
#undef ORC
#undef RC
