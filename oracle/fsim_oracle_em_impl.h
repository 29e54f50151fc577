/* fsim_oracle_em_impl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, precision-generic body).
 *
 * EXTENSION, NO REFERENCE COUNTERPART (BASELINE.json configs[2] names "Boris + Yee
 * FDTD"): the reference's E and B are static.  This file is the written specification of the electromagnetic field
 * update the B200 build adds -- an axisymmetric Yee scheme in the reference's cylindrical (r, z) geometry, driven by
 * the current the deposit already measures (moments 1 of programMoments01, empic.js:1006).  PARITY UNPINNED by
 * construction; checked against the analytic TM010 resonance of the empty cavity and against the cold-plasma shift
 * of that resonance, w^2 = w_c^2 + w_p^2 (tests/test_em.py).
 *
 * Mesh (cell (i,j) = [i dr, (i+1) dr] x [j dz, (j+1) dz], the texel of empic.js:1162):
 *   E_r (i+1/2, j)   [nz+1][nr]      E_z (i, j+1/2)   [nz][nr+1]     B_t (i+1/2, j+1/2) [nz][nr]    (TM set)
 *   E_t (i, j)       [nz+1][nr+1]    B_r (i, j+1/2)   [nz][nr+1]     B_z (i+1/2, j)     [nz+1][nr]  (TE set)
 * all row-major [j][i].  Perfectly conducting wall at r = radius and end plates at z = 0, height: tangential E = 0
 * there (E_z at i = nr, E_t at i = nr and at j = 0, nz, E_r at j = 0, nz are never updated and stay 0); E_t = 0 on the
 * axis.  One step of dt (B then E, leap-frog), every product formed as written, left to right, no fused multiply-add;
 * per-column coefficients in host fp64, rounded to REAL:
 *   kz = dt/dz, kr = dt/dr, cz = c^2 dt/dz, cr = c^2 dt/dr, cj = dt/eps0, ax = 4 c^2 dt/dr,
 *   a1_i = dt (i+1) / ((i+1/2) dr), a0_i = dt i / ((i+1/2) dr)            (B_z: -(1/r) d(r E_t)/dr)
 *   b1_i = c^2 dt (i+1/2) / (i dr), b0_i = c^2 dt (i-1/2) / (i dr), i >= 1 (E_z:  (c^2/r) d(r B_t)/dr)
 *   B_r += kz (E_t[j+1][i] - E_t[j][i])
 *   B_t -= (kz (E_r[j+1][i] - E_r[j][i])) - (kr (E_z[j][i+1] - E_z[j][i]))
 *   B_z -= (a1_i E_t[j][i+1]) - (a0_i E_t[j][i])
 *   E_r += (-(cz (B_t[j][i] - B_t[j-1][i]))) - cj (0.5 (J_r[j-1][i] + J_r[j][i]))                    1 <= j <= nz-1
 *   E_t += ((cz (B_r[j][i] - B_r[j-1][i])) - (cr (B_z[j][i] - B_z[j][i-1])))
 *          - cj (0.25 (((J_t[j-1][i-1] + J_t[j-1][i]) + J_t[j][i-1]) + J_t[j][i]))                  1 <= i <= nr-1, 1 <= j <= nz-1
 *   E_z += ((b1_i B_t[j][i]) - (b0_i B_t[j][i-1])) - cj (0.5 (J_z[j][i-1] + J_z[j][i]))             1 <= i <= nr-1
 *   E_z[j][0] += (ax B_t[j][0]) - cj J_z[j][0]                                                      (axis)
 * current density at the cell centres from the deposited moments (RGBA texture moments01: sum over the sprites of
 * 0.001 v S, v in the reference's normalised units v/c * (1/radius, 1/radius, 1/height)):
 *   J_q[j][i] = g_q,i mom_q[j][i],  g_r,i = g_t,i = G_i radius, g_z,i = G_i height,
 *   G_i = q macro_weight 1000 c / (2 pi u_i radius dr dz), u_i = (i + 1/2)/nr   (cell volume 2 pi r dr dz)
 * fields the push gathers (cell centres), B0 = the static field present at fsim_em_init:
 *   E = (0.5 (E_r[j][i] + E_r[j+1][i]),  0.25 (((E_t[j][i] + E_t[j][i+1]) + E_t[j+1][i]) + E_t[j+1][i+1]),  0.5 (E_z[j][i] + E_z[j][i+1]))
 *   B = B0 + (0.5 (B_r[j][i] + B_r[j][i+1]),  B_t[j][i],  0.5 (B_z[j][i] + B_z[j+1][i]))
 */
#define ORCE_CAT2(a, b) a##_##b
#define ORCE_CAT(a, b) ORCE_CAT2(a, b)
#define ORCE(name) ORCE_CAT(name, SFX)
#define RC(x) ((REAL)(x))

/* coef [nr+1][6] (host doubles): a1 a0 b1 b0 gR gZ per column; scal [6]: kz kr cz cr cj ax */
void ORCE(orce_step)(int64_t nr, int64_t nz, REAL *Er, REAL *Ez, REAL *Bt, REAL *Et, REAL *Br, REAL *Bz, const double *coef_d,
                     const double *scal_d, const REAL *mom /* [nz*nr][4] or NULL */, int nthreads)
{
    const REAL kz = (REAL)scal_d[0], kr = (REAL)scal_d[1], cz = (REAL)scal_d[2], cr = (REAL)scal_d[3], cj = (REAL)scal_d[4],
               ax = (REAL)scal_d[5];
    REAL *c = (REAL *)malloc(sizeof(REAL) * 6 * (size_t)(nr + 1));
    for (int64_t k = 0; k < 6 * (nr + 1); ++k) c[k] = (REAL)coef_d[k];
    const int64_t P = nr + 1; /* row length of the arrays with nr+1 columns */
    int64_t j;
    /* ---- B from E ---- */
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (j = 0; j <= nz; ++j)
        for (int64_t i = 0; i <= nr; ++i) {
            if (j < nz) Br[j * P + i] = Br[j * P + i] + kz * (Et[(j + 1) * P + i] - Et[j * P + i]);
            if (j < nz && i < nr)
                Bt[j * nr + i] = Bt[j * nr + i] - ((kz * (Er[(j + 1) * nr + i] - Er[j * nr + i])) - (kr * (Ez[j * P + i + 1] - Ez[j * P + i])));
            if (i < nr) Bz[j * nr + i] = Bz[j * nr + i] - ((c[6 * i] * Et[j * P + i + 1]) - (c[6 * i + 1] * Et[j * P + i]));
        }
    /* ---- E from B and J ---- */
#define JQ(q, jj, ii) (mom ? c[6 * (ii) + ((q) == 2 ? 5 : 4)] * mom[4 * ((jj) * nr + (ii)) + (q)] : RC(0.0))
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (j = 0; j <= nz; ++j)
        for (int64_t i = 0; i <= nr; ++i) {
            if (i < nr && j >= 1 && j <= nz - 1)
                Er[j * nr + i] = Er[j * nr + i] + ((-(cz * (Bt[j * nr + i] - Bt[(j - 1) * nr + i]))) - cj * (RC(0.5) * (JQ(0, j - 1, i) + JQ(0, j, i))));
            if (i >= 1 && i <= nr - 1 && j >= 1 && j <= nz - 1)
                Et[j * P + i] = Et[j * P + i] + (((cz * (Br[j * P + i] - Br[(j - 1) * P + i])) - (cr * (Bz[j * nr + i] - Bz[j * nr + i - 1]))) -
                                                 cj * (RC(0.25) * (((JQ(1, j - 1, i - 1) + JQ(1, j - 1, i)) + JQ(1, j, i - 1)) + JQ(1, j, i))));
            if (j < nz && i >= 1 && i <= nr - 1)
                Ez[j * P + i] = Ez[j * P + i] + (((c[6 * i + 2] * Bt[j * nr + i]) - (c[6 * i + 3] * Bt[j * nr + i - 1])) -
                                                 cj * (RC(0.5) * (JQ(2, j, i - 1) + JQ(2, j, i))));
            if (j < nz && i == 0) Ez[j * P] = Ez[j * P] + ((ax * Bt[j * nr]) - cj * JQ(2, j, 0));
        }
#undef JQ
    free(c);
}

/* cell-centred E and B = B0 + B_em as RGBA textures [nz*nr][4] (w = 1) */
void ORCE(orce_cells)(int64_t nr, int64_t nz, const REAL *Er, const REAL *Ez, const REAL *Bt, const REAL *Et, const REAL *Br,
                      const REAL *Bz, const REAL *B0, REAL *E, REAL *B)
{
    const int64_t P = nr + 1;
    for (int64_t j = 0; j < nz; ++j)
        for (int64_t i = 0; i < nr; ++i) {
            const int64_t c = i + j * nr;
            E[4 * c] = RC(0.5) * (Er[j * nr + i] + Er[(j + 1) * nr + i]);
            E[4 * c + 1] = RC(0.25) * (((Et[j * P + i] + Et[j * P + i + 1]) + Et[(j + 1) * P + i]) + Et[(j + 1) * P + i + 1]);
            E[4 * c + 2] = RC(0.5) * (Ez[j * P + i] + Ez[j * P + i + 1]);
            E[4 * c + 3] = RC(1.0);
            B[4 * c] = B0[4 * c] + RC(0.5) * (Br[j * P + i] + Br[j * P + i + 1]);
            B[4 * c + 1] = B0[4 * c + 1] + Bt[j * nr + i];
            B[4 * c + 2] = B0[4 * c + 2] + RC(0.5) * (Bz[j * nr + i] + Bz[(j + 1) * nr + i]);
            B[4 * c + 3] = RC(1.0);
        }
}

#undef ORCE_CAT2
#undef ORCE_CAT
#undef ORCE
#undef RC
