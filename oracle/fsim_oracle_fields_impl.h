/* fsim_oracle_fields_impl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, precision-generic body).
 *
 * EXTENSION, NO REFERENCE COUNTERPART: the reference never feeds the
 * deposited density back into E (empic.js:1471-1505 only draws it).  This file is the written
 * specification of the self-consistent electrostatic field solve the B200 build adds, in the
 * reference's own geometry (cylindrical r,z; cell (i,j) centred at r=(i+.5)dr, z=(j+.5)dz; texel
 * index i + j*nr, empic.js:1162) and with the reference's own solver: the weighted-Jacobi iteration
 * of matrix_webgl.makeSORIterative, x <- omega (R x + C) + (1 - omega) x  (matrix_webgl.js:224-300,
 * :392-420), applied to the sparse 5-point operator instead of a dense matrix.
 * PARITY UNPINNED by construction; checked against oracle/numpy_ref.py and a direct sparse solve.
 *
 *   div grad phi = -rho/eps0, finite volumes on cell-centred rings:
 *     [ r_{i+1/2} (phi_{i+1,j} - phi_{i,j}) - r_{i-1/2} (phi_{i,j} - phi_{i-1,j}) ] / (r_i dr^2)
 *       + (phi_{i,j+1} - 2 phi_{i,j} + phi_{i,j-1}) / dz^2 = -src_{i,j}
 *   r_{i-1/2} = i dr, so the axis needs no ghost value; ghost cells beyond r = radius, z = 0 and
 *   z = height hold phi = 0 (grounded wall).
 *   per-column coefficients (host fp64, then rounded to REAL):
 *     aE = (i+1)/((i+.5) dr^2), aW = i/((i+.5) dr^2), aZ = 1/dz^2, aC = aE + aW + 2 aZ,
 *     cE = aE/aC, cW = aW/aC, cZ = aZ/aC, cB = 1/aC
 *   one sweep, evaluated exactly in this order, no fused multiply-add:
 *     t    = ((cE phi_E + cW phi_W) + cZ (phi_N + phi_S)) + cB src
 *     phi' = omega t + (1 - omega) phi
 *   E = -grad phi by centred differences: E_r = -((phi_E - phi_W) inv2dr), phi_W := phi at i = 0;
 *   E_z = -((phi_N - phi_S) inv2dz); E_theta = 0.
 */

#define ORC_CAT2(a, b) a##_##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define ORC(name) ORC_CAT(name, SFX)
#define RC(x) ((REAL)(x))

/* src = rho/eps0 from a density texture (RGBA, channel 3 = weighted count / (2 u), empic.js:1055-1056) */
void ORC(orc_charge_source)(int64_t ncell, const REAL *dens, double rho_scale_d, REAL *src)
{
    const REAL k = (REAL)rho_scale_d;
    for (int64_t c = 0; c < ncell; ++c) src[c] = k * dens[4 * c + 3];
}
/* ... with a neutralising background given per cell in the units of the density texture (e.g. the density
 * of the same species at t = 0 = immobile ions): src = k (dens.a - background) */
void ORC(orc_charge_source_bg)(int64_t ncell, const REAL *dens, double rho_scale_d, const REAL *background, REAL *src)
{
    const REAL k = (REAL)rho_scale_d;
    for (int64_t c = 0; c < ncell; ++c) src[c] = k * (dens[4 * c + 3] - background[c]);
}

/* `sweeps` weighted-Jacobi sweeps; phi is updated in place (tmp is scratch of the same size) */
void ORC(orc_relax)(int64_t nr, int64_t nz, REAL *phi, REAL *tmp, const REAL *src, const double *coef_d,
                    double omega_d, int sweeps, int nthreads)
{
    const REAL om = (REAL)omega_d, one_m = RC(1.0) - om;
    REAL *coef = (REAL *)malloc(sizeof(REAL) * 4 * (size_t)nr);
    for (int64_t k = 0; k < 4 * nr; ++k) coef[k] = (REAL)coef_d[k];
    REAL *in = phi, *out = tmp;
    for (int s = 0; s < sweeps; ++s) {
        int64_t j;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (j = 0; j < nz; ++j)
            for (int64_t i = 0; i < nr; ++i) {
                const int64_t c = i + j * nr;
                const REAL pE = (i + 1 < nr) ? in[c + 1] : RC(0.0);
                const REAL pW = (i > 0) ? in[c - 1] : RC(0.0);
                REAL pN = (j + 1 < nz) ? in[c + nr] : RC(0.0);
                REAL pS = (j > 0) ? in[c - nr] : RC(0.0);
                if (g_periodic_z) { /* EXTENSION: periodic in z instead of the grounded end walls */
                    if (j + 1 >= nz) pN = in[i];
                    if (j == 0) pS = in[i + (nz - 1) * nr];
                }
                const REAL *k = coef + 4 * i;
                const REAL t = ((k[0] * pE + k[1] * pW) + k[2] * (pN + pS)) + k[3] * src[c];
                out[c] = om * t + one_m * in[c];
            }
        REAL *sw = in; in = out; out = sw;
    }
    if (in != phi) memcpy(phi, in, sizeof(REAL) * (size_t)(nr * nz));
    free(coef);
}

/* E texture (RGBA: E_r, E_theta, E_z, 1) from phi */
void ORC(orc_efield)(int64_t nr, int64_t nz, const REAL *phi, double inv2dr_d, double inv2dz_d, REAL *E)
{
    const REAL inv2dr = (REAL)inv2dr_d, inv2dz = (REAL)inv2dz_d;
    for (int64_t j = 0; j < nz; ++j)
        for (int64_t i = 0; i < nr; ++i) {
            const int64_t c = i + j * nr;
            const REAL pE = (i + 1 < nr) ? phi[c + 1] : RC(0.0);
            const REAL pW = (i > 0) ? phi[c - 1] : phi[c];
            REAL pN = (j + 1 < nz) ? phi[c + nr] : RC(0.0);
            REAL pS = (j > 0) ? phi[c - nr] : RC(0.0);
            if (g_periodic_z) {
                if (j + 1 >= nz) pN = phi[i];
                if (j == 0) pS = phi[i + (nz - 1) * nr];
            }
            E[4 * c] = -((pE - pW) * inv2dr);
            E[4 * c + 1] = RC(0.0);
            E[4 * c + 2] = -((pN - pS) * inv2dz);
            E[4 * c + 3] = RC(1.0);
        }
}

#undef ORC_CAT2
#undef ORC_CAT
#undef ORC
#undef RC
