/* fsim_oracle_spindle_impl.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU statement of the spindle-cusp boundary solve ("next" row N3 of SURVEY.md section 8f, second half).
 * spindle.js does not run in the reference (undefined names, a shader that does not compile, unit slips:
 * SURVEY.md section 0 row 6), so there is nothing to be bit-equal to: this file IS the specification,
 * written from the intent of spindle.js:27-30 ("solves the boundary conditions for a perfect conductor in
 * center of a spindle cusp magnetic field") and :632-654 (solve A x = b with makeSORIterative, then superpose
 * the loops).  PARITY UNPINNED.  include/fusionsim.h (fsim_add_spindle_cusp_plasma_field) holds the same text;
 * the CUDA product must equal this file bit for bit.
 *
 * Included once (double precision assembly) and twice more for the REAL-typed field superposition.       */
#ifndef FSIM_ORACLE_SPINDLE_ONCE
#define FSIM_ORACLE_SPINDLE_ONCE

/* field (tesla per ampere) at (x, z) of a loop of radius Rl at height Zl: the quadrature of
 * programCurrentLoopShape (empic.js:308-326, spindle.js:296-318) at the exact relative position, in
 * metres, with the midpoint-rule weight 2 pi / 1000 where the reference has 0.001                    */
static void orcs_loop_field(double Rl, double Zl, double x, double z, const double *costab, double *br, double *bz)
{
    const double dz = z - Zl;
    const double K = Rl * FSIM_SPINDLE_QW * FSIM_MU0 / (4.0 * FSIM_PI_GLSL);
    double Br = 0.0, Bz = 0.0;
    for (int k = 0; k < FSIM_NQUAD; ++k) {
        const double c = costab[k];
        const double rho = sqrt(Rl * Rl + x * x + dz * dz - 2.0 * x * Rl * c);
        const double f = (rho > 0.0) ? K / (rho * rho * rho) : 0.0;
        Br += dz * f * c;
        Bz += f * (Rl - x * c);
    }
    *br = Br;
    *bz = Bz;
}

/* surface geometry, spindle.js:138-176 (lower half): nodes [L+1][2], collocation points [L][2], normals [L][2] */
void orcs_geometry(double radius, double height, double *nodes, double *points, double *normals)
{
    const int L = FSIM_SPINDLE_L;
    const double a = FSIM_SPINDLE_A;
    const double s = height / (2.0 * radius);
    const double R = radius * sqrt(1.0 + a * a);
    const double alpha = atan(a);
    const double theta = alpha + FSIM_PI;
    const double arc = 0.5 * FSIM_PI - 2.0 * alpha;
    for (int l = 0; l <= L; ++l) {
        const double phi = (double)l * arc / (double)L + theta;
        nodes[2 * l] = (l == 0) ? 0.0 : R * cos(-phi) + radius; /* node 0 lies on the axis */
        nodes[2 * l + 1] = s * (R * sin(-phi));
    }
    for (int p = 0; p < L; ++p) {
        const double phi = ((double)p + 0.5) * arc / (double)L + theta;
        points[2 * p] = R * cos(-phi) + radius;
        points[2 * p + 1] = s * (R * sin(-phi));
        const double nx = -(s * cos(-phi)), nz = -sin(-phi);
        const double len = sqrt(nx * nx + nz * nz);
        normals[2 * p] = nx / len;
        normals[2 * p + 1] = nz / len;
    }
}

/* A [L][L] row-major (row = collocation point), b [L]: the system A x = -b with the gauge x_{L-1} = 0 already
 * applied (row and column L-1 are those of the identity, rhs[L-1] = 0); rhs = -b is returned in `rhs`.   */
void orcs_assemble(double radius, double height, double coil_r, double coil_I, const double *costab,
                   const double *nodes, const double *points, const double *normals, double *A, double *rhs,
                   int nthreads)
{
    const int L = FSIM_SPINDLE_L;
    double *nf = (double *)malloc(sizeof(double) * 2 * (size_t)(L + 1) * L); /* [node][point][r,z] */
    int l;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (l = 0; l <= L; ++l)
        for (int p = 0; p < L; ++p) {
            double r0, z0, r1, z1;
            orcs_loop_field(nodes[2 * l], nodes[2 * l + 1], points[2 * p], points[2 * p + 1], costab, &r0, &z0);
            orcs_loop_field(nodes[2 * l], height - nodes[2 * l + 1], points[2 * p], points[2 * p + 1], costab, &r1, &z1);
            nf[2 * ((size_t)l * L + p)] = r0 - r1;       /* the mirror image carries the opposite current */
            nf[2 * ((size_t)l * L + p) + 1] = z0 - z1;
        }
    for (int p = 0; p < L; ++p) {
        const double nx = normals[2 * p], nz = normals[2 * p + 1];
        for (int e = 0; e < L; ++e) {
            const double er = nf[2 * ((size_t)e * L + p)] - nf[2 * ((size_t)(e + 1) * L + p)];
            const double ez = nf[2 * ((size_t)e * L + p) + 1] - nf[2 * ((size_t)(e + 1) * L + p) + 1];
            A[(size_t)p * L + e] = nx * er + nz * ez;
        }
        double r0, z0, r1, z1;
        orcs_loop_field(coil_r, 0.0, points[2 * p], points[2 * p + 1], costab, &r0, &z0);
        orcs_loop_field(coil_r, height, points[2 * p], points[2 * p + 1], costab, &r1, &z1);
        rhs[p] = -(coil_I * (nx * (r0 - r1) + nz * (z0 - z1)));
    }
    for (int k = 0; k < L; ++k) {
        A[(size_t)(L - 1) * L + k] = 0.0;
        A[(size_t)k * L + (L - 1)] = 0.0;
    }
    A[(size_t)(L - 1) * L + (L - 1)] = 1.0;
    rhs[L - 1] = 0.0;
    free(nf);
}

/* element strengths x [L] -> node currents [L+1], scaled by 1 - sqrt(1 - beta) */
void orcs_node_currents(const double *x, double beta, double *cur)
{
    const int L = FSIM_SPINDLE_L;
    const double scale = 1.0 - sqrt(1.0 - beta);
    for (int l = 0; l <= L; ++l) {
        const double plus = (l < L) ? x[l] : 0.0, minus = (l > 0) ? x[l - 1] : 0.0;
        cur[l] = scale * (plus - minus);
    }
}
#endif /* FSIM_ORACLE_SPINDLE_ONCE */

#ifdef REAL
#define ORCS_CAT2(a, b) a##_##b
#define ORCS_CAT(a, b) ORCS_CAT2(a, b)
#define ORCS(name) ORCS_CAT(name, SFX)

/* B += sum over loops of I_l * (loop field at the cell centre), in the working precision; loops [n][3] = R, Z, I
 * in metres / amperes (doubles, rounded to REAL); B is the RGBA texture [nr*nz][4] (components r, theta, z).  */
void ORCS(orcs_add_loops)(int64_t nr, int64_t nz, double radius, double height, int64_t nloops, const double *loops,
                          const REAL *costab, REAL *B, int nthreads)
{
    const REAL dr = (REAL)(radius / (double)nr), dzc = (REAL)(height / (double)nz);
    int64_t c;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (c = 0; c < nr * nz; ++c) {
        const REAL x = ((REAL)(c % nr) + (REAL)0.5) * dr;
        const REAL z = ((REAL)(c / nr) + (REAL)0.5) * dzc;
        REAL sr = (REAL)0.0, sz = (REAL)0.0;
        for (int64_t l = 0; l < nloops; ++l) {
            const REAL Rl = (REAL)loops[3 * l], Zl = (REAL)loops[3 * l + 1], I = (REAL)loops[3 * l + 2];
            const REAL dz = z - Zl;
            const REAL K = Rl * (REAL)FSIM_SPINDLE_QW * (REAL)FSIM_MU0 / ((REAL)4.0 * (REAL)FSIM_PI_GLSL);
            REAL Br = (REAL)0.0, Bz = (REAL)0.0;
            for (int k = 0; k < FSIM_NQUAD; ++k) {
                const REAL cs = costab[k];
                const REAL rho = ORC_SQRT(Rl * Rl + x * x + dz * dz - (REAL)2.0 * x * Rl * cs);
                const REAL f = (rho > (REAL)0.0) ? K / (rho * rho * rho) : (REAL)0.0;
                Br += dz * f * cs;
                Bz += f * (Rl - x * cs);
            }
            sr += I * Br;
            sz += I * Bz;
        }
        B[4 * c] = B[4 * c] + sr;
        B[4 * c + 2] = B[4 * c + 2] + sz;
    }
}
#undef ORCS_CAT2
#undef ORCS_CAT
#undef ORCS
#endif
