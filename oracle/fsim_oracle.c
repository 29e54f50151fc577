/* fsim_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the particle-step path of kcdodd/fusion-sim: a restatement of
 * the reference's GLSL shaders and host-side set() code in plain C.
 * Pinned to the reference's own shader source, executed by oracle/glsl_interp.py
 * (tests/test_reference_glsl.py); what stays unpinned is listed in the header of
 * fsim_oracle_impl.h and in DESIGN.md section 2.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * Callers: tests/, __graft_entry__.smoke(), bench.py cpu_baseline and
 * --impl reference.  Never the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/fsim_constants.h"

/* EXTENSION: periodic boundary in z.  The reference has none (CLAMP_TO_EDGE textures,
 * absorbing sink mask rows, utilities.js:530-531, fusionsim.js:109-112).  When set, (1) a pushed position wraps,
 * z <- z - floor(z) (and exactly 1 -> 0), before the sink lookup; (2) the 11x11 footprint wraps in z; (3) the
 * Poisson stencil and the gradient wrap in z.  Set by OraclePusher before each call (spec.periodic_z). */
static int g_periodic_z = 0;
void orc_set_periodic_z(int on) { g_periodic_z = on; }

static inline double orc_sqrt_f64(double x) { return sqrt(x); }
static inline float orc_sqrt_f32(float x) { return sqrtf(x); }
static inline double orc_floor_f64(double x) { return floor(x); }
static inline float orc_floor_f32(float x) { return floorf(x); }

#define REAL double
#define SFX f64
#include "fsim_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL float
#define SFX f32
#include "fsim_oracle_impl.h"
#undef REAL
#undef SFX

/* dense weighted-Jacobi solver of matrix_webgl.js ("next" row N3 of SURVEY.md section 8f) */
#define REAL double
#define SFX f64
#include "fsim_oracle_jacobi_impl.h"
#undef REAL
#undef SFX
#define REAL float
#define SFX f32
#include "fsim_oracle_jacobi_impl.h"
#undef REAL
#undef SFX

/* EXTENSION: self-consistent electrostatic field solve ("next" row N4 of SURVEY.md section 8f) */
#define REAL double
#define SFX f64
#include "fsim_oracle_fields_impl.h"
#undef REAL
#undef SFX
#define REAL float
#define SFX f32
#include "fsim_oracle_fields_impl.h"
#undef REAL
#undef SFX

/* "next" row N3, second half: spindle-cusp boundary solve (specification written from the intent of spindle.js) */
#define REAL double
#define SFX f64
#define ORC_SQRT sqrt
#include "fsim_oracle_spindle_impl.h"
#undef ORC_SQRT
#undef REAL
#undef SFX
#define REAL float
#define SFX f32
#define ORC_SQRT sqrtf
#include "fsim_oracle_spindle_impl.h"
#undef ORC_SQRT
#undef REAL
#undef SFX

/* EXTENSION: axisymmetric Yee update driven by the deposited current (BASELINE configs[2]) */
#define REAL double
#define SFX f64
#include "fsim_oracle_em_impl.h"
#undef REAL
#undef SFX
#define REAL float
#define SFX f32
#include "fsim_oracle_em_impl.h"
#undef REAL
#undef SFX

/* coefficients of the Yee update (header of fsim_oracle_em_impl.h): coef [nr+1][6] = a1 a0 b1 b0 gR gZ, scal [6] =
 * kz kr cz cr cj ax; host doubles */
void orce_coeffs(int64_t nr, int64_t nz, double radius, double height, double dt, double charge, double macro_weight,
                 double *coef, double *scal)
{
    const double dr = radius / (double)nr, dz = height / (double)nz, c2 = FSIM_C_LIGHT * FSIM_C_LIGHT;
    scal[0] = dt / dz; scal[1] = dt / dr; scal[2] = c2 * dt / dz; scal[3] = c2 * dt / dr; scal[4] = dt / FSIM_EPS0;
    scal[5] = 4.0 * c2 * dt / dr;
    for (int64_t i = 0; i <= nr; ++i) {
        const double rh = ((double)i + 0.5) * dr;
        coef[6 * i] = dt * ((double)i + 1.0) * dr / (rh * dr);
        coef[6 * i + 1] = dt * (double)i * dr / (rh * dr);
        coef[6 * i + 2] = i ? c2 * dt * rh / ((double)i * dr * dr) : 0.0;
        coef[6 * i + 3] = i ? c2 * dt * (((double)i - 0.5) * dr) / ((double)i * dr * dr) : 0.0;
        const double u = ((double)i + 0.5) / (double)nr;
        const double G = charge * macro_weight * 1000.0 * FSIM_C_LIGHT / (2.0 * FSIM_PI * u * radius * dr * dz);
        coef[6 * i + 4] = G * radius;
        coef[6 * i + 5] = G * height;
    }
}

/* per-column Jacobi coefficients [nr][4] = cE cW cZ cB (header of fsim_oracle_fields_impl.h) */
void orc_relax_coeffs(int64_t nr, double dr, double dz, double *out)
{
    for (int64_t i = 0; i < nr; ++i) {
        const double rc = ((double)i + 0.5) * dr * dr;
        const double aE = ((double)i + 1.0) / rc, aW = (double)i / rc, aZ = 1.0 / (dz * dz);
        const double aC = aE + aW + 2.0 * aZ;
        out[4 * i] = aE / aC;
        out[4 * i + 1] = aW / aC;
        out[4 * i + 2] = aZ / aC;
        out[4 * i + 3] = 1.0 / aC;
    }
}

/* N(num) = num.toFixed(20) (empic.js:23-25) parsed back by the GLSL compiler. */
double orc_tofixed20(double x)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%.20f", x);
    return strtod(buf, NULL);
}

/* cos(PI*(k+0.5)/1000) of empic.js:317 (the GLSL literal 3.14159265359), one value per quadrature
 * point.  The ARGUMENT is formed in the shader's working precision -- float arithmetic when
 * as_f32 != 0, exactly as the GLSL expression is written -- and the cosine of that argument is the
 * host libm's (GLSL ES 1.00 gives cos no accuracy bound).                                     */
void orc_cos_table(double *out, int as_f32)
{
    for (int k = 0; k < FSIM_NQUAD; ++k) {
        if (as_f32) {
            const float a = (float)FSIM_PI_GLSL * ((float)k + 0.5f) / 1000.0f;
            out[k] = (double)(float)cos((double)a);
        } else {
            out[k] = cos(FSIM_PI_GLSL * ((double)k + 0.5) / 1000.0);
        }
    }
}

/* Deposit footprint, empic.js:949-971.  as_f32 != 0 reproduces the
 * Float32Array round trips of the reference (store, re-read, divide, store). */
void orc_shape_table(double *out /* 121 */, int as_f32)
{
    const int n = FSIM_NSHAPE;
    const double mid = (n - 1) / 2.0;
    double sum = 0.0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            double d = sqrt(pow(i - mid, 2) + pow(j - mid, 2));
            double c = cos(0.5 * M_PI * d / mid);
            double v = pow(c > 0.0 ? c : 0.0, 2);
            if (as_f32) v = (double)(float)v;
            out[i + n * j] = v;
            sum += v;
        }
    for (int k = 0; k < n * n; ++k) {
        double v = out[k] / sum;
        out[k] = as_f32 ? (double)(float)v : v;
    }
}

/* source_pdf -> 512x512 inverse-cdf table, empic.js:1268-1339, host doubles.
 * pdf is [n0][n1] row-major (pdf[i][j], i along r).  out is [512*512][2]
 * with texel (i,j) at 2*(i + j*512): (x, y).  NaN behaviour of the JS code is
 * kept (0/0 rows, comparisons with NaN are false).                           */
int orc_inv_cdf(const double *pdf, int64_t n0, int64_t n1, double *out)
{
    double *cdf_y = (double *)malloc(sizeof(double) * (size_t)(n0 * n1));
    double *cdf_x = (double *)malloc(sizeof(double) * (size_t)n0);
    if (!cdf_y || !cdf_x) return -1;
    double sum_x = 0.0;
    for (int64_t i = 0; i < n0; ++i) {
        double sum_y = 0.0;
        for (int64_t j = 0; j < n1; ++j) {
            sum_y += pdf[i * n1 + j];
            cdf_y[i * n1 + j] = sum_y;
        }
        for (int64_t j = 0; j < n1; ++j) cdf_y[i * n1 + j] /= sum_y;
        sum_x += sum_y;
        cdf_x[i] = sum_x;
    }
    for (int64_t i = 0; i < n0; ++i) cdf_x[i] /= sum_x;

    const int T = FSIM_N_INVCDF;
    for (int ti = 0; ti < T; ++ti) {
        double f1 = (double)ti / 511.0;
        /* inverse_cdf_x(f1), empic.js:1293-1309 */
        double x;
        {
            int64_t i = 0;
            while (i < n0 && cdf_x[i] < f1) i++;
            if (i >= n0) { x = NAN; }
            else if (i == 0) x = (f1 / cdf_x[0]) / (double)n0;
            else x = ((double)i + (f1 - cdf_x[i - 1]) / (cdf_x[i] - cdf_x[i - 1])) / (double)n0;
        }
        for (int tj = 0; tj < T; ++tj) {
            double f2 = (double)tj / 511.0;
            /* inverse_cdf_y(x, f2), empic.js:1311-1326 */
            double y;
            {
                double fi = floor(x * (double)n0);
                int64_t i = (fi < (double)(n0 - 1)) ? (int64_t)fi : n0 - 1;
                if (!(fi == fi) || i < 0) { y = NAN; }
                else {
                    const double *cy = cdf_y + i * n1;
                    int64_t j = 0;
                    while (j < n1 && cy[j] < f2) j++;
                    if (j >= n1) y = NAN;
                    else if (j == 0) y = (f2 / cy[0]) / (double)n1;
                    else y = ((double)j + (f2 - cy[j - 1]) / (cy[j] - cy[j - 1])) / (double)n1;
                }
            }
            out[2 * (ti + tj * T) + 0] = x;
            out[2 * (ti + tj * T) + 1] = y;
        }
    }
    free(cdf_y);
    free(cdf_x);
    return 0;
}

int orc_abi_version(void) { return 1; }
