/* fsim_oracle_jacobi_impl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, precision-generic body).
 *
 * CPU restatement of matrix_webgl.makeSORIterative (public/javascripts/matrix_webgl.js:35-711):
 * dense weighted Jacobi  x <- omega (R x + C) + (1-omega) x  on RGBA-packed textures.
 * Pinned to the reference's shader source: programR, programC, programMVproduct, the sum_frag
 * chain, programResult and programStats are executed from matrix_webgl.js by oracle/glsl_interp.py
 * (tests/golden/make_reference_vectors_jacobi.py) and the LITERAL mode below reproduces them bit
 * for bit (tests/test_reference_glsl.py) -- which also confirms defect (1) by execution.
 *
 * Packing restated from the shaders: vec_height vh = 2^n_power; the vector has L = 4 vh^2
 * entries, entry e = 4 (px + vh py) + channel (:107-127).  Row `row` of the iteration matrix is a
 * vh x vh tile of RGBA texels, column col = 4 (mx + vh my) + channel (programR :238-241).  A row
 * sum is: per texel the product with x (programMVproduct :323-327), then n_power passes that
 * add 2x2 texel blocks in the order (+x,+y), (-x,+y), (+x,-y), (-x,-y) (sum_frag :350), then the
 * four channels dot(v, vec4(1.0)) = ((r+g)+b)+a (programResult :408-411).
 *
 * Two reference defects are switchable (literal != 0 reproduces them, 0 = evident intent):
 *   (1) programResult gathers the row sums through  row = {2px, 2px+1, 2px+2vh, 2px+2vh+1} + 4 vh py
 *       (:408-411) while programR / programC number rows and entries e = 4 (px + vh py) + k
 *       (:238, :290): for vh >= 2 entry e receives the sum of a DIFFERENT matrix row.  Invisible in
 *       the only test the author left (a diagonal matrix, fusionsim.js:35-67).
 *   (2) solve() never resets x1, x2, x1x2, x1x1, x2x2 between iterations (:628-634, :675-682).
 * Sampling A at texel EDGES (c * vec2(col,row), :243-251) is taken as A[row][col].
 */
#define ORCJ_CAT2(a, b) a##_##b
#define ORCJ_CAT(a, b) ORCJ_CAT2(a, b)
#define ORCJ(name) ORCJ_CAT(name, SFX)
#define RC(x) ((REAL)(x))

/* programR (:238-254) and programC (:287-296): A is [L][L] row-major, R is [L][L] row-major in
 * NATURAL column order, C is [L].  omega_lit = N(omega) parsed back; omega == 1 skips the product. */
void ORCJ(orcj_setup)(int64_t L, const REAL *A, const REAL *b, double omega_d, int omega_is_one,
                      REAL *R, REAL *C, int nthreads)
{
    const REAL omega = (REAL)omega_d;
    int64_t row;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (row = 0; row < L; ++row) {
        const REAL d = A[row * L + row];
        for (int64_t col = 0; col < L; ++col) {
            REAL v = (row == col) ? RC(0.0) : -A[row * L + col] / d;
            R[row * L + col] = omega_is_one ? v : omega * v;
        }
        REAL c = b[row] / d;
        C[row] = omega_is_one ? c : omega * c;
    }
}

/* the row sum in the reference's order; prod is scratch of L reals */
static REAL ORCJ(row_sum)(int64_t vh, const REAL *Rrow, const REAL *x, REAL *prod)
{
    const int64_t L = 4 * vh * vh;
    for (int64_t c = 0; c < L; ++c) prod[c] = Rrow[c] * x[c];
    int64_t w = vh; /* current texel grid is w x w, texel (X,Y) channel k at prod[4 (X + w Y) + k] */
    while (w > 1) {
        const int64_t h = w / 2;
        for (int64_t Y = 0; Y < h; ++Y)
            for (int64_t X = 0; X < h; ++X)
                for (int k = 0; k < 4; ++k) {
                    REAL a = prod[4 * ((2 * X + 1) + w * (2 * Y + 1)) + k];
                    REAL bq = prod[4 * ((2 * X) + w * (2 * Y + 1)) + k];
                    REAL cq = prod[4 * ((2 * X + 1) + w * (2 * Y)) + k];
                    REAL dq = prod[4 * ((2 * X) + w * (2 * Y)) + k];
                    /* in-place is safe: target index 4 (X + h Y) + k <= every source index */
                    prod[4 * (X + h * Y) + k] = a + bq + cq + dq;
                }
        w = h;
    }
    return prod[0] * RC(1.0) + prod[1] * RC(1.0) + prod[2] * RC(1.0) + prod[3] * RC(1.0);
}

/* out.mv_product (:539-562): xnew = R x + C (+ (1-omega) x).  one_m_omega = N(1-omega). */
void ORCJ(orcj_mv_product)(int64_t vh, const REAL *R, const REAL *C, const REAL *x, REAL *xnew,
                           double one_m_omega_d, int omega_is_one, int literal, int nthreads)
{
    const int64_t L = 4 * vh * vh;
    const REAL omo = (REAL)one_m_omega_d;
#pragma omp parallel num_threads(nthreads)
    {
        REAL *prod = (REAL *)malloc(sizeof(REAL) * (size_t)L);
        int64_t e;
#pragma omp for schedule(static)
        for (e = 0; e < L; ++e) {
            int64_t row = e;
            if (literal) { /* defect (1): the gather of programResult :408-411 */
                const int64_t pix = e / 4, k = e % 4, px = pix % vh, py = pix / vh;
                row = 2 * px + 4 * vh * py + (k & 1) + ((k >> 1) ? 2 * vh : 0);
            }
            REAL s = ORCJ(row_sum)(vh, R + row * L, x, prod);
            REAL v = s + C[e];
            if (!omega_is_one) v = v + omo * x[e];
            xnew[e] = v;
        }
        free(prod);
    }
}

/* programStats (:443-449): per texel (dot(x1,x2)/4, dot(x1,x1)/4, dot(x2,x2)/4, max |x2-x1|) */
void ORCJ(orcj_stats)(int64_t npix, const REAL *x1, const REAL *x2, REAL *stats)
{
    for (int64_t i = 0; i < npix; ++i) {
        const REAL *a = x1 + 4 * i, *b = x2 + 4 * i;
        REAL d[4];
        for (int k = 0; k < 4; ++k) {
            REAL t = b[k] - a[k];
            d[k] = t < RC(0.0) ? -t : t;
        }
        stats[4 * i + 0] = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]) * RC(0.25);
        stats[4 * i + 1] = (a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]) * RC(0.25);
        stats[4 * i + 2] = (b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3]) * RC(0.25);
        REAL m = d[0] > d[1] ? d[0] : d[1];
        m = m > d[2] ? m : d[2];
        m = m > d[3] ? m : d[3];
        stats[4 * i + 3] = m;
    }
}

#undef ORCJ_CAT2
#undef ORCJ_CAT
#undef ORCJ
#undef RC
