// jacobi.cu -- dense weighted-Jacobi solver: the device routine behind matrix_webgl.makeSORIterative
// (public/javascripts/matrix_webgl.js:35-711; "next" row N3 of SURVEY.md section 8f).
//
//   x <- omega (R x + C) + (1 - omega) x,   R = -offdiag(A)/diag(A),  C = b/diag(A)
//
// The reference spreads one iteration over 3 + n_power render passes on RGBA-packed textures
// (programMVproduct :309, sum_programs :346-388, programResult :392).  Here one kernel does a
// whole mat-vec: one block per matrix row streams the row once (the kernel is HBM-bound: L*L
// reals per iteration), forms the products and adds them IN THE REFERENCE'S ORDER -- 2x2 texel
// blocks (+x,+y), (-x,+y), (+x,-y), (-x,-y) per channel, level by level, then ((r+g)+b)+a -- so
// the result is bit-identical to the CPU restatement.  The convergence statistics are computed per
// texel on the device (programStats :428) and reduced on the host in double, as solve() does.
//
// flags & FSIM_JACOBI_LITERAL reproduces two defects of the reference (see the oracle header):
// the row gather of programResult and the never-reset statistics accumulators.
#include <math.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

struct fsim_jacobi {
    int n_power = 0, vh = 0, prec = FSIM_F64, device = 0;
    unsigned flags = 0;
    size_t rs = 8;
    int64_t L = 0;
    double omega = 1.0, omega_lit = 1.0, omo_lit = 0.0;  // N(omega), N(1-omega) (:254, :413)
    bool omega_is_one = true;
    cudaStream_t stream = nullptr;
    void *A = nullptr, *R = nullptr, *b = nullptr, *C = nullptr, *xg = nullptr, *xr = nullptr, *stats = nullptr;
    void *stage = nullptr;
    bool have_A = false, have_b = false;
    int64_t launches = 0;
    double mv_ms = 0.0;
    int64_t mv_launches = 0;
};

namespace fsim {

static double tofixed20j(double x)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%.20f", x);
    return strtod(buf, nullptr);
}

template <typename F>
static int jdispatch(const fsim_jacobi *j, F &&f)
{
    if (j->prec == FSIM_F64) return f(double{});
    return f(float{});
}

// programR :238-254 and programC :287-296 in one pass
template <typename Real>
__global__ void __launch_bounds__(256)
jacobi_setup_kernel(const Real *__restrict__ A, const Real *__restrict__ b, Real *__restrict__ R,
                    Real *__restrict__ C, int64_t L, Real omega, int omega_is_one)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L * L) return;
    const int64_t row = t / L, col = t % L;
    const Real d = A[row * L + row];
    const Real v = (row == col) ? (Real)0.0 : -A[t] / d;
    R[t] = omega_is_one ? v : omega * v;
    if (col == 0) {
        const Real c = b[row] / d;
        C[row] = omega_is_one ? c : omega * c;
    }
}

// 8 consecutive reals: the matrix is streamed (evict-first), the vector stays cached
__device__ __forceinline__ void ld8_stream(const double *p, double (&o)[8])
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 t = __ldcs(reinterpret_cast<const double2 *>(p) + k);
        o[2 * k] = t.x; o[2 * k + 1] = t.y;
    }
}
__device__ __forceinline__ void ld8_stream(const float *p, float (&o)[8])
{
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(p) + k);
        o[4 * k] = t.x; o[4 * k + 1] = t.y; o[4 * k + 2] = t.z; o[4 * k + 3] = t.w;
    }
}
__device__ __forceinline__ void ld8(const double *p, double (&o)[8])
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p) + k);
        o[2 * k] = t.x; o[2 * k + 1] = t.y;
    }
}
__device__ __forceinline__ void ld8(const float *p, float (&o)[8])
{
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + k);
        o[4 * k] = t.x; o[4 * k + 1] = t.y; o[4 * k + 2] = t.z; o[4 * k + 3] = t.w;
    }
}

// one block per output entry e; see the header comment for the summation order
template <typename Real>
__global__ void __launch_bounds__(256)
jacobi_mv_kernel(const Real *__restrict__ R, const Real *__restrict__ C, const Real *__restrict__ x,
                 Real *__restrict__ xnew, int vh, Real omo, int omega_is_one, int literal)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Real *s0 = reinterpret_cast<Real *>(smem_raw);  // level buffers, ping-pong
    const int64_t L = 4ll * vh * vh;
    const int64_t e = blockIdx.x;
    int64_t row = e;
    if (literal) {  // programResult :408-411 gathers the row sums through a different numbering
        const int64_t pix = e / 4, k = e % 4, px = pix % vh, py = pix / vh;
        row = 2 * px + 4 * vh * py + (k & 1) + ((k >> 1) ? 2 * vh : 0);
    }
    const Real *Rrow = R + row * L;
    const int h = vh / 2;
    const int q1 = h * h;
    // level 1: products of four texels (2X..2X+1, 2Y..2Y+1), added (+x,+y), (-x,+y), (+x,-y), (-x,-y).
    // A thread reads the two 8-real runs (texels 2X and 2X+1 of rows 2Y and 2Y+1) with 128-bit loads;
    // the loop is unrolled so that several runs are in flight per thread (the kernel streams the
    // matrix row once: HBM-bound).
#pragma unroll 4
    for (int q = threadIdx.x; q < q1; q += blockDim.x) {
        const int X = q % h, Y = q / h;
        const size_t i00 = 4 * ((size_t)(2 * X) + (size_t)vh * (2 * Y));
        const size_t i01 = i00 + 4 * (size_t)vh;  // (2X, 2Y+1)
        Real r0[8], r1[8], x0[8], x1[8];
        ld8_stream(Rrow + i00, r0);
        ld8_stream(Rrow + i01, r1);
        ld8(x + i00, x0);
        ld8(x + i01, x1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const Real d = r0[k] * x0[k];
            const Real c = r0[4 + k] * x0[4 + k];
            const Real bq = r1[k] * x1[k];
            const Real a = r1[4 + k] * x1[4 + k];
            s0[4 * q + k] = a + bq + c + d;
        }
    }
    __syncthreads();
    Real *src = s0, *dst = s0 + 4 * q1;
    for (int w = h; w > 1; w >>= 1) {
        const int hh = w / 2;
        for (int q = threadIdx.x; q < hh * hh * 4; q += blockDim.x) {
            const int k = q & 3, X = (q >> 2) % hh, Y = (q >> 2) / hh;
            const Real a = src[4 * ((2 * X + 1) + w * (2 * Y + 1)) + k];
            const Real bq = src[4 * ((2 * X) + w * (2 * Y + 1)) + k];
            const Real c = src[4 * ((2 * X + 1) + w * (2 * Y)) + k];
            const Real d = src[4 * ((2 * X) + w * (2 * Y)) + k];
            dst[4 * (X + hh * Y) + k] = a + bq + c + d;
        }
        __syncthreads();
        Real *t = src; src = dst; dst = t;
    }
    if (threadIdx.x == 0) {
        const Real s = src[0] * (Real)1.0 + src[1] * (Real)1.0 + src[2] * (Real)1.0 + src[3] * (Real)1.0;
        Real v = s + C[e];
        if (!omega_is_one) v = v + omo * x[e];
        xnew[e] = v;
    }
}

// programStats :443-449
template <typename Real>
__global__ void __launch_bounds__(256)
jacobi_stats_kernel(const Real *__restrict__ x1, const Real *__restrict__ x2, Real *__restrict__ st, int64_t npix)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const Real *a = x1 + 4 * i, *b = x2 + 4 * i;
    Real d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const Real t = b[k] - a[k];
        d[k] = t < (Real)0.0 ? -t : t;
    }
    st[4 * i + 0] = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]) * (Real)0.25;
    st[4 * i + 1] = (a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]) * (Real)0.25;
    st[4 * i + 2] = (b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3]) * (Real)0.25;
    Real m = d[0] > d[1] ? d[0] : d[1];
    m = m > d[2] ? m : d[2];
    m = m > d[3] ? m : d[3];
    st[4 * i + 3] = m;
}

template <typename Real>
__global__ void jconvert_in(const double *__restrict__ in, Real *__restrict__ out, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (Real)in[k];
}
template <typename Real>
__global__ void jconvert_out(const Real *__restrict__ in, double *__restrict__ out, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (double)in[k];
}

static int jfail(int code, const std::string &m)
{
    set_error(m);
    return code;
}

static int jupload(fsim_jacobi *j, const double *host, void *dst, int64_t n)
{
    if (!host) return jfail(FSIM_ERR_INVALID, "null array");
    // stage holds L*L doubles; chunked so that it also serves when Real == double
    FSIM_CUDA(cudaMemcpyAsync(j->stage, host, sizeof(double) * n, cudaMemcpyHostToDevice, j->stream));
    return jdispatch(j, [&](auto tag) {
        using Real = decltype(tag);
        jconvert_in<Real><<<grid_for(n, 256), 256, 0, j->stream>>>((const double *)j->stage, (Real *)dst, n);
        FSIM_CUDA(cudaGetLastError());
        j->launches++;
        return (int)FSIM_OK;
    });
}

static int jdownload(fsim_jacobi *j, const void *src, double *host, int64_t n)
{
    int rc = jdispatch(j, [&](auto tag) {
        using Real = decltype(tag);
        jconvert_out<Real><<<grid_for(n, 256), 256, 0, j->stream>>>((const Real *)src, (double *)j->stage, n);
        FSIM_CUDA(cudaGetLastError());
        j->launches++;
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    FSIM_CUDA(cudaMemcpyAsync(host, j->stage, sizeof(double) * n, cudaMemcpyDeviceToHost, j->stream));
    FSIM_CUDA(cudaStreamSynchronize(j->stream));
    return FSIM_OK;
}

static int jmv(fsim_jacobi *j, const void *x, void *xnew)
{
    return jdispatch(j, [&](auto tag) {
        using Real = decltype(tag);
        const int h = j->vh / 2;
        const int q1 = h * h;
        const int threads = std::max(32, std::min(256, (q1 + 31) / 32 * 32));
        const size_t smem = sizeof(Real) * 4 * ((size_t)q1 + (size_t)std::max(1, q1 / 4));
        if (smem > 48 * 1024)
            FSIM_CUDA(cudaFuncSetAttribute(jacobi_mv_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jacobi_mv_kernel<Real><<<(unsigned)j->L, threads, smem, j->stream>>>(
            (const Real *)j->R, (const Real *)j->C, (const Real *)x, (Real *)xnew, j->vh, (Real)j->omo_lit,
            j->omega_is_one ? 1 : 0, (j->flags & FSIM_JACOBI_LITERAL) ? 1 : 0);
        FSIM_CUDA(cudaGetLastError());
        j->launches++;
        return (int)FSIM_OK;
    });
}

}  // namespace fsim

using namespace fsim;

extern "C" {

int fsim_jacobi_create(int32_t n_power, double relaxation, int32_t precision, int32_t device, uint32_t flags,
                       fsim_jacobi **out)
{
    if (!out) return jfail(FSIM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_power < 1 || n_power > 7) return jfail(FSIM_ERR_INVALID, ".n_power <- must be in [1, 7]");
    if (!(relaxation == relaxation)) return jfail(FSIM_ERR_INVALID, ".relaxation <- must be a number");
    if (precision != FSIM_F64 && precision != FSIM_F32) return jfail(FSIM_ERR_INVALID, ".precision <- must be FSIM_F64 or FSIM_F32");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return jfail(FSIM_ERR_CUDA, "no CUDA device: libfusionsim has no CPU fallback");
    if (device < 0 || device >= ndev) return jfail(FSIM_ERR_INVALID, ".device <- no such CUDA device");
    fsim_jacobi *j = new fsim_jacobi();
    j->n_power = n_power;
    j->vh = 1 << n_power;
    j->L = 4ll * j->vh * j->vh;
    j->prec = precision;
    j->rs = precision == FSIM_F64 ? 8 : 4;
    j->device = device;
    j->flags = flags;
    j->omega = relaxation == 0.0 ? 1.0 : relaxation;  // spec.relaxation || 1.0 (:55)
    j->omega_is_one = j->omega == 1.0;
    j->omega_lit = tofixed20j(j->omega);
    j->omo_lit = tofixed20j(1.0 - j->omega);
    cudaSetDevice(device);
    const size_t LL = (size_t)j->L * (size_t)j->L;
    bool ok = cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&j->A, LL * j->rs) == cudaSuccess && cudaMalloc(&j->R, LL * j->rs) == cudaSuccess;
    ok = ok && cudaMalloc(&j->stage, LL * sizeof(double)) == cudaSuccess;
    void **vecs[] = {&j->b, &j->C, &j->xg, &j->xr, &j->stats};
    for (void **v : vecs) ok = ok && cudaMalloc(v, (size_t)j->L * j->rs) == cudaSuccess &&
                               cudaMemset(*v, 0, (size_t)j->L * j->rs) == cudaSuccess;
    if (!ok) {
        cudaError_t e = cudaGetLastError();
        fsim_jacobi_destroy(j);
        return cuda_fail(e, "fsim_jacobi_create allocation", __FILE__, __LINE__);
    }
    *out = j;
    return FSIM_OK;
}

int fsim_jacobi_destroy(fsim_jacobi *j)
{
    if (!j) return FSIM_OK;
    cudaSetDevice(j->device);
    if (j->stream) cudaStreamSynchronize(j->stream);
    void *p[] = {j->A, j->R, j->b, j->C, j->xg, j->xr, j->stats, j->stage};
    for (void *q : p) cudaFree(q);
    if (j->stream) cudaStreamDestroy(j->stream);
    delete j;
    return FSIM_OK;
}

int64_t fsim_jacobi_vec_length(const fsim_jacobi *j) { return j ? j->L : -1; }
int64_t fsim_jacobi_launch_count(const fsim_jacobi *j) { return j ? j->launches : -1; }

int fsim_jacobi_set_matrix(fsim_jacobi *j, const double *A)
{
    if (!j) return jfail(FSIM_ERR_INVALID, "null handle");
    FSIM_CUDA(cudaSetDevice(j->device));
    FSIM_TRY(jupload(j, A, j->A, j->L * j->L));
    j->have_A = true;
    return FSIM_OK;
}
int fsim_jacobi_set_b(fsim_jacobi *j, const double *b)
{
    if (!j) return jfail(FSIM_ERR_INVALID, "null handle");
    FSIM_CUDA(cudaSetDevice(j->device));
    FSIM_TRY(jupload(j, b, j->b, j->L));
    j->have_b = true;
    return FSIM_OK;
}
int fsim_jacobi_init_vector(fsim_jacobi *j, const double *x)
{
    if (!j) return jfail(FSIM_ERR_INVALID, "null handle");
    FSIM_CUDA(cudaSetDevice(j->device));
    return jupload(j, x, j->xr, j->L);  // init_vector renders into x_result (:515-522)
}
int fsim_jacobi_get_result(fsim_jacobi *j, double *x)
{
    if (!j || !x) return jfail(FSIM_ERR_INVALID, "null argument");
    FSIM_CUDA(cudaSetDevice(j->device));
    return jdownload(j, j->xr, x, j->L);
}

// out.solve(params), matrix_webgl.js:576-699
int fsim_jacobi_solve(fsim_jacobi *j, double tolerance, int32_t substep, int32_t max_iterations,
                      double *correlation, double *diff_out, int32_t *iterations, double *result)
{
    if (!j) return jfail(FSIM_ERR_INVALID, "null handle");
    if (!j->have_A || !j->have_b) return jfail(FSIM_ERR_STATE, "solve: set_matrix and set_b first");
    if (!(tolerance == tolerance)) return jfail(FSIM_ERR_INVALID, ".tolerance <- Non-optional property is undefined!");
    FSIM_CUDA(cudaSetDevice(j->device));
    const int64_t L = j->L, npix = L / 4;
    int rc = jdispatch(j, [&](auto tag) {
        using Real = decltype(tag);
        jacobi_setup_kernel<Real><<<grid_for(L * L, 256), 256, 0, j->stream>>>(
            (const Real *)j->A, (const Real *)j->b, (Real *)j->R, (Real *)j->C, L, (Real)j->omega_lit,
            j->omega_is_one ? 1 : 0);
        FSIM_CUDA(cudaGetLastError());
        j->launches++;
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    std::vector<double> st((size_t)L), x1a((size_t)L), x2a((size_t)L);
    double corr = 0.0, x1 = 0, x2 = 0, x1x2 = 0, x1x1 = 0, x2x2 = 0;
    double diff = tolerance + 1;
    int it = 0;
    const int sub_n = substep > 0 ? substep : 1;  // params.substep || 1
    const bool literal = (j->flags & FSIM_JACOBI_LITERAL) != 0;
    struct EventPair {  // destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        EventPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
        ~EventPair() { cudaEventDestroy(a); cudaEventDestroy(b); }
    } ev;
    const cudaEvent_t e0 = ev.a, e1 = ev.b;
    while (it < max_iterations && diff > tolerance) {
        for (int s = 0; s < sub_n; ++s) {
            // programSet: x_guess <- x_result (:649-652), then x_result <- R x_guess + C (:655)
            FSIM_CUDA(cudaMemcpyAsync(j->xg, j->xr, (size_t)L * j->rs, cudaMemcpyDeviceToDevice, j->stream));
            cudaEventRecord(e0, j->stream);
            FSIM_TRY(jmv(j, j->xg, j->xr));
            cudaEventRecord(e1, j->stream);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            j->mv_ms += ms;
            j->mv_launches++;
        }
        rc = jdispatch(j, [&](auto tag) {
            using Real = decltype(tag);
            jacobi_stats_kernel<Real><<<grid_for(npix, 256), 256, 0, j->stream>>>(
                (const Real *)j->xg, (const Real *)j->xr, (Real *)j->stats, npix);
            FSIM_CUDA(cudaGetLastError());
            j->launches++;
            return (int)FSIM_OK;
        });
        FSIM_TRY(rc);
        FSIM_TRY(jdownload(j, j->stats, st.data(), L));
        FSIM_TRY(jdownload(j, j->xg, x1a.data(), L));
        FSIM_TRY(jdownload(j, j->xr, x2a.data(), L));
        if (!literal) x1 = x2 = x1x2 = x1x1 = x2x2 = 0;  // the reference never resets them (:628-634)
        double max_diff = 0.0;
        for (int64_t i = 0; i < npix; ++i) {  // :675-683
            x1 += x1a[4 * i] + x1a[4 * i + 1] + x1a[4 * i + 2] + x1a[4 * i + 3];
            x2 += x2a[4 * i] + x2a[4 * i + 1] + x2a[4 * i + 2] + x2a[4 * i + 3];
            x1x2 += st[4 * i];
            x1x1 += st[4 * i + 1];
            x2x2 += st[4 * i + 2];
            max_diff = std::max(max_diff, st[4 * i + 3]);
        }
        const double Ld = (double)L;
        corr = (Ld * x1x2 - x1 * x2) / sqrt((Ld * x1x1 - x1 * x1) * (Ld * x2x2 - x2 * x2));  // :686
        diff = 2 * Ld * max_diff / (fabs(x1) + fabs(x2));                                    // :687
        it++;
    }
    if (correlation) *correlation = corr;
    if (diff_out) *diff_out = diff;
    if (iterations) *iterations = it;
    if (result) {
        if (it == 0) FSIM_TRY(jdownload(j, j->xr, result, L));
        else memcpy(result, x2a.data(), sizeof(double) * (size_t)L);
    }
    return FSIM_OK;
}

int fsim_jacobi_timing(fsim_jacobi *j, double *mv_ms, int64_t *mv_launches)
{
    if (!j) return jfail(FSIM_ERR_INVALID, "null handle");
    if (mv_ms) *mv_ms = j->mv_ms;
    if (mv_launches) *mv_launches = j->mv_launches;
    return FSIM_OK;
}

}  // extern "C"
