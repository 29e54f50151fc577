// api.cu -- the C ABI of include/fusionsim.h: handle lifetime, out.set() conversions, the
// step()/density() orchestration and the accessors.  Host code + small layout-conversion kernels.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace fsim {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[1024];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line,
             what);
    set_error(buf);
    return FSIM_ERR_CUDA;
}

int upload_shape(const double *shape64, const double *shape32);
int upload_costab(const double *c64, const float *c32);

static int fail(int code, const std::string &msg)
{
    set_error(msg);
    return code;
}

int ensure_stage(fsim_sim *s, size_t bytes)
{
    if (bytes <= s->stage_bytes) return FSIM_OK;
    if (s->stage) FSIM_CUDA(cudaFree(s->stage));
    s->stage = nullptr;
    s->stage_bytes = 0;
    FSIM_CUDA(cudaMalloc(&s->stage, bytes));
    s->stage_bytes = bytes;
    return FSIM_OK;
}

// N(num) = num.toFixed(20) (empic.js:23-25), parsed back as the GLSL compiler would
static double tofixed20(double x)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%.20f", x);
    return strtod(buf, nullptr);
}

// ---- layout-conversion kernels ------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// default rand / entropy when the caller supplies none (the reference uses Math.random and
// window.crypto, empic.js:148-173: unseeded).  kind 0: [0,1) like Math.random; kind 1:
// uint32 / 0xFFFFFFFF like empic.js:151-154.
template <typename Real>
__global__ void fill_random_kernel(Real *out, int64_t n, int64_t stride, uint64_t seed, int kind)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t hsh = splitmix64(seed ^ splitmix64((uint64_t)k));
    double v = kind ? (double)(uint32_t)(hsh >> 32) / (double)0xFFFFFFFFu
                    : (double)(hsh >> 11) * (1.0 / 9007199254740992.0);
    out[k * stride] = (Real)v;
}

// default rand of a particle is a function of its ID (not of the slot it happens to be stored in)
template <typename Real>
__global__ void fill_random_by_id_kernel(Real *out, const uint32_t *__restrict__ pid, uint32_t id_base, int64_t n, uint64_t seed)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t hsh = splitmix64(seed ^ splitmix64((uint64_t)(pid[k] - id_base)));
    out[k] = (Real)((double)(hsh >> 11) * (1.0 / 9007199254740992.0));
}

// Sort on ingest (fsim_set_position on a freshly initialised handle): the gather cell of every particle straight from
// the staged host array [N][3], indexed by particle id; histogram for the counting sort.
template <typename Real>
__global__ void __launch_bounds__(256)
ingest_keys_kernel(const double *__restrict__ in, int64_t n, double f0, double f1, double f2, int nr, int nz, int row0, int rows,
                   uint32_t *__restrict__ key, uint32_t *__restrict__ counts)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < n;
    uint32_t c = 0xffffffffu;
    if (valid) {
        const Real x = (Real)(in[3 * p] * f0), y = (Real)(in[3 * p + 1] * f1), z = (Real)(in[3 * p + 2] * f2);  // as part3_in_kernel
        const Real r = fsqrt(x * x + y * y);
        int cj = tex_idx(z, nz) - row0;
        cj = cj < 0 ? 0 : (cj >= rows ? rows - 1 : cj);
        c = (uint32_t)tex_idx(r, nr) + (uint32_t)cj * (uint32_t)nr;
        key[p] = c;
    }
    int leader;
    uint32_t len, rank;
    warp_runs(c, (int)(threadIdx.x & 31), leader, len, rank);
    if (valid && rank == 0) atomicAdd(counts + c, len);
}

__global__ void ids_from_perm_kernel(uint32_t *__restrict__ pid, const uint32_t *__restrict__ perm, int64_t n, uint32_t id_base)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) pid[j] = id_base + (perm[j] & KEY_MASK);
}

template <typename Real>
__global__ void iota_ids_kernel(uint32_t *id, int64_t n, uint32_t base)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) id[k] = base + (uint32_t)k;
}

// host [nr][nz][3] doubles -> local table [cell][3], cell = i + (j-row0)*nr  (empic.js:1162)
template <typename Real>
__global__ void field_in_kernel(const double *__restrict__ in, Real *__restrict__ out, int nr, int nz,
                                int row0, int rows)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)nr * rows) return;
    const int i = (int)(c % nr);
    int j = (int)(c / nr) + row0;
    if (j < 0) j += nz;  // periodic z: ghost rows take the wrapped rows
    if (j >= nz) j -= nz;
    const double *p = in + 3 * ((size_t)i * nz + j);
    out[3 * c] = (Real)p[0]; out[3 * c + 1] = (Real)p[1]; out[3 * c + 2] = (Real)p[2];
}

// value.position / value.velocity: [N][3] doubles * (factor_r, factor_r, factor_z), stored to the
// typed array (empic.js:1202-1205, :1226-1229); particle id -> storage slot through pid[].
template <typename Real>
__global__ void part3_in_kernel(const double *__restrict__ in, Real *a0, Real *a1, Real *a2,
                                const uint32_t *__restrict__ pid, uint32_t id_base, int64_t n, double f0,
                                double f1, double f2, uint8_t *alive, int by_id)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const size_t src = by_id ? (size_t)(pid[p] - id_base) : (size_t)p;
    a0[p] = (Real)(in[3 * src] * f0);
    a1[p] = (Real)(in[3 * src + 1] * f1);
    a2[p] = (Real)(in[3 * src + 2] * f2);
    if (alive) alive[p] = 1;  // position.w = 1.0, empic.js:1205
}

template <typename Real>
__global__ void part4_in_kernel(const double *__restrict__ in, Real *a0, Real *a1, Real *a2, Real *a3,
                                const uint32_t *__restrict__ pid, uint32_t id_base, int64_t n, int by_id)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const size_t src = by_id ? (size_t)(pid[p] - id_base) : (size_t)p;
    a0[p] = (Real)in[4 * src]; a1[p] = (Real)in[4 * src + 1];
    a2[p] = (Real)in[4 * src + 2]; a3[p] = (Real)in[4 * src + 3];
}

// checkpoint restore: [N][4] normalised x, y, z, alive (exactly what fsim_get_position returns)
template <typename Real>
__global__ void state_pos_in_kernel(const double *__restrict__ in, Real *x, Real *y, Real *z, uint8_t *alive,
                                    const uint32_t *__restrict__ pid, uint32_t id_base, int64_t n, int by_id)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const size_t src = by_id ? (size_t)(pid[p] - id_base) : (size_t)p;
    x[p] = (Real)in[4 * src]; y[p] = (Real)in[4 * src + 1]; z[p] = (Real)in[4 * src + 2];
    alive[p] = in[4 * src + 3] > 0.5 ? 1 : 0;
}

// [cells][4] doubles -> planar grid field (inverse of planar_out_kernel); [cells] -> one plane when nch == 1
template <typename Real>
__global__ void planar_in_kernel(const double *__restrict__ in, Real *__restrict__ out, int nr, int rows, int pitch,
                                 int64_t plane, int nch)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)nr * rows) return;
    const size_t o = (size_t)(c / nr) * pitch + (size_t)(c % nr);
    for (int q = 0; q < nch; ++q) out[q * plane + o] = (Real)in[nch * c + q];
}

// storage order -> id order (or storage order when by_id == 0)
template <typename Real>
__global__ void part_out_kernel(double *__restrict__ out, int width, const Real *a0, const Real *a1,
                                const Real *a2, const Real *a3, const uint8_t *alive,
                                const uint32_t *__restrict__ pid, uint32_t id_base, int64_t n, int by_id)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const size_t d = by_id ? (size_t)(pid[p] - id_base) : (size_t)p;
    out[width * d] = (double)a0[p];
    out[width * d + 1] = (double)a1[p];
    out[width * d + 2] = (double)a2[p];
    if (width == 4) out[4 * d + 3] = alive ? (double)alive[p] : (double)a3[p];
}

template <typename Real>
__global__ void cells_out_kernel(int64_t *__restrict__ out, const Real *x, const Real *y, const Real *z,
                                 const uint32_t *__restrict__ pid, uint32_t id_base, int64_t n, int nr,
                                 int nz, int by_id)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const size_t d = by_id ? (size_t)(pid[p] - id_base) : (size_t)p;
    const Real r = fsqrt(x[p] * x[p] + y[p] * y[p]);
    out[d] = (int64_t)tex_idx(r, nr) + (int64_t)nr * tex_idx(z[p], nz);
}

template <typename Real>
__global__ void convert_in_kernel(const double *__restrict__ in, Real *__restrict__ out, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (Real)in[k];
}
template <typename Real>
__global__ void convert_out_kernel(const Real *__restrict__ in, double *__restrict__ out, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (double)in[k];
}

// planar grid field (common.cuh) -> [cell][4] doubles, cell = i + j*nr
template <typename Real>
__global__ void planar_out_kernel(const Real *__restrict__ in, double *__restrict__ out, int nr, int rows,
                                  int pitch, int64_t plane)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)nr * rows) return;
    const size_t o = (size_t)(c / nr) * pitch + (size_t)(c % nr);
#pragma unroll
    for (int q = 0; q < 4; ++q) out[4 * c + q] = (double)in[q * plane + o];
}

// ---- run invariants (fsim_check_digest): particle ids and the deposit, reduced on the device ----------
// d[0] ^= id, d[1] += id over the live slots; d[2] += count over the owned cells (u64 atomics: exact, order-free)
__global__ void __launch_bounds__(256)
digest_ids_kernel(const uint32_t *__restrict__ id, int64_t n_host, const uint32_t *__restrict__ n_dev, unsigned long long *d)
{
    const int64_t n = live_count(n_dev, n_host);
    unsigned long long x = 0, sum = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        x ^= id[p];
        sum += id[p];
    }
    for (int o = 16; o; o >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicXor(d, x);
        atomicAdd(d + 1, sum);
    }
}
// counts and the weight channel of the per-cell sums over the owned rows; the floating-point partial sums
// are formed in a fixed order (one per block, tree inside the block) and added in block order by the host
template <typename Real>
__global__ void __launch_bounds__(256)
digest_cells_kernel(const uint32_t *__restrict__ count, const Real *__restrict__ alpha, int nr, int pitch, int row_lo,
                    int nrows, unsigned long long *d, double *partial)
{
    __shared__ double sh[8];
    __shared__ unsigned long long shc[8];
    const int64_t ncell = (int64_t)nr * nrows;
    const int64_t per = (ncell + gridDim.x - 1) / gridDim.x;
    const int64_t c0 = (int64_t)blockIdx.x * per, c1 = min(ncell, c0 + per);
    double a = 0.0;
    unsigned long long cnt = 0;
    for (int64_t c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
        const int64_t j = row_lo + c / nr, i = c % nr;
        cnt += count[j * nr + i];
        a += (double)alpha[j * pitch + i];
    }
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = a; shc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        unsigned long long tc = 0;
        for (int w = 0; w < 8; ++w) { t += sh[w]; tc += shc[w]; }
        partial[blockIdx.x] = t;
        atomicAdd(d + 2, tc);
    }
}

// value.sink_mask [nr][nz] -> 1 bit per global cell i + j*nr; the shader tests .r > 0.5 (:719).
// One warp per 32-cell word (ballot): 2 MB instead of 16.7 MB at 8192 x 2048, resident in L2 for good.
__global__ void sink_in_kernel(const double *__restrict__ in, uint32_t *__restrict__ out, int nr, int nz)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // blockDim is a multiple of 32
    bool keep = false;
    if (c < (int64_t)nr * nz) {
        const int i = (int)(c % nr), j = (int)(c / nr);
        keep = in[(size_t)i * nz + j] > 0.5;
    }
    const unsigned word = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && c < (int64_t)nr * nz) out[c >> 5] = word;
}
__global__ void sink_out_kernel(const uint32_t *__restrict__ in, uint8_t *__restrict__ out, int64_t ncell)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncell) out[c] = (in[c >> 5] >> (c & 31)) & 1u;
}

// ---- host-side restatement of the source_pdf -> inverse-cdf table build (empic.js:1268-1339) --
static int build_inv_cdf(const double *pdf, int64_t n0, int64_t n1, std::vector<double> &out)
{
    std::vector<double> cdf_y((size_t)(n0 * n1)), cdf_x((size_t)n0);
    double sum_x = 0.0;
    for (int64_t i = 0; i < n0; ++i) {
        double sum_y = 0.0;
        double *row = cdf_y.data() + i * n1;
        for (int64_t j = 0; j < n1; ++j) {
            sum_y += pdf[i * n1 + j];
            row[j] = sum_y;
        }
        for (int64_t j = 0; j < n1; ++j) row[j] /= sum_y;  // 0/0 = NaN rows stay NaN, as in JS
        sum_x += sum_y;
        cdf_x[i] = sum_x;
    }
    for (int64_t i = 0; i < n0; ++i) cdf_x[i] /= sum_x;
    const int T = FSIM_N_INVCDF;
    out.assign((size_t)T * T * 2, 0.0);
    for (int ti = 0; ti < T; ++ti) {
        const double f1 = (double)ti / 511.0;
        if (f1 < 0 || f1 > 1) return fail(FSIM_ERR_RANGE, "function out of range");  // :1294-1296
        int64_t i = 0;
        while (i < n0 && cdf_x[i] < f1) ++i;
        double x;
        if (i >= n0) x = NAN;
        else if (i == 0) x = (f1 / cdf_x[0]) / (double)n0;
        else x = ((double)i + (f1 - cdf_x[i - 1]) / (cdf_x[i] - cdf_x[i - 1])) / (double)n0;
        const double fi = floor(x * (double)n0);
        const int64_t iy = (fi < (double)(n0 - 1)) ? (int64_t)fi : n0 - 1;
        for (int tj = 0; tj < T; ++tj) {
            const double f2 = (double)tj / 511.0;
            double y;
            if (!(fi == fi) || iy < 0) y = NAN;
            else {
                const double *cy = cdf_y.data() + iy * n1;
                int64_t j = 0;
                while (j < n1 && cy[j] < f2) ++j;
                if (j >= n1) y = NAN;
                else if (j == 0) y = (f2 / cy[0]) / (double)n1;
                else y = ((double)j + (f2 - cy[j - 1]) / (cy[j] - cy[j - 1])) / (double)n1;
            }
            out[2 * ((size_t)ti + (size_t)tj * T)] = x;
            out[2 * ((size_t)ti + (size_t)tj * T) + 1] = y;
        }
    }
    return FSIM_OK;
}

// deposit footprint, empic.js:949-971 (as_f32: with the Float32Array round trips)
static void build_shape(double *out, bool as_f32)
{
    const int n = FSIM_NSHAPE;
    const double mid = (n - 1) / 2.0;
    double sum = 0.0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            const double d = sqrt(pow(i - mid, 2) + pow(j - mid, 2));
            const double c = cos(0.5 * M_PI * d / mid);
            double v = pow(c > 0.0 ? c : 0.0, 2);
            if (as_f32) v = (double)(float)v;
            out[i + n * j] = v;
            sum += v;
        }
    for (int k = 0; k < n * n; ++k) {
        const double v = out[k] / sum;
        out[k] = as_f32 ? (double)(float)v : v;
    }
}

int check_handle(fsim_sim *s)
{
    if (!s) return fail(FSIM_ERR_INVALID, "null simulation handle");
    if (s->sticky_error) return fail(FSIM_ERR_CUDA, "handle is in a sticky CUDA error state: " + g_last_error);
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
    return FSIM_OK;
}
// Second stream for the stencil and the canvas draws (common.cuh).  Created on first use, one priority level
// above the main stream: its blocks are scheduled as soon as the sweep's retire.
cudaStream_t post_begin(fsim_sim *s)
{
    if (!(s->spec.flags & FSIM_FLAG_POST_STREAM)) return s->stream;  // default: one stream (measured faster)
    if (!s->post_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (s->spec.flags & FSIM_FLAG_POST_NO_PRIORITY) hi = 0;
        if (cudaStreamCreateWithPriority(&s->post_stream, cudaStreamNonBlocking, hi) != cudaSuccess) return s->stream;
        cudaEventCreateWithFlags(&s->post_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s->post_done, cudaEventDisableTiming);
    }
    cudaEventRecord(s->post_fork, s->stream);
    cudaStreamWaitEvent(s->post_stream, s->post_fork, 0);
    return s->post_stream;
}
int post_end(fsim_sim *s)
{
    if (!s->post_stream || !(s->spec.flags & FSIM_FLAG_POST_STREAM)) return FSIM_OK;
    FSIM_CUDA(cudaEventRecord(s->post_done, s->post_stream));
    s->post_pending = true;
    return FSIM_OK;
}
int post_join(fsim_sim *s)
{
    if (!s->post_pending) return FSIM_OK;
    FSIM_CUDA(cudaStreamWaitEvent(s->stream, s->post_done, 0));
    s->post_pending = false;
    return FSIM_OK;
}
// every entry point except the frame's own (step, density, migrate, canvas draws) first joins the post stream.
// check / check_n are the prologue of entry points that may CHANGE what a frame launches or works on (they bump
// config_epoch: a captured frame graph, fsim_run_frames, is dropped); check_ro / check_n_ro of those that only read.
static int check_ro(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    return post_join(s);
}
static int check(fsim_sim *s)
{
    FSIM_TRY(check_ro(s));
    s->config_epoch++;
    return FSIM_OK;
}
// entry points that address particles by count: the asynchronous slab exchange keeps the exact count on
// the device (settle_count synchronises and refreshes fsim_sim::n)
static int check_n_ro(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    return settle_count(s);
}
static int check_n(fsim_sim *s)
{
    FSIM_TRY(check_n_ro(s));
    s->config_epoch++;
    return FSIM_OK;
}

static int finish(fsim_sim *s, int rc)
{
    if (rc == FSIM_ERR_CUDA) s->sticky_error = true;
    return rc;
}

// allocations are zeroed ON THE HANDLE'S STREAM (a cudaStreamNonBlocking stream is not ordered after
// the legacy default stream a plain cudaMemset runs on)
static thread_local cudaStream_t g_alloc_stream = nullptr;
template <typename T>
static int dalloc(T **p, size_t count, bool zero = true)
{
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    FSIM_CUDA(cudaMalloc((void **)p, bytes));
    if (zero) FSIM_CUDA(cudaMemsetAsync(*p, 0, bytes, g_alloc_stream));
    return FSIM_OK;
}
static int dalloc_bytes(void **p, size_t bytes, bool zero = true)
{
    bytes = std::max<size_t>(bytes, 16);
    FSIM_CUDA(cudaMalloc(p, bytes));
    if (zero) FSIM_CUDA(cudaMemsetAsync(*p, 0, bytes, g_alloc_stream));
    return FSIM_OK;
}

static int stage_in(fsim_sim *s, const void *host, size_t bytes)
{
    FSIM_TRY(ensure_stage(s, bytes));
    FSIM_CUDA(cudaMemcpyAsync(s->stage, host, bytes, cudaMemcpyHostToDevice, s->stream));
    return FSIM_OK;
}
static int stage_out(fsim_sim *s, void *host, size_t bytes)
{
    FSIM_CUDA(cudaMemcpyAsync(host, s->stage, bytes, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    return FSIM_OK;
}

// seed of the default rand channel q (the reference draws Math.random(), empic.js:168-173: unseeded)
static uint64_t default_seed(const fsim_sim *s, int q)
{
    return 0x5EEDF0510Cull + 17 * (uint64_t)(q + 1) + ((uint64_t)s->id_base << 20);
}

static int create_impl(const fsim_spec *sp, fsim_sim *s)
{
    s->spec = *sp;
    s->prec = sp->precision;
    s->rs = sp->precision == FSIM_F64 ? 8 : 4;
    s->device = sp->device;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(FSIM_ERR_CUDA, std::string("no CUDA device: libfusionsim has no CPU fallback (") +
                                       cudaGetErrorString(e) + ")");
    if (sp->device < 0 || sp->device >= ndev) return fail(FSIM_ERR_INVALID, ".device <- no such CUDA device");
    FSIM_CUDA(cudaSetDevice(s->device));
    FSIM_CUDA(cudaDeviceGetAttribute(&s->nsm, cudaDevAttrMultiProcessorCount, s->device));
    FSIM_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    g_alloc_stream = s->stream;

    s->nr = (int)sp->nr;
    s->nz = (int)sp->nz;
    s->ncell_global = (int64_t)s->nr * s->nz;
    if (sp->slab_rows > 0) {
        s->slab = true;
        s->own0 = (int)sp->slab_row0;
        s->own_rows = (int)sp->slab_rows;
        const int lo = std::max<int64_t>(0, sp->slab_row0 - sp->halo_rows);
        const int hi = std::min<int64_t>(s->nz, sp->slab_row0 + sp->slab_rows + sp->halo_rows);
        s->row0 = lo;
        s->rows = hi - lo;
    } else if (sp->flags & FSIM_FLAG_PERIODIC_Z) {  // EXTENSION: nz owned rows + ghost rows that hold the wrapped rows
        const int h = (int)std::max<int64_t>(sp->halo_rows, 8);
        s->ring = true;
        s->own0 = 0; s->own_rows = s->nz; s->row0 = -h; s->rows = s->nz + 2 * h;
    } else {
        s->own0 = 0; s->own_rows = s->nz; s->row0 = 0; s->rows = s->nz;
    }
    s->ncell_local = (int64_t)s->nr * s->rows;
    s->pitch = (s->nr + 3) / 4 * 4;
    s->plane = (int64_t)s->pitch * s->rows;
#ifdef FSIM_TUNE
    if (const char *e = getenv("FSIM_PLANE_PAD")) s->plane += atoll(e) / 4 * 4;  // tuning build only: channel-plane stride padding (reals)
#endif

    // physical quantities, empic.js:44-46, :852 and the toFixed(20) literals of :527,:606,:647
    s->h = sp->particle_charge * sp->dt / (2 * sp->particle_mass);
    s->factor_r = 1 / sp->radius;
    s->factor_z = 1 / sp->height;
    s->step_factor = sp->dt * FSIM_C_LIGHT;
    s->k13 = tofixed20(s->factor_r / s->factor_z);
    s->k31 = tofixed20(s->factor_z / s->factor_r);
    s->kr = tofixed20(s->factor_r);
    s->kz = tofixed20(s->factor_z);

    s->n = sp->nparticles_total > 0 ? sp->nparticles_total : sp->nparticles * sp->nparticles;
    s->cap = std::max<int64_t>(sp->capacity, s->n);
    s->cap = (s->cap + 1023) / 1024 * 1024 + 1024;
    s->id_base = (uint32_t)sp->id_base;

    for (int b = 0; b < 2; ++b) {
        for (int k = 0; k < NPART_ARRAYS; ++k) FSIM_TRY(dalloc_bytes(&s->part[b][k], s->rs * s->cap));
        FSIM_TRY(dalloc(&s->alive[b], s->cap));
        FSIM_TRY(dalloc(&s->pid[b], s->cap));
    }
    FSIM_TRY(dalloc(&s->key, s->cap));
    FSIM_TRY(dalloc(&s->perm, s->cap));
    for (int q = 0; q < 2; ++q) FSIM_TRY(dalloc_bytes(&s->dcol[q], s->rs * s->cap));
    FSIM_TRY(dalloc(&s->counts, s->ncell_local + 1));
    FSIM_TRY(dalloc(&s->starts, s->ncell_local + 2));
    FSIM_TRY(dalloc(&s->cursor, s->ncell_local + 1));
    FSIM_TRY(dalloc(&s->blocksums, s->ncell_local / 2048 + 2));
    FSIM_TRY(dalloc_bytes(&s->cellrec, s->rs * RECSTRIDE * s->ncell_local));
    FSIM_TRY(dalloc_bytes(&s->E, s->rs * 3 * s->ncell_local));
    FSIM_TRY(dalloc_bytes(&s->B, s->rs * 3 * s->ncell_local));
    FSIM_TRY(dalloc(&s->sink, (s->ncell_global + 31) / 32));
    FSIM_TRY(dalloc_bytes(&s->entropy, s->rs * 4 * FSIM_N_ENTROPY * FSIM_N_ENTROPY));
    FSIM_TRY(dalloc_bytes(&s->invcdf, s->rs * 2 * FSIM_N_INVCDF * FSIM_N_INVCDF));
    FSIM_TRY(dalloc_bytes(&s->cellsum, s->rs * 4 * s->plane));
    FSIM_TRY(make_sums_tensor_map(s, 18));  // re-encoded by launch_conv if its tile height differs
    FSIM_TRY(dalloc(&s->cellcount, s->ncell_local));
    FSIM_TRY(dalloc_bytes(&s->avg, s->rs * 4 * s->plane));
    if (sp->flags & FSIM_FLAG_KEEP_MOMENTS) {
        FSIM_TRY(dalloc_bytes(&s->mom, s->rs * 4 * s->plane));
        FSIM_TRY(dalloc_bytes(&s->norm, s->rs * 4 * s->plane));
    }
    FSIM_TRY(dalloc(&s->heavy_list, s->cap / 256 + 2));
    FSIM_TRY(dalloc(&s->medium_list, s->cap / 16 + 2));
    FSIM_TRY(dalloc(&s->heavy_n, 2));
    FSIM_TRY(dalloc(&s->oob, 1));
    FSIM_TRY(dalloc(&s->mscratch, MC_WORDS));
    FSIM_TRY(dalloc(&s->hole_flag, s->cap));
    if (s->slab) FSIM_TRY(dalloc(&s->leavers, s->cap));

    // host-computed constant tables (libm): deposit footprint and quadrature cosines
    double shape64[FSIM_NSHAPE * FSIM_NSHAPE], shape32[FSIM_NSHAPE * FSIM_NSHAPE];
    build_shape(shape64, false);
    build_shape(shape32, true);
    FSIM_TRY(upload_shape(shape64, shape32));
    // empic.js:317: the argument is formed in the shader's working precision, its cosine is the host libm's
    double costab[FSIM_NQUAD];
    float costab32[FSIM_NQUAD];
    host_cos_tables(costab, costab32);
    FSIM_TRY(upload_costab(costab, costab32));

    // ids, default rand / entropy
    const uint64_t seed = 0x5EEDF0510Cull;
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        if (s->n) {
            for (int b = 0; b < 2; ++b)
                iota_ids_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(s->pid[b], s->n, s->id_base);
            for (int q = 0; q < 4; ++q)
                fill_random_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
                    (Real *)s->part[0][AQ0 + q], s->n, 1, default_seed(s, q), 0);
        }
        const int64_t ne = 4ll * FSIM_N_ENTROPY * FSIM_N_ENTROPY;
        fill_random_kernel<Real><<<grid_for(ne, 256), 256, 0, s->stream>>>((Real *)s->entropy, ne, 1, seed, 1);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    return FSIM_OK;
}

static void free_all(fsim_sim *s)
{
    for (int b = 0; b < 2; ++b) {
        for (int k = 0; k < NPART_ARRAYS; ++k) cudaFree(s->part[b][k]);
        cudaFree(s->alive[b]);
        cudaFree(s->pid[b]);
    }
    void *ptrs[] = {s->key, s->perm, s->dcol[0], s->dcol[1], s->counts, s->starts, s->cursor, s->blocksums, s->cellrec, s->E, s->B, s->sink,
                    s->entropy, s->invcdf, s->cellsum, s->cellcount, s->mom, s->norm, s->avg,
                    s->heavy_list, s->medium_list, s->heavy_n, s->oob, s->stage, s->migr, s->mscratch, s->hole_flag, s->halo_buf, s->leavers};
    for (void *p : ptrs) cudaFree(p);
    cudaFree(s->phi[0]); cudaFree(s->phi[1]); cudaFree(s->rho_src); cudaFree(s->relax_coef); cudaFree(s->background);
    em_free(s);
    if (s->frame_graph) cudaGraphExecDestroy(s->frame_graph);
    s->frame_graph = nullptr;
    cudaFree(s->bmag); cudaFree(s->plan.send); cudaFree(s->plan.recv); cudaFree(s->plan.holes); cudaFree(s->plan.targets); cudaFree(s->plan.sources);
    if (s->n_pinned) cudaFreeHost(s->n_pinned);
    for (auto &e : s->n_event)
        if (e) cudaEventDestroy(e);
    for (auto &kv : s->timers)
        for (auto &pe : kv.second.pending) {
            cudaEventDestroy(pe.first);
            cudaEventDestroy(pe.second);
        }
    for (auto &m : s->marks)
        if (m) cudaEventDestroy(m);
    for (int k = 0; k < 2; ++k) {
        cudaFree(s->canvas_dev[k]);
        if (s->render_done[k]) cudaEventDestroy(s->render_done[k]);
        if (s->copy_done[k]) cudaEventDestroy(s->copy_done[k]);
    }
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->post_stream) {
        cudaStreamSynchronize(s->post_stream);
        cudaStreamDestroy(s->post_stream);
        cudaEventDestroy(s->post_fork);
        cudaEventDestroy(s->post_done);
    }
    if (s->stream && !s->ext_stream) cudaStreamDestroy(s->stream);
}

static int particles_in3(fsim_sim *s, const double *host, int a0, double f0, double f1, double f2,
                         bool set_alive)
{
    if (!host) return fail(FSIM_ERR_INVALID, "null array");
    // single GPU: element p of the host array is particle id p wherever it is stored;
    // slab mode: element p is storage slot p (ids are global there, see fsim_set_ids)
    if (s->n == 0) return FSIM_OK;
    FSIM_TRY(stage_in(s, host, sizeof(double) * 3 * s->n));
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        part3_in_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
            (const double *)s->stage, (Real *)s->part[s->cur][a0], (Real *)s->part[s->cur][a0 + 1],
            (Real *)s->part[s->cur][a0 + 2], s->pid[s->cur], s->id_base, s->n, f0, f1, f2,
            set_alive ? s->alive[s->cur] : nullptr, s->slab ? 0 : 1);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
}

static int field_in(fsim_sim *s, const double *host, void *dst)
{
    if (!host) return fail(FSIM_ERR_INVALID, "null array");
    FSIM_TRY(stage_in(s, host, sizeof(double) * 3 * s->ncell_global));
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        field_in_kernel<Real><<<grid_for(s->ncell_local, 256), 256, 0, s->stream>>>(
            (const double *)s->stage, (Real *)dst, s->nr, s->nz, s->row0, s->rows);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
}

static int table_in(fsim_sim *s, const double *host, void *dst, int64_t count)
{
    if (!host) return fail(FSIM_ERR_INVALID, "null array");
    FSIM_TRY(stage_in(s, host, sizeof(double) * count));
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        convert_in_kernel<Real><<<grid_for(count, 256), 256, 0, s->stream>>>((const double *)s->stage,
                                                                            (Real *)dst, count);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
}

static int table_out(fsim_sim *s, const void *src, double *host, int64_t count)
{
    if (!host) return fail(FSIM_ERR_INVALID, "null array");
    FSIM_TRY(ensure_stage(s, sizeof(double) * count));
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        convert_out_kernel<Real><<<grid_for(count, 256), 256, 0, s->stream>>>((const Real *)src,
                                                                             (double *)s->stage, count);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    return stage_out(s, host, sizeof(double) * count);
}

static int planar_out(fsim_sim *s, const void *src, double *host)
{
    if (!host) return fail(FSIM_ERR_INVALID, "null array");
    const int64_t nc = s->ncell_local;
    FSIM_TRY(ensure_stage(s, sizeof(double) * 4 * nc));
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        planar_out_kernel<Real><<<grid_for(nc, 256), 256, 0, s->stream>>>((const Real *)src, (double *)s->stage,
                                                                         s->nr, s->rows, s->pitch, s->plane);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    return stage_out(s, host, sizeof(double) * 4 * nc);
}

static int collect_timers(fsim_sim *s)
{
    FSIM_TRY(post_join(s));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    for (auto &kv : s->timers) {
        for (auto &pe : kv.second.pending) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, pe.first, pe.second);
            kv.second.ms += ms;
            cudaEventDestroy(pe.first);
            cudaEventDestroy(pe.second);
        }
        kv.second.pending.clear();
    }
    return FSIM_OK;
}

}  // namespace fsim

using namespace fsim;

extern "C" {

const char *fsim_last_error(void) { return g_last_error.c_str(); }
int fsim_abi_version(void) { return FSIM_ABI_VERSION; }

int fsim_create(const fsim_spec *sp, fsim_sim **out)
{
    if (!sp || !out) return fail(FSIM_ERR_INVALID, "null argument");
    *out = nullptr;
    // spec validation: the reference checks that each property is a number (empic.js:31-41,
    // utilities.js:118-127); here they are numbers by type, so reject what cannot be simulated.
    auto bad = [](double v) { return !(v == v) || isinf(v); };
    if (bad(sp->radius) || sp->radius == 0) return fail(FSIM_ERR_INVALID, ".radius <- must be a finite non-zero number");
    if (bad(sp->height) || sp->height == 0) return fail(FSIM_ERR_INVALID, ".height <- must be a finite non-zero number");
    if (sp->nr < 1 || sp->nr > (1 << 20)) return fail(FSIM_ERR_INVALID, ".nr <- must be in [1, 2^20]");
    if (sp->nz < 1 || sp->nz > (1 << 20)) return fail(FSIM_ERR_INVALID, ".nz <- must be in [1, 2^20]");
    if (sp->nr * sp->nz > (1ll << 31) - 2) return fail(FSIM_ERR_INVALID, ".nr <- nr*nz exceeds 2^31 cells");
    if (bad(sp->dt)) return fail(FSIM_ERR_INVALID, ".dt <- must be a finite number");
    if (sp->nparticles < 0 || sp->nparticles > 65535) return fail(FSIM_ERR_INVALID, ".nparticles <- must be in [0, 65535]");
    if (bad(sp->particle_mass) || sp->particle_mass == 0) return fail(FSIM_ERR_INVALID, ".particle_mass <- must be a finite non-zero number");
    if (bad(sp->particle_charge)) return fail(FSIM_ERR_INVALID, ".particle_charge <- must be a finite number");
    if (sp->precision != FSIM_F64 && sp->precision != FSIM_F32) return fail(FSIM_ERR_INVALID, ".precision <- must be FSIM_F64 or FSIM_F32");
    // slot indices travel with a flag in bit 31 (perm[], key[]) and particle ids are 32-bit: keep both below 2^31 / 2^32
    const int64_t n_req = sp->nparticles_total > 0 ? sp->nparticles_total : sp->nparticles * sp->nparticles;
    const int64_t slots = std::max<int64_t>(n_req, sp->capacity);
    if (sp->nparticles_total < 0 || sp->capacity < 0 || slots > 0x7fffffffll - 4096)
        return fail(FSIM_ERR_INVALID, ".nparticles_total <- particle slots per handle must stay below 2^31");
    if (sp->id_base > 0xffffffffull || sp->id_base + (uint64_t)slots > 0xffffffffull)
        return fail(FSIM_ERR_INVALID, ".id_base <- id_base + particle slots must stay below 2^32 (ids are 32-bit)");
    if (sp->slab_rows < 0 || sp->slab_row0 < 0 || sp->slab_row0 + sp->slab_rows > sp->nz)
        return fail(FSIM_ERR_INVALID, ".slab_rows <- slab outside the grid");
    if ((sp->flags & FSIM_FLAG_PERIODIC_Z) && (sp->slab_rows > 0 || sp->nz < 8))
        return fail(FSIM_ERR_UNSUPPORTED, ".flags <- periodic z runs on one GPU (no slab) and needs at least 8 rows");
    if (sp->slab_rows > 0 && sp->halo_rows < FSIM_SHAPE_MID)
        return fail(FSIM_ERR_INVALID, ".halo_rows <- slab mode needs at least 5 halo rows (deposit footprint)");
    fsim_sim *s = new fsim_sim();
    int rc = create_impl(sp, s);
    if (rc != FSIM_OK) {
        free_all(s);
        delete s;
        return rc;
    }
    *out = s;
    return FSIM_OK;
}

int fsim_destroy(fsim_sim *s)
{
    if (!s) return FSIM_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    free_all(s);
    delete s;
    return FSIM_OK;
}

int fsim_set_E(fsim_sim *s, const double *E)
{
    FSIM_TRY(check(s));
    return finish(s, field_in(s, E, s->E));
}
int fsim_set_B(fsim_sim *s, const double *B)
{
    FSIM_TRY(check(s));
    s->bmag_valid = false;
    return finish(s, field_in(s, B, s->B));
}
// Sort on ingest.  On a freshly initialised handle (create, fsim_set_particle_count) no per-particle state exists yet
// that would have to move along, so the storage can be put into cell order WHILE the positions come in: keys from the
// staged host array, counting sort in id space, slot j := particle perm[j], and the conversion kernels -- which address
// the host arrays by particle id anyway -- then gather 24 contiguous bytes per particle.  The first step() finds the
// storage sorted instead of sorting it with ten random gathers per particle (13-18 ms at 64 Mi particles).
static int ingest_sorted(fsim_sim *s, const double *pos)
{
    if (!pos) return fail(FSIM_ERR_INVALID, "null array");
    FSIM_TRY(stage_in(s, pos, sizeof(double) * 3 * s->n));
    if (s->counts_dirty) {
        FSIM_CUDA(cudaMemsetAsync(s->counts, 0, sizeof(uint32_t) * (s->ncell_local + 1), s->stream));
        s->counts_dirty = false;
    }
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        ingest_keys_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>((const double *)s->stage, s->n, s->factor_r, s->factor_r,
                                                                            s->factor_z, s->nr, s->nz, s->row0, s->rows, s->key, s->counts);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    s->counts_dirty = true;
    FSIM_TRY(launch_bin(s));  // scan + index scatter: perm[j] = id (relative to id_base) of the j-th particle in cell order
    const int c = s->cur;
    ids_from_perm_kernel<<<grid_for(s->n, 256), 256, 0, s->stream>>>(s->pid[c], s->perm, s->n, s->id_base);
    FSIM_CUDA(cudaGetLastError());
    s->launches++;
    rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        part3_in_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
            (const double *)s->stage, (Real *)s->part[c][AX], (Real *)s->part[c][AY], (Real *)s->part[c][AZ], s->pid[c], s->id_base,
            s->n, s->factor_r, s->factor_r, s->factor_z, s->alive[c], 1);
        if (s->rand_default)  // the default rand follows the particle id
            for (int q = 0; q < 4; ++q)
                fill_random_by_id_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
                    (Real *)s->part[c][AQ0 + q], s->pid[c], s->id_base, s->n, default_seed(s, q));
        FSIM_CUDA(cudaGetLastError());
        s->launches += 5;
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    s->binned = false;  // perm[] was in id space
    s->keys_valid = false;
    s->ever_sorted = true;
    s->ids_identity = false;
    s->steps_since_sort = 0;
    s->resort_due = false;
    return FSIM_OK;
}

int fsim_set_position(fsim_sim *s, const double *pos)
{
    FSIM_TRY(check_n(s));
    s->binned = false;
    s->keys_valid = false;
    s->have_leavers = false;
    if (s->fresh && !s->slab && s->n > 0) {
        s->fresh = false;
        return finish(s, ingest_sorted(s, pos));
    }
    s->steps_since_sort = 1 << 20;  // the storage order says nothing about the new positions: re-sort at the next density()
    return finish(s, particles_in3(s, pos, AX, s->factor_r, s->factor_r, s->factor_z, true));
}
int fsim_set_velocity(fsim_sim *s, const double *vel)
{
    FSIM_TRY(check_n(s));
    s->fresh = false;
    return finish(s, particles_in3(s, vel, AVX, s->factor_r, s->factor_r, s->factor_z, false));
}
int fsim_set_sink_mask(fsim_sim *s, const double *mask)
{
    FSIM_TRY(check(s));
    if (!mask) return fail(FSIM_ERR_INVALID, "null array");
    FSIM_TRY(finish(s, stage_in(s, mask, sizeof(double) * s->ncell_global)));
    sink_in_kernel<<<grid_for(s->ncell_global, 256), 256, 0, s->stream>>>((const double *)s->stage, s->sink,
                                                                         s->nr, s->nz);
    s->launches++;
    FSIM_CUDA(cudaGetLastError());
    return FSIM_OK;
}
int fsim_set_source_pdf(fsim_sim *s, const double *pdf, int64_t n0, int64_t n1)
{
    FSIM_TRY(check(s));
    if (!pdf || n0 < 1 || n1 < 1) return fail(FSIM_ERR_INVALID, ".source_pdf <- empty");
    std::vector<double> tab;
    FSIM_TRY(build_inv_cdf(pdf, n0, n1, tab));
    return finish(s, table_in(s, tab.data(), s->invcdf, (int64_t)tab.size()));
}
int fsim_set_inv_cdf(fsim_sim *s, const double *table)
{
    FSIM_TRY(check(s));
    return finish(s, table_in(s, table, s->invcdf, 2ll * FSIM_N_INVCDF * FSIM_N_INVCDF));
}
int fsim_set_entropy(fsim_sim *s, const double *entropy)
{
    FSIM_TRY(check(s));
    return finish(s, table_in(s, entropy, s->entropy, 4ll * FSIM_N_ENTROPY * FSIM_N_ENTROPY));
}
int fsim_set_rand(fsim_sim *s, const double *rnd)
{
    FSIM_TRY(check_n(s));
    if (!rnd) return fail(FSIM_ERR_INVALID, "null array");
    s->fresh = false;
    s->rand_default = false;
    if (s->n == 0) return FSIM_OK;
    FSIM_TRY(finish(s, stage_in(s, rnd, sizeof(double) * 4 * s->n)));
    return finish(s, dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        part4_in_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
            (const double *)s->stage, (Real *)s->part[s->cur][AQ0], (Real *)s->part[s->cur][AQ1],
            (Real *)s->part[s->cur][AQ2], (Real *)s->part[s->cur][AQ3], s->pid[s->cur], s->id_base, s->n,
            s->slab ? 0 : 1);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    }));
}
// Checkpoint restore (extension): the exact inverses of fsim_get_position / fsim_get_velocity /
// fsim_get_rand -- normalised units, alive flags included -- so that a run can be resumed, or started
// from a state no set() call can express (particles that were "just respawned").
int fsim_set_state(fsim_sim *s, const double *pos4, const double *vel3, const double *rand4)
{
    FSIM_TRY(check_n(s));
    s->fresh = false;
    if (s->n == 0) return FSIM_OK;
    if (pos4) {
        FSIM_TRY(finish(s, stage_in(s, pos4, sizeof(double) * 4 * s->n)));
        FSIM_TRY(finish(s, dispatch(s, [&](auto tag) {
            using Real = decltype(tag);
            const int c = s->cur;
            state_pos_in_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
                (const double *)s->stage, (Real *)s->part[c][AX], (Real *)s->part[c][AY], (Real *)s->part[c][AZ],
                s->alive[c], s->pid[c], s->id_base, s->n, s->slab ? 0 : 1);
            FSIM_CUDA(cudaGetLastError());
            s->launches++;
            return (int)FSIM_OK;
        })));
        s->binned = false;
        s->keys_valid = false;
        s->have_leavers = false;
        s->steps_since_sort = 1 << 20;  // re-sort at the next density()
    }
    if (vel3) FSIM_TRY(finish(s, particles_in3(s, vel3, AVX, 1.0, 1.0, 1.0, false)));
    if (rand4) FSIM_TRY(fsim_set_rand(s, rand4));
    if (vel3 && s->keys_valid) s->keys_valid = false;  // the sprite colours carry the velocity
    return FSIM_OK;
}

// "moments01_avg" [cells][4] (the running average density() blends into) or "phi" [cells]
int fsim_set_field(fsim_sim *s, const char *name, const double *data)
{
    FSIM_TRY(check(s));
    if (!name || !data) return fail(FSIM_ERR_INVALID, "null argument");
    const std::string n(name);
    const int64_t nc = s->ncell_local;
    int nch = 0;
    void *dst = nullptr;
    if (n == "moments01_avg") { nch = 4; dst = s->avg; }
    else if (n == "phi" || n == "background") {
        FSIM_TRY(finish(s, ensure_fieldsolve(s)));
        nch = 1; dst = n == "phi" ? s->phi[s->phi_cur] : s->background;
    } else return fail(FSIM_ERR_INVALID, "set field: unknown name " + n);
    FSIM_TRY(finish(s, stage_in(s, data, sizeof(double) * nch * nc)));
    return finish(s, dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        planar_in_kernel<Real><<<grid_for(nc, 256), 256, 0, s->stream>>>((const double *)s->stage, (Real *)dst, s->nr,
                                                                        s->rows, s->pitch, s->plane, nch);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    }));
}

int fsim_set_particle_count(fsim_sim *s, int64_t n)
{
    FSIM_TRY(check(s));
    if (n < 0 || n > s->cap - 1024) return fail(FSIM_ERR_RANGE, "particle count exceeds capacity");
    s->n = n;
    s->fresh = true;         // a re-initialisation: no per-particle state that a sort on ingest would have to carry along
    s->rand_default = true;  // (the default rand is re-drawn by particle id when the positions come in)
    if (s->n_async) {  // the device-resident count follows (pageable source: staged before the call returns)
        const uint32_t n32 = (uint32_t)n;
        FSIM_CUDA(cudaMemcpyAsync(s->mscratch + MC_NLIVE, &n32, sizeof n32, cudaMemcpyHostToDevice, s->stream));
        s->n_inflight[0] = s->n_inflight[1] = false;
    }
    s->binned = false;
    s->keys_valid = false;
    s->have_leavers = false;
    // slots get fresh ids id_base + slot (a re-initialisation; fsim_set_ids may override)
    if (n) {
        iota_ids_kernel<double><<<grid_for(n, 256), 256, 0, s->stream>>>(s->pid[s->cur], n, s->id_base);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
    }
    return FSIM_OK;
}
int fsim_set_ids(fsim_sim *s, const uint64_t *ids)
{
    FSIM_TRY(check_n(s));
    s->fresh = false;  // the caller's ids stay where they are
    if (!ids) return fail(FSIM_ERR_INVALID, "null array");
    std::vector<uint32_t> tmp((size_t)s->n);
    for (int64_t k = 0; k < s->n; ++k) {
        // outside slab mode the accessors place particle `id` at row id - id_base of the caller's arrays
        if (!s->slab && (ids[k] < s->id_base || ids[k] >= (uint64_t)s->id_base + (uint64_t)s->n))
            return fail(FSIM_ERR_RANGE, ".ids <- outside [id_base, id_base + N): only a slab rank holds arbitrary global ids");
        tmp[k] = (uint32_t)ids[k];
    }
    FSIM_CUDA(cudaMemcpyAsync(s->pid[s->cur], tmp.data(), sizeof(uint32_t) * s->n, cudaMemcpyHostToDevice, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    return FSIM_OK;
}

int fsim_add_current_loop(fsim_sim *s, double r, double z, double I)
{
    FSIM_TRY(check(s));
    s->bmag_valid = false;  // the |B| layer of the canvas is stale
    // u_R = r*factor_r, u_Z = z*factor_z, empic.js:1355-1357
    return finish(s, launch_add_loop(s, r * s->factor_r, z * s->factor_z, I));
}
int fsim_add_current_z(fsim_sim *s, double I)
{
    FSIM_TRY(check(s));
    s->bmag_valid = false;  // the |B| layer of the canvas is stale
    return finish(s, launch_add_uniform(s, 0, I));
}
int fsim_add_bz(fsim_sim *s, double Bz)
{
    FSIM_TRY(check(s));
    s->bmag_valid = false;  // the |B| layer of the canvas is stale
    return finish(s, launch_add_uniform(s, 1, Bz));
}
int fsim_add_btheta(fsim_sim *s, double Bt)
{
    FSIM_TRY(check(s));
    s->bmag_valid = false;  // the |B| layer of the canvas is stale
    return finish(s, launch_add_uniform(s, 2, Bt));
}
int fsim_add_spindle_cusp_plasma_field(fsim_sim *s, double r, double B_c, double beta_c)
{
    FSIM_TRY(check(s));
    s->bmag_valid = false;  // the |B| layer of the canvas is stale
    // spindle.makeSpindleCuspPlasmaField does not run in the reference (spindle.js:57,328,333,624,643,651); this is
    // the boundary solve it was written towards, specified in include/fusionsim.h (spindle.cu)
    auto bad = [](double v) { return !(v == v) || isinf(v); };
    if (bad(r) || r <= 0) return fail(FSIM_ERR_INVALID, ".r <- coil radius must be a positive number");
    if (bad(B_c)) return fail(FSIM_ERR_INVALID, ".B_c <- must be a finite number");
    if (bad(beta_c) || beta_c < 0 || beta_c > 1) return fail(FSIM_ERR_INVALID, ".beta_c <- must lie in [0, 1]");
    return finish(s, spindle_solve(s, r, B_c, beta_c));
}
// what the last fsim_add_spindle_cusp_plasma_field found: element strengths x [256], node currents [257] (amperes),
// the solved system A [256][256] and rhs [256]; checks of the solver and its final `diff` (matrix_webgl.js:687)
int fsim_get_spindle(fsim_sim *s, double *x, double *currents, double *A, double *rhs, int32_t *iterations, double *diff)
{
    FSIM_TRY(check_ro(s));
    if (s->spindle_x.empty()) return fail(FSIM_ERR_STATE, "no spindle-cusp boundary solve on this handle yet");
    if (x) memcpy(x, s->spindle_x.data(), sizeof(double) * s->spindle_x.size());
    if (currents) memcpy(currents, s->spindle_currents.data(), sizeof(double) * s->spindle_currents.size());
    if (A) memcpy(A, s->spindle_A.data(), sizeof(double) * s->spindle_A.size());
    if (rhs) memcpy(rhs, s->spindle_rhs.data(), sizeof(double) * s->spindle_rhs.size());
    if (iterations) *iterations = s->spindle_iterations;
    if (diff) *diff = s->spindle_diff;
    return FSIM_OK;
}

int fsim_precalc(fsim_sim *s)
{
    FSIM_TRY(check(s));
    FSIM_TRY(finish(s, launch_precalc(s)));
    s->have_precalc = true;
    return FSIM_OK;
}

// EXTENSION: charge density from the deposited moments -> weighted-Jacobi sweeps on
// the potential (warm start) -> E = -grad(phi) -> precalc().  Specification:
// include/fusionsim.h (fsim_solve_fields).
static int solve_args_ok(fsim_sim *s, int32_t sweeps, double omega, int32_t source)
{
    if (sweeps < 0 || !(omega > 0.0 && omega < 2.0)) return fail(FSIM_ERR_INVALID, ".sweeps/.omega <- out of range");
    if (source != 0 && source != 1) return fail(FSIM_ERR_INVALID, ".source <- 0 (running average) or 1 (instantaneous)");
    if (source == 1 && !s->norm)
        return fail(FSIM_ERR_STATE, "solveFields: the instantaneous density needs FSIM_FLAG_KEEP_MOMENTS");
    return FSIM_OK;
}

static int solve_charge_source(fsim_sim *s, double macro_weight, int32_t source)
{
    const double dr = s->spec.radius / (double)s->nr, dz = s->spec.height / (double)s->nz;
    const double rho_scale = s->spec.particle_charge * macro_weight / (FSIM_PI * s->spec.radius * dr * dz * FSIM_EPS0);
    const char *dens = (const char *)(source ? s->norm : s->avg) + s->rs * 3 * (size_t)s->plane;  // channel a
    return launch_charge_source(s, dens, rho_scale);
}

int fsim_solve_fields(fsim_sim *s, double macro_weight, int32_t sweeps, double omega, int32_t source)
{
    FSIM_TRY(check(s));
    if (s->slab)
        return fail(FSIM_ERR_UNSUPPORTED, "solveFields on a slab: drive fsim_solve_fields_stage with halo exchanges "
                                          "between the stages (fusion_sim_b200/dist.py)");
    FSIM_TRY(solve_args_ok(s, sweeps, omega, source));
    FSIM_TRY(finish(s, ensure_fieldsolve(s)));
    FSIM_TRY(finish(s, solve_charge_source(s, macro_weight, source)));
    // periodic z: what a slab rank exchanges with its neighbours between the stages (dist.py, solve_fields_slab) this
    // handle exchanges with itself -- 4 rows of the source once, 4 rows of the potential before every launch of <= 4 sweeps
    auto wrap = [&](void *plane) { return s->ring ? finish(s, ring_wrap_rows(s, plane, 4)) : (int)FSIM_OK; };
    FSIM_TRY(wrap(s->rho_src));
    int left = sweeps;
    for (; left >= 4; left -= 4) { FSIM_TRY(wrap(s->phi[s->phi_cur])); FSIM_TRY(finish(s, launch_relax(s, 4, omega))); }
    if (left >= 2) { FSIM_TRY(wrap(s->phi[s->phi_cur])); FSIM_TRY(finish(s, launch_relax(s, 2, omega))); left -= 2; }
    if (left >= 1) { FSIM_TRY(wrap(s->phi[s->phi_cur])); FSIM_TRY(finish(s, launch_relax(s, 1, omega))); }
    FSIM_TRY(wrap(s->phi[s->phi_cur]));
    FSIM_TRY(finish(s, launch_efield(s)));
    FSIM_TRY(finish(s, launch_precalc(s)));
    s->have_precalc = true;
    return FSIM_OK;
}

// EXTENSION (BASELINE configs[2]): electromagnetic field update on an axisymmetric Yee mesh (em.cu).
// Specification: include/fusionsim.h (fsim_em_step).
static int em_ready(fsim_sim *s)
{
    if (!s->em_on) return fail(FSIM_ERR_STATE, "no electromagnetic fields on this handle: call fsim_em_init first");
    return FSIM_OK;
}
int fsim_em_init(fsim_sim *s)
{
    FSIM_TRY(check(s));
    if (s->slab || s->ring)
        return fail(FSIM_ERR_UNSUPPORTED, "emInit: the electromagnetic update runs on one GPU holding the whole, non-periodic grid");
    FSIM_TRY(finish(s, em_init(s)));
    s->em_on = true;
    return FSIM_OK;
}
int fsim_em_set(fsim_sim *s, const char *name, const double *data)
{
    FSIM_TRY(check(s));
    FSIM_TRY(em_ready(s));
    if (!name || !data) return fail(FSIM_ERR_INVALID, "null argument");
    const int f = em_field_index(name);
    if (f < 0) return fail(FSIM_ERR_INVALID, std::string("emSet: unknown field ") + name + " (Er Ez Bt Et Br Bz)");
    const int64_t cnt = em_field_count(s, f);
    return finish(s, table_in(s, data, s->em[f], cnt));
}
int fsim_em_get(fsim_sim *s, const char *name, double *out)
{
    FSIM_TRY(check_ro(s));
    FSIM_TRY(em_ready(s));
    if (!name || !out) return fail(FSIM_ERR_INVALID, "null argument");
    const int f = em_field_index(name);
    if (f < 0) return fail(FSIM_ERR_INVALID, std::string("emGet: unknown field ") + name + " (Er Ez Bt Et Br Bz)");
    return finish(s, table_out(s, s->em[f], out, em_field_count(s, f)));
}
int fsim_em_step(fsim_sim *s, double macro_weight, int32_t with_current)
{
    FSIM_TRY(check(s));
    FSIM_TRY(em_ready(s));
    if (!(macro_weight == macro_weight) || isinf(macro_weight)) return fail(FSIM_ERR_INVALID, ".macro_weight <- must be a finite number");
    if (with_current && !s->mom)
        return fail(FSIM_ERR_STATE, "emStep: the current comes from moments01, kept only with FSIM_FLAG_KEEP_MOMENTS");
    FSIM_TRY(finish(s, em_step(s, macro_weight, with_current != 0)));
    s->bmag_valid = false;  // the |B| layer of the canvas is stale
    FSIM_TRY(finish(s, launch_precalc(s)));
    s->have_precalc = true;
    return FSIM_OK;
}

// slab mode: one stage at a time; the caller moves boundary rows between neighbouring ranks
int fsim_solve_fields_stage(fsim_sim *s, int32_t stage, double macro_weight, int32_t sweeps, double omega, int32_t source)
{
    FSIM_TRY(check(s));
    FSIM_TRY(finish(s, ensure_fieldsolve(s)));
    switch (stage) {
    case 0:
        FSIM_TRY(solve_args_ok(s, 0, omega, source));
        return finish(s, solve_charge_source(s, macro_weight, source));
    case 1:
        FSIM_TRY(solve_args_ok(s, sweeps, omega, 0));
        return finish(s, launch_relax(s, sweeps, omega));
    case 2:
        return finish(s, launch_efield(s));
    case 3:
        FSIM_TRY(finish(s, launch_precalc(s)));
        s->have_precalc = true;
        return FSIM_OK;
    default:
        return fail(FSIM_ERR_INVALID, "solve stage 0..3");
    }
}

// device address of rows [first_row, first_row + nrows) of the LOCAL table of "phi", "rho_src" or "E"
int fsim_field_rows(fsim_sim *s, const char *name, int64_t first_row, int64_t nrows, void **ptr, int64_t *nbytes)
{
    FSIM_TRY(check(s));
    if (!name || !ptr || !nbytes) return fail(FSIM_ERR_INVALID, "null argument");
    if (first_row < 0 || nrows < 0 || first_row + nrows > s->rows) return fail(FSIM_ERR_RANGE, "rows outside the local table");
    const std::string n(name);
    char *base = nullptr;
    size_t row_bytes = 0;
    if (n == "E") {
        base = (char *)s->E;
        row_bytes = s->rs * 3 * (size_t)s->nr;
    } else if (n == "phi" || n == "rho_src") {
        FSIM_TRY(finish(s, ensure_fieldsolve(s)));
        base = (char *)(n == "phi" ? s->phi[s->phi_cur] : s->rho_src);
        row_bytes = s->rs * (size_t)s->pitch;
    } else {
        return fail(FSIM_ERR_INVALID, "field rows: unknown name " + n);
    }
    *ptr = base + row_bytes * (size_t)first_row;
    *nbytes = (int64_t)(row_bytes * (size_t)nrows);
    return FSIM_OK;
}

// Physical re-sort of the particle storage every `sort_interval` frames (default 8): between
// re-sorts density() bins through a 4-byte index list and the push tolerates the slowly decaying
// order (particles move a fraction of a cell per half-step).
static int sort_interval(const fsim_sim *s) { return s->spec.sort_interval > 0 ? s->spec.sort_interval : 8; }

// keys/colours (prepass) if the push did not emit them, then scan + index scatter
static int bin_particles(fsim_sim *s)
{
    if (s->binned) return FSIM_OK;
    if (!s->keys_valid) FSIM_TRY(launch_keys(s));
    return launch_bin(s);
}

static int physical_sort(fsim_sim *s)
{
    FSIM_TRY(bin_particles(s));
    return launch_apply_perm(s);
}

int fsim_half_step(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    s->config_epoch++;  // leaves the step()/density() cycle a captured frame graph replays
    s->fresh = false;
    return finish(s, launch_push(s, false, 1));
}

int fsim_step(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    s->fresh = false;
    // out.step, empic.js:1436-1469: B-buffers then A-buffers = two half-steps.  The second one also
    // emits the deposit prepass (sort key, sprite colour, histogram) of the new state for the
    // density() that follows.
    // Both are done in ONE sweep over the particle storage (push.cu, NH = 2).
    // After set({position}) the storage order says nothing about the new positions: put the storage into
    // cell order BEFORE the sweep (an unsorted sweep costs 6.8 ms instead of 2.9 at 64 Mi particles).
    // The sort itself is FUSED into the sweep: the binning leaves the index list, the sweep reads through it and
    // writes the other copy of the storage in cell order (push.cu, PERM).
    if (s->steps_since_sort >= (1 << 20) && s->n) {
        FSIM_TRY(finish(s, bin_particles(s)));
        s->resort_due = true;
    }
    FSIM_TRY(finish(s, launch_push(s, true, 2, s->resort_due)));
    s->resort_due = false;  // (a binning that no longer matches the positions cannot be used: wait for the next one)
    s->steps_since_sort++;
    if (s->steps_since_sort >= 4 * sort_interval(s))  // push-only loops: keep the gather coherent
        FSIM_TRY(finish(s, physical_sort(s)));
    return FSIM_OK;
}

// Run every kernel of this handle on a stream the caller owns (a cudaStream_t), e.g. the stream a
// communication library orders its collectives on: the multi-GPU driver then needs no host
// synchronisation between a kernel and the collective that consumes its output.
int fsim_set_stream(fsim_sim *s, void *cuda_stream)
{
    FSIM_TRY(check(s));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    if (s->stream && !s->ext_stream) FSIM_CUDA(cudaStreamDestroy(s->stream));
    s->stream = (cudaStream_t)cuda_stream;
    s->ext_stream = true;
    return FSIM_OK;
}

int fsim_sort(fsim_sim *s)
{
    FSIM_TRY(check(s));
    return finish(s, physical_sort(s));
}

int fsim_density_begin(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    FSIM_TRY(finish(s, bin_particles(s)));
    if (s->spec.flags & FSIM_FLAG_ATOMIC_DEPOSIT)
        FSIM_TRY(finish(s, launch_cellsum_atomic(s)));  // measured alternative, not bit-reproducible
    else
        FSIM_TRY(finish(s, launch_cellsum(s)));
    if (s->slab || s->ring) FSIM_TRY(finish(s, launch_halo_pack(s)));  // own boundary rows -> send buffers (the caller's exchange starts here)
    // the deposit is done with the index list; now, every sort_interval frames, put the storage
    // itself into cell order for the pushes that follow
    // -- fused into the next step()'s sweep, which reads through this frame's index list (no pass of its own)
    if (!s->ever_sorted || s->steps_since_sort >= sort_interval(s)) {
        if (s->spec.flags & FSIM_FLAG_UNFUSED_SORT) FSIM_TRY(finish(s, launch_apply_perm(s)));  // measurement only
        else s->resort_due = true;
    }
    s->conv_interior_done = false;
    return FSIM_OK;
}
// slab mode: the stencil on the rows that need no halo row, to run WHILE the halo exchange is in flight
int fsim_density_interior(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    if (!s->slab) return FSIM_OK;
    FSIM_TRY(finish(s, launch_conv_rows(s, 1)));
    s->conv_interior_done = true;
    return FSIM_OK;
}
int fsim_density_end(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    if (s->ring) {  // periodic z: the exchange is with itself -- bottom rows above the top, top rows below the bottom
        void *p[4];
        int64_t each = 0;
        FSIM_TRY(fsim_halo_ptrs(s, &p[0], &p[1], &p[2], &p[3], &each));
        FSIM_CUDA(cudaMemcpyAsync(p[3], p[0], (size_t)each, cudaMemcpyDeviceToDevice, s->stream));
        FSIM_CUDA(cudaMemcpyAsync(p[2], p[1], (size_t)each, cudaMemcpyDeviceToDevice, s->stream));
    }
    if (s->slab || s->ring) FSIM_TRY(finish(s, launch_halo_unpack(s)));  // neighbours' boundary rows -> halo rows of the sums
    const int part = s->conv_interior_done ? 2 : 0;
    s->conv_interior_done = false;
    return finish(s, launch_conv_rows(s, part));
}
// device addresses of the per-cell sums (planar, 4 planes of rows x pitch reals) and the per-cell
// counts -- for a caller that reduces them across ranks between fsim_density_begin and _end
// (the replicated-table alternative of the multi-GPU driver)
int fsim_cellsum_ptrs(fsim_sim *s, void **sums, int64_t *sum_bytes, void **counts, int64_t *count_bytes)
{
    FSIM_TRY(check(s));
    if (!sums || !sum_bytes || !counts || !count_bytes) return fail(FSIM_ERR_INVALID, "null argument");
    *sums = s->cellsum;
    *sum_bytes = (int64_t)(s->rs * 4 * (size_t)s->plane);
    *counts = s->cellcount;
    *count_bytes = (int64_t)(sizeof(uint32_t) * (size_t)s->ncell_local);
    return FSIM_OK;
}

int fsim_density(fsim_sim *s)
{
    FSIM_TRY(fsim_density_begin(s));
    return fsim_density_end(s);
}

int fsim_render_rgba8(fsim_sim *s, uint8_t *rgba)
{
    FSIM_TRY(check_ro(s));
    if (!rgba) return fail(FSIM_ERR_INVALID, "null array");
    const size_t bytes = 4 * (size_t)s->ncell_global;
    FSIM_TRY(finish(s, ensure_stage(s, bytes)));
    FSIM_TRY(finish(s, launch_render(s, (uint8_t *)s->stage, s->stream)));
    // canvas rows run top-down (row nz-1-j): the owned rows [own0, own0+own_rows) are one block;
    // a slab rank fills only its block of the caller's full-size image.
    const size_t off = 4 * (size_t)s->nr * (size_t)(s->nz - s->own0 - s->own_rows);
    const size_t len = 4 * (size_t)s->nr * (size_t)s->own_rows;
    FSIM_CUDA(cudaMemcpyAsync(rgba + off, (uint8_t *)s->stage + off, len, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    return FSIM_OK;
}

// two device images alternate; the copy of image k (if any) runs on the copy stream under the next frame
static int canvas_draw(fsim_sim *s, uint8_t *host, bool own_rows_only)
{
    const size_t bytes = 4 * (size_t)s->ncell_global;
    if (!s->copy_stream) {
        FSIM_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            FSIM_CUDA(cudaMalloc((void **)&s->canvas_dev[k], bytes));
            FSIM_CUDA(cudaEventCreateWithFlags(&s->render_done[k], cudaEventDisableTiming));
            FSIM_CUDA(cudaEventCreateWithFlags(&s->copy_done[k], cudaEventDisableTiming));
        }
    }
    const int k = (s->canvas_slot ^= 1);
    if (s->copy_pending[k]) FSIM_CUDA(cudaEventSynchronize(s->copy_done[k]));  // image k is free again
    s->copy_pending[k] = false;
    const cudaStream_t ps = post_begin(s);  // after the stencil, on its stream: the next sweep does not wait for the image
    FSIM_TRY(finish(s, launch_render(s, s->canvas_dev[k], ps)));
    FSIM_TRY(post_end(s));
    if (!host) return FSIM_OK;
    FSIM_CUDA(cudaEventRecord(s->render_done[k], ps));
    FSIM_CUDA(cudaStreamWaitEvent(s->copy_stream, s->render_done[k], 0));
    const size_t off = 4 * (size_t)s->nr * (size_t)(s->nz - s->own0 - s->own_rows);
    const size_t len = 4 * (size_t)s->nr * (size_t)s->own_rows;
    FSIM_CUDA(cudaMemcpyAsync(host + (own_rows_only ? 0 : off), s->canvas_dev[k] + off, len, cudaMemcpyDeviceToHost, s->copy_stream));
    FSIM_CUDA(cudaEventRecord(s->copy_done[k], s->copy_stream));
    s->copy_pending[k] = true;
    return FSIM_OK;
}

int fsim_render_rgba8_async(fsim_sim *s, uint8_t *rgba)
{
    FSIM_TRY(check_handle(s));
    if (!rgba) return fail(FSIM_ERR_INVALID, "null array");
    return canvas_draw(s, rgba, false);
}
// slab mode: `rows` holds only this rank's rows [own_rows][nr][4] (top row first) -- small enough to pin
int fsim_render_rows_async(fsim_sim *s, uint8_t *rows)
{
    FSIM_TRY(check_handle(s));
    if (!rows) return fail(FSIM_ERR_INVALID, "null array");
    return canvas_draw(s, rows, true);
}
// the two canvas draws of out.density (programBMag + programDensity, empic.js:1497-1504) into the
// device-resident canvas, without a read-back: where the reference leaves its canvas, too
int fsim_draw_canvas(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    return canvas_draw(s, nullptr, false);
}

// ---- N frames of the page loop, launch-bound scenes through a CUDA graph ------------------------------
// The reference's page runs step(); density() once per animation frame (fusionsim.js:170-178).  On the
// reference's own demo scene (160 000 particles) a frame is ~14 launches of 3-30 us each: the GPU idles between
// them.  fsim_run_frames(n) = n x (step, density, the two canvas draws), bit-identical to calling them one by
// one; after one cycle of frames launched normally it CAPTURES the next cycle -- 2 x sort_interval frames: the
// re-sort fused into every sort_interval-th sweep flips the two copies of the particle storage, so the kernel
// arguments repeat with that period -- into a CUDA graph and replays it while whole cycles remain.  The host-side
// frame state (which copy is current, frames since the re-sort, canvas image, ...) returns to its starting value
// after a cycle, which is checked at capture; any entry point that can change what a frame launches bumps
// config_epoch and the graph is dropped.
static void frame_phase(const fsim_sim *s, int (&ph)[12])
{
    const int v[12] = {s->cur, s->steps_since_sort, s->canvas_slot, s->resort_due, s->ever_sorted, s->keys_valid, s->binned,
                       s->counts_dirty, s->bmag_valid, s->conv_interior_done, s->tm_sums_rows, (int)s->fresh};
    for (int k = 0; k < 12; ++k) ph[k] = v[k];
}
static void frame_phase_restore(fsim_sim *s, const int (&ph)[12])
{
    s->cur = ph[0]; s->steps_since_sort = ph[1]; s->canvas_slot = ph[2]; s->resort_due = ph[3]; s->ever_sorted = ph[4];
    s->keys_valid = ph[5]; s->binned = ph[6]; s->counts_dirty = ph[7]; s->bmag_valid = ph[8]; s->conv_interior_done = ph[9];
    s->tm_sums_rows = ph[10]; s->fresh = ph[11];
}
static bool frame_phase_is(const fsim_sim *s, const int (&want)[12])
{
    int ph[12];
    frame_phase(s, ph);
    for (int k = 0; k < 12; ++k)
        if (ph[k] != want[k]) return false;
    return true;
}
static void drop_frame_graph(fsim_sim *s)
{
    if (s->frame_graph) cudaGraphExecDestroy(s->frame_graph);
    s->frame_graph = nullptr;
}
static int one_frame(fsim_sim *s)
{
    FSIM_TRY(fsim_step(s));
    FSIM_TRY(fsim_density(s));
    return fsim_draw_canvas(s);
}
// capture one cycle; on any failure the host-side state is put back and frames go on one by one
static int capture_frame_graph(fsim_sim *s, int period)
{
    if (s->copy_stream && (s->copy_pending[0] || s->copy_pending[1])) {  // no read-back may be in flight on a canvas image
        FSIM_CUDA(cudaStreamSynchronize(s->copy_stream));
        s->copy_pending[0] = s->copy_pending[1] = false;
    }
    int ph0[12];
    frame_phase(s, ph0);
    const int64_t launches0 = s->launches;
    std::map<std::string, int64_t> timer0;
    for (auto &kv : s->timers) timer0[kv.first] = kv.second.launches;
    if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        s->graph_failed = true;
        return FSIM_OK;
    }
    int rc = FSIM_OK;
    for (int k = 0; k < period && rc == FSIM_OK; ++k) rc = one_frame(s);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s->stream, &g);
    const bool cyclic = frame_phase_is(s, ph0);
    const int64_t per_cycle = s->launches - launches0;
    // nothing of the above ran: the host state goes back to where the device is
    frame_phase_restore(s, ph0);
    s->launches = launches0;
    for (auto &kv : s->timers) kv.second.launches = timer0.count(kv.first) ? timer0[kv.first] : 0;
    if (rc != FSIM_OK || e != cudaSuccess || !g || !cyclic) {
        cudaGetLastError();
        if (g) cudaGraphDestroy(g);
        s->sticky_error = false;  // a launch refused under capture is not a device fault
        s->graph_failed = true;
        return FSIM_OK;
    }
    cudaGraphExec_t ex = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess || !ex) {
        cudaGetLastError();
        s->graph_failed = true;
        return FSIM_OK;
    }
    s->frame_graph = ex;
    s->graph_epoch = s->config_epoch;
    s->graph_frames = period;
    s->graph_launches = per_cycle;
    for (int k = 0; k < 12; ++k) s->graph_phase[k] = ph0[k];
    return FSIM_OK;
}

int fsim_run_frames(fsim_sim *s, int64_t nframes)
{
    FSIM_TRY(check_handle(s));
    if (nframes < 0) return fail(FSIM_ERR_INVALID, ".nframes <- must not be negative");
    if (s->slab) return fail(FSIM_ERR_UNSUPPORTED, "run_frames on a slab: the frame of a slab has exchanges between its parts "
                                                   "(fusion_sim_b200/dist.py drives them)");
    const int period = 2 * sort_interval(s);
    // a graph is worth it, and safe, on a plain handle: one stream, no per-launch timing events, no periodic self-exchange
    const bool allow = !s->ring && !s->timing && !s->ext_stream && !(s->spec.flags & FSIM_FLAG_POST_STREAM);
    if (s->frame_graph && (s->graph_epoch != s->config_epoch || !allow)) {
        drop_frame_graph(s);
        s->frames_run = 0;  // whatever changed may bring its own lazy initialisation: one cycle by hand first
    }
    int off_cycle = 0;
    while (nframes > 0) {
        if (allow && s->frame_graph && nframes >= s->graph_frames && frame_phase_is(s, s->graph_phase)) {
            // the replay draws into both canvas images: a read-back of one of them (fsim_render_rgba8_async) must have landed
            if (s->copy_stream && (s->copy_pending[0] || s->copy_pending[1])) {
                FSIM_CUDA(cudaStreamSynchronize(s->copy_stream));
                s->copy_pending[0] = s->copy_pending[1] = false;
            }
            FSIM_CUDA(cudaGraphLaunch(s->frame_graph, s->stream));
            s->launches += s->graph_launches;
            s->graph_replays++;
            nframes -= s->graph_frames;
            off_cycle = 0;
            continue;
        }
        if (allow && !s->frame_graph && !s->graph_failed && s->frames_run >= period && nframes >= period && s->ever_sorted &&
            s->steps_since_sort <= sort_interval(s)) {
            FSIM_TRY(capture_frame_graph(s, period));
            if (s->frame_graph) continue;
        }
        FSIM_TRY(one_frame(s));
        s->frames_run++;
        nframes--;
        // step() or density() called by hand between two run_frames shifts the cycle: the captured phase never comes back
        if (s->frame_graph && ++off_cycle > 2 * s->graph_frames) {
            drop_frame_graph(s);
            off_cycle = 0;
        }
    }
    return FSIM_OK;
}
// statistics of fsim_run_frames: frames per captured cycle (0: no graph), kernel launches per cycle, replays so far
int fsim_frame_graph_info(fsim_sim *s, int32_t *frames_per_cycle, int64_t *launches_per_cycle, int64_t *replays)
{
    FSIM_TRY(check_handle(s));
    if (frames_per_cycle) *frames_per_cycle = s->frame_graph ? s->graph_frames : 0;
    if (launches_per_cycle) *launches_per_cycle = s->frame_graph ? s->graph_launches : 0;
    if (replays) *replays = s->graph_replays;
    return FSIM_OK;
}

int fsim_sync(fsim_sim *s)
{
    FSIM_TRY(check_ro(s));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    if (s->copy_stream) {
        FSIM_CUDA(cudaStreamSynchronize(s->copy_stream));
        s->copy_pending[0] = s->copy_pending[1] = false;
    }
    uint32_t oob = 0, merr = 0;
    FSIM_CUDA(cudaMemcpyAsync(&oob, s->oob, sizeof oob, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaMemcpyAsync(&merr, s->mscratch + MC_ERR, sizeof merr, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    FSIM_TRY(settle_count(s));
    if (merr) {
        FSIM_CUDA(cudaMemsetAsync(s->mscratch + MC_ERR, 0, sizeof merr, s->stream));
        return fail(FSIM_ERR_RANGE, std::string("slab exchange: ") +
                    ((merr & MERR_SEND_OVERFLOW) ? "more particles left for one rank in a frame than its send region holds: the "
                                                   "stragglers stayed behind (they migrate in a later frame), so that frame's "
                                                   "density misses them -- raise the exchange capacity, or use the exact exchange; " : "") +
                    ((merr & MERR_CAPACITY) ? "arrivals exceeded the particle capacity of this rank and were DROPPED: the "
                                              "particle state is no longer valid; " : ""));
    }
    if (oob) {
        FSIM_CUDA(cudaMemsetAsync(s->oob, 0, sizeof oob, s->stream));  // ordered before the next push by the stream
        char buf[256];
        snprintf(buf, sizeof buf, "%u particle pushes gathered outside the local slab table (halo_rows too small "
                 "or migration overdue)", oob);
        return fail(FSIM_ERR_RANGE, buf);
    }
    return FSIM_OK;
}

#ifdef FSIM_TUNE
// tuning build only (make EXTRA=-DFSIM_TUNE; not declared in include/fusionsim.h, not in the product)
int fsim_tune_set(int push_variant, int conv_variant)
{
    fsim::g_push_variant = push_variant;
    fsim::g_conv_variant = conv_variant;
    return FSIM_OK;
}
#endif

int64_t fsim_particle_count(const fsim_sim *s)
{
    if (!s) return -1;
    if (s->n_async && settle_count(const_cast<fsim_sim *>(s)) != FSIM_OK) return -1;
    return s->n;
}
int64_t fsim_local_cells(const fsim_sim *s) { return s ? s->ncell_local : -1; }
int64_t fsim_launch_count(const fsim_sim *s) { return s ? s->launches : -1; }

static int part_out(fsim_sim *s, double *out, int width, int a0, bool with_alive)
{
    if (!out) return fail(FSIM_ERR_INVALID, "null array");
    if (s->n == 0) return FSIM_OK;
    FSIM_TRY(ensure_stage(s, sizeof(double) * width * s->n));
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int c = s->cur;
        part_out_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
            (double *)s->stage, width, (const Real *)s->part[c][a0], (const Real *)s->part[c][a0 + 1],
            (const Real *)s->part[c][a0 + 2], width == 4 && !with_alive ? (const Real *)s->part[c][a0 + 3] : nullptr,
            with_alive ? s->alive[c] : nullptr, s->pid[c], s->id_base, s->n, s->slab ? 0 : 1);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    return stage_out(s, out, sizeof(double) * width * s->n);
}

int fsim_get_position(fsim_sim *s, double *out)
{
    FSIM_TRY(check_n_ro(s));
    return finish(s, part_out(s, out, 4, AX, true));
}
int fsim_get_velocity(fsim_sim *s, double *out)
{
    FSIM_TRY(check_n_ro(s));
    return finish(s, part_out(s, out, 3, AVX, false));
}
int fsim_get_rand(fsim_sim *s, double *out)
{
    FSIM_TRY(check_n_ro(s));
    return finish(s, part_out(s, out, 4, AQ0, false));
}
int fsim_get_ids(fsim_sim *s, uint64_t *out)
{
    FSIM_TRY(check_n_ro(s));
    if (!out) return fail(FSIM_ERR_INVALID, "null array");
    std::vector<uint32_t> tmp((size_t)s->n);
    FSIM_CUDA(cudaMemcpyAsync(tmp.data(), s->pid[s->cur], sizeof(uint32_t) * s->n, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    for (int64_t k = 0; k < s->n; ++k) out[k] = tmp[k];
    return FSIM_OK;
}
int fsim_get_cells(fsim_sim *s, int64_t *out)
{
    FSIM_TRY(check_n_ro(s));
    if (!out) return fail(FSIM_ERR_INVALID, "null array");
    if (s->n == 0) return FSIM_OK;
    FSIM_TRY(finish(s, ensure_stage(s, sizeof(int64_t) * s->n)));
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int c = s->cur;
        cells_out_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
            (int64_t *)s->stage, (const Real *)s->part[c][AX], (const Real *)s->part[c][AY],
            (const Real *)s->part[c][AZ], s->pid[c], s->id_base, s->n, s->nr, s->nz, s->slab ? 0 : 1);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
    FSIM_TRY(finish(s, rc));
    return finish(s, stage_out(s, out, sizeof(int64_t) * s->n));
}

int fsim_get_field(fsim_sim *s, const char *name, double *out)
{
    FSIM_TRY(check_ro(s));
    if (!name) return fail(FSIM_ERR_INVALID, "null field name");
    if (!out) return fail(FSIM_ERR_INVALID, "null array");
    const std::string n(name);
    const int64_t nc = s->ncell_local;
    if (n == "E") return finish(s, table_out(s, s->E, out, 3 * nc));
    if (n == "B") return finish(s, table_out(s, s->B, out, 3 * nc));
    if (n == "R1" || n == "R2" || n == "R3" || n == "A") {
        std::vector<double> rec((size_t)(FSIM_CELLREC * nc));
        FSIM_TRY(finish(s, ensure_stage(s, sizeof(double) * FSIM_CELLREC * nc)));
        FSIM_TRY(finish(s, launch_expand_records(s, (double *)s->stage)));
        FSIM_TRY(finish(s, stage_out(s, rec.data(), sizeof(double) * FSIM_CELLREC * nc)));
        const int off = n == "R1" ? 0 : n == "R2" ? 3 : n == "R3" ? 6 : 9;
        for (int64_t c = 0; c < nc; ++c)
            for (int k = 0; k < 3; ++k) out[3 * c + k] = rec[FSIM_CELLREC * c + off + k];
        return FSIM_OK;
    }
    if (n == "cell_sums") return finish(s, planar_out(s, s->cellsum, out));
    if (n == "moments01_avg") return finish(s, planar_out(s, s->avg, out));
    if (n == "moments01" || n == "moments01_norm") {
        if (!s->mom) return fail(FSIM_ERR_STATE, "moments01/moments01_norm are kept only with FSIM_FLAG_KEEP_MOMENTS");
        return finish(s, planar_out(s, n == "moments01" ? s->mom : s->norm, out));
    }
    if (n == "phi" || n == "rho_src") {
        if (!s->phi[0]) return fail(FSIM_ERR_STATE, "phi/rho_src exist after the first solveFields()");
        FSIM_TRY(finish(s, ensure_stage(s, sizeof(double) * nc)));
        FSIM_TRY(finish(s, launch_plane_out(s, n == "phi" ? s->phi[s->phi_cur] : s->rho_src, (double *)s->stage)));
        return finish(s, stage_out(s, out, sizeof(double) * nc));
    }
    if (n == "inv_cdf") return finish(s, table_out(s, s->invcdf, out, 2ll * FSIM_N_INVCDF * FSIM_N_INVCDF));
    if (n == "entropy") return finish(s, table_out(s, s->entropy, out, 4ll * FSIM_N_ENTROPY * FSIM_N_ENTROPY));
    return fail(FSIM_ERR_INVALID, "unknown field name: " + n);
}
int fsim_get_cell_count(fsim_sim *s, uint32_t *out)
{
    FSIM_TRY(check_ro(s));
    if (!out) return fail(FSIM_ERR_INVALID, "null array");
    FSIM_CUDA(cudaMemcpyAsync(out, s->cellcount, sizeof(uint32_t) * s->ncell_local, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    return FSIM_OK;
}
int fsim_get_sink_mask(fsim_sim *s, uint8_t *out)
{
    FSIM_TRY(check_ro(s));
    if (!out) return fail(FSIM_ERR_INVALID, "null array");
    FSIM_TRY(finish(s, ensure_stage(s, (size_t)s->ncell_global)));
    sink_out_kernel<<<grid_for(s->ncell_global, 256), 256, 0, s->stream>>>(s->sink, (uint8_t *)s->stage, s->ncell_global);
    FSIM_CUDA(cudaGetLastError());
    s->launches++;
    return finish(s, stage_out(s, out, (size_t)s->ncell_global));
}

// Invariants of a run, reduced on the device (extension; bench.py prints them with every line):
// out[0] live particles, out[1] XOR of their ids, out[2] sum of their ids (mod 2^64), out[3] particles the
// last density() deposited on the owned rows; *sum_alpha = sum of the weight channel of the per-cell sums
// over the owned rows (= 0.001 * out[3] up to rounding).  Over all ranks of a slab run the first three
// must equal those of ids 0..N-1: no particle lost or duplicated by the migration.
int fsim_check_digest(fsim_sim *s, uint64_t *out, double *sum_alpha)
{
    FSIM_TRY(check_n_ro(s));
    if (!out) return fail(FSIM_ERR_INVALID, "null array");
    constexpr int NB = 256;
    FSIM_TRY(finish(s, ensure_stage(s, 4 * sizeof(unsigned long long) + NB * sizeof(double))));
    unsigned long long *d = (unsigned long long *)s->stage;
    double *partial = (double *)(d + 4);
    FSIM_CUDA(cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long) + NB * sizeof(double), s->stream));
    if (s->n) digest_ids_kernel<<<s->nsm * 4, 256, 0, s->stream>>>(s->pid[s->cur], s->n, nullptr, d);
    const int lo = s->own0 - s->row0;
    FSIM_TRY(finish(s, dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        digest_cells_kernel<Real><<<NB, 256, 0, s->stream>>>(s->cellcount, (const Real *)s->cellsum + 3 * s->plane, s->nr,
                                                              s->pitch, lo, s->own_rows, d, partial);
        FSIM_CUDA(cudaGetLastError());
        s->launches += 2;
        return (int)FSIM_OK;
    })));
    unsigned long long hd[4];
    double hp[NB];
    FSIM_CUDA(cudaMemcpyAsync(hd, d, sizeof hd, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaMemcpyAsync(hp, partial, sizeof hp, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    out[0] = (uint64_t)s->n; out[1] = hd[0]; out[2] = hd[1]; out[3] = hd[2];
    if (sum_alpha) {
        double t = 0.0;
        for (int k = 0; k < NB; ++k) t += hp[k];
        *sum_alpha = t;
    }
    return FSIM_OK;
}

int fsim_timing_enable(fsim_sim *s, int on)
{
    FSIM_TRY(check(s));
    FSIM_TRY(collect_timers(s));
    s->timing = on != 0;
    return FSIM_OK;
}
int fsim_timing_reset(fsim_sim *s)
{
    FSIM_TRY(check(s));
    FSIM_TRY(collect_timers(s));
    for (auto &kv : s->timers) {
        kv.second.ms = 0;
        kv.second.launches = 0;
    }
    return FSIM_OK;
}
int fsim_timing_get(fsim_sim *s, const char *name, double *ms, int64_t *launches)
{
    FSIM_TRY(check_ro(s));
    if (!name) return fail(FSIM_ERR_INVALID, "null kernel name");
    FSIM_TRY(collect_timers(s));
    auto it = s->timers.find(name);
    if (ms) *ms = it == s->timers.end() ? 0.0 : it->second.ms;
    if (launches) *launches = it == s->timers.end() ? 0 : it->second.launches;
    return FSIM_OK;
}

int fsim_mark(fsim_sim *s, int slot)
{
    FSIM_TRY(check_ro(s));
    if (slot < 0 || slot >= 16) return fail(FSIM_ERR_INVALID, "mark slot out of range");
    if (!s->marks[slot]) FSIM_CUDA(cudaEventCreate(&s->marks[slot]));
    FSIM_CUDA(cudaEventRecord(s->marks[slot], s->stream));
    return FSIM_OK;
}
int fsim_elapsed_ms(fsim_sim *s, int a, int b, double *ms)
{
    FSIM_TRY(check_ro(s));
    if (a < 0 || a >= 16 || b < 0 || b >= 16 || !s->marks[a] || !s->marks[b] || !ms)
        return fail(FSIM_ERR_INVALID, "elapsed: unrecorded mark");
    FSIM_CUDA(cudaEventSynchronize(s->marks[b]));
    float f = 0.f;
    FSIM_CUDA(cudaEventElapsedTime(&f, s->marks[a], s->marks[b]));
    *ms = f;
    return FSIM_OK;
}

}  // extern "C"
