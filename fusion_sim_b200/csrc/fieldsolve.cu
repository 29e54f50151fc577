// fieldsolve.cu -- EXTENSION (no reference counterpart): the
// self-consistent electrostatic field solve that closes the PIC loop (sm_100a).
//
// The reference pushes test particles in static fields: density() only feeds the display
// (empic.js:1471-1505).  fsim_solve_fields() turns the deposited density into a charge density,
// relaxes the potential with the reference's own iteration -- the weighted Jacobi
// x <- omega (R x + C) + (1 - omega) x of matrix_webgl.makeSORIterative (matrix_webgl.js:224-300,
// :392-420), here on the sparse 5-point cylindrical operator instead of a dense matrix --, takes
// E = -grad(phi) and re-runs precalc().  Specification (operation order, boundary rules):
// include/fusionsim.h, fsim_solve_fields; the tests hold the result bit-identical to a CPU restatement.
//
//   charge_source : src = rho_scale * density.a                          24 B/cell  (HBM-bound)
//   relax<T>      : T Jacobi sweeps per launch on TMA-staged shared-memory tiles (temporal
//                   blocking): two 2-D cp.async.bulk.tensor box loads (phi and src, tile + 4-cell
//                   halo, out-of-grid cells zero-filled by the TMA unit = the grounded wall), T
//                   sweeps ping-ponging between two shared-memory copies on a region that shrinks by
//                   one cell per sweep, one coalesced store of the tile.  Algorithmic bytes per
//                   launch: 24 B/cell (phi in, src in, phi out) for T sweeps, i.e. 6 B per
//                   cell-sweep at T = 4 against 24 for a sweep-per-launch stencil.
//   efield        : centred differences -> E [cell][3]                    32 B/cell  (HBM-bound)
#include <cuda.h>

#include "common.cuh"

namespace fsim {

constexpr int RT_I = 56, RT_J = 32;     // output tile
constexpr int RT_H = 4;                 // halo = most sweeps per launch
constexpr int RB_W = RT_I + 2 * RT_H;   // 64: box width = two 32-lane column chunks (x start and width are multiples of 16 bytes)
constexpr int RB_H = RT_J + 2 * RT_H;   // 40
constexpr int RT_WARPS = 8, RT_THREADS = 32 * RT_WARPS;
constexpr int RT_CHUNKS = RB_W / 32;

template <typename Real>
struct RelaxArgs {
    Real *out;               // planar [rows][pitch]
    const Real *coef;        // [nr][4] = cE cW cZ cB
    int nr, rows, pitch;
    Real omega, one_m;
};

__device__ __forceinline__ uint32_t fs_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Work split of a sweep: a warp owns a band of rows, a lane owns columns lane and lane + 32 and
// walks down its band with the centre and south values in registers: 3 shared-memory loads (east,
// west, the next north) + the source + 1 store per cell, consecutive lanes on consecutive words.
template <typename Real, int T>
__global__ void __launch_bounds__(RT_THREADS)
relax_kernel(const __grid_constant__ CUtensorMap tmPhi, const __grid_constant__ CUtensorMap tmSrc,
             const RelaxArgs<Real> a)
{
    static_assert(T >= 1 && T <= RT_H, "sweeps per launch");
    extern __shared__ __align__(128) unsigned char relax_smem[];
    Real *A = reinterpret_cast<Real *>(relax_smem);  // [RB_H][RB_W] phi, written by the TMA unit
    Real *S = A + RB_H * RB_W;                        // [RB_H][RB_W] src,  written by the TMA unit
    Real *B = S + RB_H * RB_W;                        // [RB_H][RB_W] ping-pong partner of A
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * RT_I - RT_H, j0 = blockIdx.y * RT_J - RT_H;  // grid coordinates of box cell (0,0)
    constexpr uint32_t kBytes = 2u * RB_H * RB_W * sizeof(Real);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fs_smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fs_smem_u32(&bar)), "r"(kBytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(fs_smem_u32(A)), "l"(reinterpret_cast<unsigned long long>(&tmPhi)), "r"(i0), "r"(j0),
            "r"(fs_smem_u32(&bar))
            : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(fs_smem_u32(S)), "l"(reinterpret_cast<unsigned long long>(&tmSrc)), "r"(i0), "r"(j0),
            "r"(fs_smem_u32(&bar))
            : "memory");
    }
    // while the boxes fly: the coefficients of this lane's columns (registers)
    Real kE[RT_CHUNKS], kW[RT_CHUNKS], kZ[RT_CHUNKS], kB[RT_CHUNKS];
    bool col_in[RT_CHUNKS];
#pragma unroll
    for (int q = 0; q < RT_CHUNKS; ++q) {
        const int gi = i0 + lane + 32 * q;
        col_in[q] = gi >= 0 && gi < a.nr;
        const Real *k = a.coef + 4 * (size_t)(col_in[q] ? gi : 0);
        kE[q] = k[0]; kW[q] = k[1]; kZ[q] = k[2]; kB[q] = k[3];
    }
    {
        uint32_t done = 0;
        for (uint32_t spin = 0; !done; ++spin) {
            if (spin > (1u << 24)) __trap();  // a TMA that never completes must not hang the GPU
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(fs_smem_u32(&bar)), "r"(0u)
                : "memory");
        }
    }
    // cB * src once (the same product every sweep would form)
#pragma unroll
    for (int q = 0; q < RT_CHUNKS; ++q)
        for (int r = warp; r < RB_H; r += RT_WARPS) S[r * RB_W + lane + 32 * q] = kB[q] * S[r * RB_W + lane + 32 * q];
    __syncthreads();

    Real *in = A, *out = B;
#pragma unroll
    for (int s = 1; s <= T; ++s) {
        const int m = RT_H - T + s;                 // margin of the region this sweep can still compute
        const int h = RB_H - 2 * m;
        const int per = (h + RT_WARPS - 1) / RT_WARPS;
        const int r0 = m + warp * per, r1 = min(r0 + per, RB_H - m);
#pragma unroll
        for (int q = 0; q < RT_CHUNKS; ++q) {
            const int c = lane + 32 * q;
            if (c < m || c >= RB_W - m || r0 >= r1) continue;
            const Real *p = in + r0 * RB_W + c;
            Real pS = p[-RB_W], pC = p[0];
            for (int r = r0; r < r1; ++r, p += RB_W) {
                const Real pN = p[RB_W];
                Real v = (Real)0;  // ghost cells keep phi = 0
                const int gj = j0 + r;
                if (col_in[q] && gj >= 0 && gj < a.rows) {
                    const Real t = ((kE[q] * p[1] + kW[q] * p[-1]) + kZ[q] * (pN + pS)) + S[r * RB_W + c];
                    v = a.omega * t + a.one_m * pC;
                }
                out[r * RB_W + c] = v;
                pS = pC; pC = pN;
            }
        }
        __syncthreads();
        Real *sw = in; in = out; out = sw;
    }
    // `in` holds the result of the last sweep; store the tile
    for (int k = tid; k < RT_I * RT_J; k += RT_THREADS) {
        const int r = RT_H + k / RT_I, c = RT_H + k % RT_I;
        const int gi = i0 + c, gj = j0 + r;
        if (gi < a.nr && gj < a.rows) a.out[(size_t)gj * a.pitch + gi] = in[r * RB_W + c];
    }
}

template <typename Real>
__global__ void __launch_bounds__(256)
charge_source_kernel(const Real *__restrict__ dens_a, const Real *__restrict__ background, Real *__restrict__ src, int nr,
                     int rows, int pitch, Real scale)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)nr * rows) return;
    const size_t o = (size_t)(c / nr) * pitch + (size_t)(c % nr);
    src[o] = scale * (dens_a[o] - background[o]);  // background is zero until set: x - 0 = x, bit for bit
}

template <typename Real>
__global__ void __launch_bounds__(256)
efield_kernel(const Real *__restrict__ phi, Real *__restrict__ E, int nr, int rows, int pitch, Real inv2dr, Real inv2dz)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)nr * rows) return;
    const int i = (int)(c % nr), j = (int)(c / nr);
    const Real *p = phi + (size_t)j * pitch + i;
    const Real pE = (i + 1 < nr) ? p[1] : (Real)0;
    const Real pW = (i > 0) ? p[-1] : p[0];
    const Real pN = (j + 1 < rows) ? p[pitch] : (Real)0;
    const Real pS = (j > 0) ? p[-pitch] : (Real)0;
    E[3 * c] = -((pE - pW) * inv2dr);
    E[3 * c + 1] = (Real)0;
    E[3 * c + 2] = -((pN - pS) * inv2dz);
}

// planar single-channel field -> [cell] doubles (accessor)
template <typename Real>
__global__ void __launch_bounds__(256)
plane_out_kernel(const Real *__restrict__ in, double *__restrict__ out, int nr, int rows, int pitch)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)nr * rows) return;
    out[c] = (double)in[(size_t)(c / nr) * pitch + (size_t)(c % nr)];
}

int launch_plane_out(fsim_sim *s, const void *plane, double *dev_out)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        plane_out_kernel<Real><<<grid_for(s->ncell_local, 256), 256, 0, s->stream>>>((const Real *)plane, dev_out, s->nr,
                                                                                    s->rows, s->pitch);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
}

static int encode_plane_map(fsim_sim *s, void *base, unsigned char *storage)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    FSIM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return FSIM_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)s->nr, (cuuint64_t)s->rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)s->pitch * s->rs};
    const cuuint32_t box[2] = {RB_W, RB_H};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = ((EncodeFn)fn)(reinterpret_cast<CUtensorMap *>(storage),
                                      s->prec == FSIM_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                      2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (potential) failed with CUresult " + std::to_string((int)r));
        return FSIM_ERR_CUDA;
    }
    return FSIM_OK;
}

// device buffers of the solve, allocated at the first fsim_solve_fields()
int ensure_fieldsolve(fsim_sim *s)
{
    if (s->phi[0]) return FSIM_OK;
    const size_t bytes = s->rs * (size_t)s->plane;
    for (int k = 0; k < 2; ++k) {
        FSIM_CUDA(cudaMalloc(&s->phi[k], bytes));
        FSIM_CUDA(cudaMemsetAsync(s->phi[k], 0, bytes, s->stream));
        FSIM_TRY(encode_plane_map(s, s->phi[k], s->tm_phi[k]));
    }
    FSIM_CUDA(cudaMalloc(&s->rho_src, bytes));
    FSIM_CUDA(cudaMemsetAsync(s->rho_src, 0, bytes, s->stream));
    FSIM_CUDA(cudaMalloc(&s->background, bytes));
    FSIM_CUDA(cudaMemsetAsync(s->background, 0, bytes, s->stream));
    FSIM_TRY(encode_plane_map(s, s->rho_src, s->tm_src));
    FSIM_CUDA(cudaMalloc(&s->relax_coef, s->rs * 4 * (size_t)s->nr));
    // per-column coefficients in host fp64 (specification: include/fusionsim.h)
    const double dr = s->spec.radius / (double)s->nr, dz = s->spec.height / (double)s->nz;
    std::vector<double> cd(4 * (size_t)s->nr);
    for (int i = 0; i < s->nr; ++i) {
        const double rc = ((double)i + 0.5) * dr * dr;
        const double aE = ((double)i + 1.0) / rc, aW = (double)i / rc, aZ = 1.0 / (dz * dz);
        const double aC = aE + aW + 2.0 * aZ;
        cd[4 * i] = aE / aC; cd[4 * i + 1] = aW / aC; cd[4 * i + 2] = aZ / aC; cd[4 * i + 3] = 1.0 / aC;
    }
    if (s->prec == FSIM_F64) {
        FSIM_CUDA(cudaMemcpyAsync(s->relax_coef, cd.data(), sizeof(double) * cd.size(), cudaMemcpyHostToDevice, s->stream));
    } else {
        std::vector<float> cf(cd.begin(), cd.end());
        FSIM_CUDA(cudaMemcpyAsync(s->relax_coef, cf.data(), sizeof(float) * cf.size(), cudaMemcpyHostToDevice, s->stream));
    }
    FSIM_CUDA(cudaStreamSynchronize(s->stream));  // the host vectors go out of scope
    return FSIM_OK;
}

template <typename Real, int T>
static int relax_launch(fsim_sim *s, Real omega)
{
    RelaxArgs<Real> a;
    const int cur = s->phi_cur;
    a.out = (Real *)s->phi[cur ^ 1];
    a.coef = (const Real *)s->relax_coef;
    a.nr = s->nr; a.rows = s->rows; a.pitch = s->pitch;
    a.omega = omega; a.one_m = (Real)1.0 - omega;
    const size_t smem = sizeof(Real) * 3 * RB_H * RB_W;
    constexpr uint32_t bit = 1u << T;  // per handle, hence per device: the attribute belongs to the device's context
    if (!(s->smem_opt_in & bit)) {
        FSIM_CUDA(cudaFuncSetAttribute(relax_kernel<Real, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s->smem_opt_in |= bit;
    }
    dim3 grid((s->nr + RT_I - 1) / RT_I, (s->rows + RT_J - 1) / RT_J);
    Bracket b(s, T == 4 ? "relax4" : (T == 3 ? "relax3" : (T == 2 ? "relax2" : "relax1")));
    relax_kernel<Real, T><<<grid, RT_THREADS, smem, s->stream>>>(
        *reinterpret_cast<const CUtensorMap *>(s->tm_phi[cur]), *reinterpret_cast<const CUtensorMap *>(s->tm_src), a);
    FSIM_CUDA(cudaGetLastError());
    s->phi_cur = cur ^ 1;
    return FSIM_OK;
}

// The stages of a solve.  One GPU runs them back to back (api.cu, fsim_solve_fields); in slab mode
// the caller exchanges boundary rows between them (fusion_sim_b200/dist.py): every stage works on
// ALL local rows, rows further than the halo from an owned row simply hold unused values.
int launch_charge_source(fsim_sim *s, const void *dens_a, double rho_scale)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        Bracket b(s, "charge_source");
        charge_source_kernel<Real><<<grid_for(s->ncell_local, 256), 256, 0, s->stream>>>(
            (const Real *)dens_a, (const Real *)s->background, (Real *)s->rho_src, s->nr, s->rows, s->pitch, (Real)rho_scale);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
}

// `sweeps` in 1..4: ONE launch
int launch_relax(fsim_sim *s, int sweeps, double omega)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        switch (sweeps) {
        case 4: return relax_launch<Real, 4>(s, (Real)omega);
        case 3: return relax_launch<Real, 3>(s, (Real)omega);
        case 2: return relax_launch<Real, 2>(s, (Real)omega);
        case 1: return relax_launch<Real, 1>(s, (Real)omega);
        default: set_error("relax: 1..4 sweeps per launch"); return (int)FSIM_ERR_INVALID;
        }
    });
}

// Periodic z (FSIM_FLAG_PERIODIC_Z): the `nrows` owned rows at either end of a planar field go into the ghost rows
// beyond the other end -- the exchange a slab rank does with its neighbours, here with itself.  Rows are contiguous.
int ring_wrap_rows(fsim_sim *s, void *plane_base, int nrows)
{
    const int h = s->own0 - s->row0;  // ghost rows below the owned block
    if (nrows > h || nrows > s->own_rows) {
        set_error("periodic z: more rows to wrap than ghost rows");
        return FSIM_ERR_RANGE;
    }
    const size_t row = s->rs * (size_t)s->pitch;
    char *b = (char *)plane_base;
    // bottom owned rows -> ghost rows above the top; top owned rows -> ghost rows below the bottom
    FSIM_CUDA(cudaMemcpyAsync(b + row * (size_t)(h + s->own_rows), b + row * (size_t)h, row * (size_t)nrows, cudaMemcpyDeviceToDevice,
                              s->stream));
    FSIM_CUDA(cudaMemcpyAsync(b + row * (size_t)(h - nrows), b + row * (size_t)(h + s->own_rows - nrows), row * (size_t)nrows,
                              cudaMemcpyDeviceToDevice, s->stream));
    return FSIM_OK;
}

int launch_efield(fsim_sim *s)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const double dr = s->spec.radius / (double)s->nr, dz = s->spec.height / (double)s->nz;
        Bracket b(s, "efield");
        efield_kernel<Real><<<grid_for(s->ncell_local, 256), 256, 0, s->stream>>>(
            (const Real *)s->phi[s->phi_cur], (Real *)s->E, s->nr, s->rows, s->pitch, (Real)(1.0 / (2.0 * dr)),
            (Real)(1.0 / (2.0 * dz)));
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
}

}  // namespace fsim
