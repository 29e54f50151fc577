// deposit.cu -- density(): deterministic moment deposition, normalisation, running average.
//
// The reference renders every particle as an 11x11 additive point sprite carrying
// 0.001*(v_r, v_a, v_z, 1) * shape (programMoments01, empic.js:980-1035, launched :1473-1478).
// For a sprite centre (r*nr, z*nz) the fragment at pixel xf samples shape texel
// xf - floor(r*nr) + 5, i.e. the weight depends only on the integer cell offset, so
//     moments01 = (per-cell nearest-grid-point sums of the sprite colours)  (*)  shape[11][11]
// with sprites clipped at the target edge (sources outside the grid do not exist).  That is
// what runs here:
//   cellsum : per cell, the colours of its particles are added sequentially in ASCENDING
//             PARTICLE-ID order (GL primitive order) -> bit-reproducible and equal to a
//             sequential CPU loop over particles, independent of the storage order;
//   conv    : 11x11 stencil on shared-memory tiles fused with programNormalizeMoments01
//             (empic.js:1053-1056), programAvgMoments (avg_frag :274-277, ratio :1083) and the
//             avgA -> avgB copy (:1490-1495).
#include "common.cuh"

namespace fsim {

constexpr int THREAD_CELL_MAX = 64;  // larger cells go to the block-per-cell path

// vertex shader of programMoments01, empic.js:994-1006.  Returns false when the sprite is
// clipped (centre outside the target, or NaN) or does not belong to local cell `c`.
template <typename Real>
__device__ __forceinline__ bool sprite_colour(Real x, Real y, Real z, Real vx, Real vy, Real vz,
                                              int nr, int nz, int row0, uint32_t c, Real (&col)[4])
{
    const Real r = fsqrt(x * x + y * y);
    const Real dx = x / r, dy = y / r;
    const Real vr = vx * dx + vy * dy;
    const Real va = vy * dx - vx * dy;
    col[0] = (Real)FSIM_DEPOSIT_WEIGHT * vr;
    col[1] = (Real)FSIM_DEPOSIT_WEIGHT * va;
    col[2] = (Real)FSIM_DEPOSIT_WEIGHT * vz;
    col[3] = (Real)FSIM_DEPOSIT_WEIGHT * (Real)1.0;
    const Real xw = r * (Real)nr, yw = z * (Real)nz;
    if (!(xw >= (Real)0) || !(xw < (Real)nr)) return false;
    if (!(yw >= (Real)0) || !(yw < (Real)nz)) return false;
    return (uint32_t)((int)xw + ((int)yw - row0) * nr) == c;
}

template <typename Real>
struct CellSumArgs {
    const Real *x, *y, *z, *vx, *vy, *vz;
    const uint32_t *id;
    const uint32_t *starts;
    Real *S;          // [ncell][4]
    uint32_t *count;  // [ncell]
    uint32_t *heavy_list, *heavy_n;
    int64_t ncell;
    int nr, nz, row0;
    // scratch of the block-per-cell path (the idle half of the particle double buffer)
    Real *scol[4];
    uint32_t *sid, *sidx;
};

template <typename Real>
__global__ void __launch_bounds__(128) cellsum_kernel(const CellSumArgs<Real> a)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.ncell) return;
    const uint32_t s = a.starts[c], e = a.starts[c + 1];
    const uint32_t k = e - s;
    Real acc[4] = {(Real)0, (Real)0, (Real)0, (Real)0};
    uint32_t cnt = 0;
    if (k > THREAD_CELL_MAX) {
        a.heavy_list[atomicAdd(a.heavy_n, 1u)] = (uint32_t)c;
        return;
    }
    unsigned long long done = 0ull;
    for (uint32_t t = 0; t < k; ++t) {
        // selection: the not-yet-used particle of this cell with the smallest id
        // (ids are unique and < 0xffffffff)
        uint32_t best = 0xffffffffu, bj = 0;
        for (uint32_t j = 0; j < k; ++j) {
            const uint32_t v = a.id[s + j];
            if (!((done >> j) & 1ull) && v < best) {
                best = v;
                bj = j;
            }
        }
        done |= 1ull << bj;
        const size_t p = (size_t)s + bj;
        Real col[4];
        if (sprite_colour<Real>(a.x[p], a.y[p], a.z[p], a.vx[p], a.vy[p], a.vz[p], a.nr, a.nz, a.row0,
                                (uint32_t)c, col)) {
            acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2]; acc[3] += col[3];
            cnt++;
        }
    }
    Real *o = a.S + 4 * (size_t)c;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = acc[3];
    a.count[c] = cnt;
}

// One block per crowded cell: colours in parallel, ids sorted by an ascending-only bitonic
// network (virtual +inf padding), then one thread adds the colours in id order.
template <typename Real>
__global__ void __launch_bounds__(256) cellsum_heavy_kernel(const CellSumArgs<Real> a)
{
  const uint32_t nheavy = *a.heavy_n;
  for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
    const uint32_t c = a.heavy_list[h];
    const uint32_t s = a.starts[c], e = a.starts[c + 1];
    const uint32_t k = e - s;
    uint32_t *sid = a.sid + s, *sidx = a.sidx + s;
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {
        const size_t p = (size_t)s + j;
        Real col[4];
        const bool ok = sprite_colour<Real>(a.x[p], a.y[p], a.z[p], a.vx[p], a.vy[p], a.vz[p], a.nr,
                                            a.nz, a.row0, c, col);
        a.scol[0][p] = col[0]; a.scol[1][p] = col[1]; a.scol[2][p] = col[2]; a.scol[3][p] = col[3];
        sid[j] = a.id[p];
        sidx[j] = j | (ok ? 0u : 0x80000000u);
    }
    __syncthreads();
    uint32_t np2 = 1;
    while (np2 < k) np2 <<= 1;
    for (uint32_t kk = 2; kk <= np2; kk <<= 1) {
        for (uint32_t jj = kk >> 1; jj > 0; jj >>= 1) {
            for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) {
                const uint32_t l = (jj == (kk >> 1)) ? (i ^ (kk - 1)) : (i ^ jj);
                if (l > i && l < k) {
                    const uint32_t vi = sid[i], vl = sid[l];
                    if (vi > vl) {
                        sid[i] = vl; sid[l] = vi;
                        const uint32_t t = sidx[i]; sidx[i] = sidx[l]; sidx[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        Real acc[4] = {(Real)0, (Real)0, (Real)0, (Real)0};
        uint32_t cnt = 0;
        for (uint32_t t = 0; t < k; ++t) {
            const uint32_t j = sidx[t];
            if (j & 0x80000000u) continue;
            const size_t p = (size_t)s + j;
            acc[0] += a.scol[0][p]; acc[1] += a.scol[1][p];
            acc[2] += a.scol[2][p]; acc[3] += a.scol[3][p];
            cnt++;
        }
        Real *o = a.S + 4 * (size_t)c;
        o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = acc[3];
        a.count[c] = cnt;
    }
    __syncthreads();
  }
}

int launch_cellsum(fsim_sim *s)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int cur = s->cur, alt = s->cur ^ 1;
        CellSumArgs<Real> a;
        a.x = (const Real *)s->part[cur][AX]; a.y = (const Real *)s->part[cur][AY];
        a.z = (const Real *)s->part[cur][AZ]; a.vx = (const Real *)s->part[cur][AVX];
        a.vy = (const Real *)s->part[cur][AVY]; a.vz = (const Real *)s->part[cur][AVZ];
        a.id = s->pid[cur];
        a.starts = s->starts;
        a.S = (Real *)s->cellsum;
        a.count = s->cellcount;
        a.heavy_list = s->heavy_list; a.heavy_n = s->heavy_n;
        a.ncell = s->ncell_local;
        a.nr = s->nr; a.nz = s->nz; a.row0 = s->row0;
        for (int k = 0; k < 4; ++k) a.scol[k] = (Real *)s->part[alt][k];
        a.sid = (uint32_t *)s->part[alt][4];
        a.sidx = (uint32_t *)s->part[alt][5];
        FSIM_CUDA(cudaMemsetAsync(s->heavy_n, 0, sizeof(uint32_t), s->stream));
        {
            Bracket b(s, "cellsum");
            cellsum_kernel<Real><<<grid_for(s->ncell_local, 128), 128, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        {   // crowded cells (count read on the device: no host round trip)
            Bracket b(s, "cellsum_heavy");
            cellsum_heavy_kernel<Real><<<148 * 4, 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        return (int)FSIM_OK;
    });
}

// ---- 11x11 stencil + normalise + running average -------------------------------------------
constexpr int CT_I = 32, CT_J = 16;             // output tile
constexpr int CH = FSIM_SHAPE_MID;              // halo = 5
constexpr int CS_I = CT_I + 2 * CH, CS_J = CT_J + 2 * CH;

__constant__ double c_shape_f64[FSIM_NSHAPE * FSIM_NSHAPE];
__constant__ float c_shape_f32[FSIM_NSHAPE * FSIM_NSHAPE];
template <typename Real> __device__ __forceinline__ Real shape_w(int k);
template <> __device__ __forceinline__ double shape_w<double>(int k) { return c_shape_f64[k]; }
template <> __device__ __forceinline__ float shape_w<float>(int k) { return c_shape_f32[k]; }

template <typename Real>
struct ConvArgs {
    const Real *S;
    Real *mom, *norm, *avg;
    int nr, rows;        // local table: nr x rows
    int j0, j1;          // output rows [j0, j1) (owned rows, local index)
};

template <typename Real>
__global__ void __launch_bounds__(CT_I * 8) conv_kernel(const ConvArgs<Real> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Real *sm = reinterpret_cast<Real *>(smem_raw);  // [CS_J][CS_I][4]
    const int i0 = blockIdx.x * CT_I, jb = a.j0 + blockIdx.y * CT_J;
    const int tid = threadIdx.y * CT_I + threadIdx.x;
    for (int t = tid; t < CS_I * CS_J; t += CT_I * 8) {
        const int li = t % CS_I, lj = t / CS_I;
        const int gi = i0 + li - CH, gj = jb + lj - CH;
        Real v0 = (Real)0, v1 = (Real)0, v2 = (Real)0, v3 = (Real)0;
        if (gi >= 0 && gi < a.nr && gj >= 0 && gj < a.rows) {
            const Real *p = a.S + 4 * ((size_t)gi + (size_t)gj * a.nr);
            v0 = p[0]; v1 = p[1]; v2 = p[2]; v3 = p[3];
        }
        Real *q = sm + 4 * t;
        q[0] = v0; q[1] = v1; q[2] = v2; q[3] = v3;
    }
    __syncthreads();
    // validity of a source (it must exist in the GLOBAL grid rows held locally)
#pragma unroll
    for (int rep = 0; rep < CT_J / 8; ++rep) {
        const int lj = threadIdx.y + rep * 8, li = threadIdx.x;
        const int gi = i0 + li, gj = jb + lj;
        if (gi >= a.nr || gj >= a.j1) continue;
        Real acc0 = (Real)0, acc1 = (Real)0, acc2 = (Real)0, acc3 = (Real)0;
        for (int tj = 0; tj < FSIM_NSHAPE; ++tj) {
            const int sj = gj - tj + CH;  // source row (local)
            if (sj < 0 || sj >= a.rows) continue;
            const Real *row = sm + 4 * ((lj + 2 * CH - tj) * CS_I);
            for (int ti = 0; ti < FSIM_NSHAPE; ++ti) {
                const Real w = shape_w<Real>(ti + FSIM_NSHAPE * tj);
                const int si = gi - ti + CH;
                if (w == (Real)0 || si < 0 || si >= a.nr) continue;
                const Real *q = row + 4 * (li + 2 * CH - ti);
                acc0 = acc0 + q[0] * w; acc1 = acc1 + q[1] * w;
                acc2 = acc2 + q[2] * w; acc3 = acc3 + q[3] * w;
            }
        }
        const size_t c = (size_t)gi + (size_t)gj * a.nr;
        if (a.mom) {
            Real *m = a.mom + 4 * c;
            m[0] = acc0; m[1] = acc1; m[2] = acc2; m[3] = acc3;
        }
        // programNormalizeMoments01, empic.js:1055-1056
        const Real u = ((Real)gi + (Real)0.5) / (Real)a.nr;
        Real M[4];
        if (acc3 > (Real)0) {
            M[0] = acc0 / acc3; M[1] = acc1 / acc3; M[2] = acc2 / acc3; M[3] = acc3;
        } else {
            M[0] = M[1] = M[2] = M[3] = (Real)0;
        }
        const Real ratio = (Real)FSIM_EMA_RATIO;
        Real *av = a.avg + 4 * c;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const Real v = (Real)FSIM_NORM_SCALE * M[q] * (Real)FSIM_NORM_HALF / u;
            if (a.norm) a.norm[4 * c + q] = v;
            av[q] = ratio * v + ((Real)1.0 - ratio) * av[q];  // avg_frag, empic.js:277
        }
    }
}

// shape64: footprint in double; shape32: the Float32Array variant (empic.js:950-971)
int upload_shape(const double *shape64, const double *shape32)
{
    float f[FSIM_NSHAPE * FSIM_NSHAPE];
    for (int k = 0; k < FSIM_NSHAPE * FSIM_NSHAPE; ++k) f[k] = (float)shape32[k];
    FSIM_CUDA(cudaMemcpyToSymbol(c_shape_f64, shape64, sizeof(double) * FSIM_NSHAPE * FSIM_NSHAPE));
    FSIM_CUDA(cudaMemcpyToSymbol(c_shape_f32, f, sizeof(f)));
    return FSIM_OK;
}

int launch_conv(fsim_sim *s)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        ConvArgs<Real> a;
        a.S = (const Real *)s->cellsum;
        a.mom = (Real *)s->mom; a.norm = (Real *)s->norm; a.avg = (Real *)s->avg;
        a.nr = s->nr; a.rows = s->rows;
        a.j0 = s->own0 - s->row0; a.j1 = a.j0 + s->own_rows;
        dim3 block(CT_I, 8);
        dim3 grid((s->nr + CT_I - 1) / CT_I, (s->own_rows + CT_J - 1) / CT_J);
        const size_t smem = sizeof(Real) * 4 * CS_I * CS_J;
        Bracket b(s, "conv");
        conv_kernel<Real><<<grid, block, smem, s->stream>>>(a);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
}

}  // namespace fsim
