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
#include <cuda.h>

#include "common.cuh"
#include "sortnet.cuh"

namespace fsim {

#ifdef FSIM_TUNE
int g_conv_variant = 0;
#endif

constexpr int THREAD_CELL_MAX = 16;   // up to here: one thread per cell, (id, slot) pairs sorted in registers
constexpr int WARP_CELL_MAX = 256;    // up to here: one warp per cell; larger cells: one block per cell

template <typename Real>
struct CellSumArgs {
    const Real *dcol[2];   // sprite colour 0.001 (v_r, v_a) of each slot (deposit prepass)
    const Real *vz;        // stored v_z of each slot: the third colour is 0.001 * v_z (empic.js:1006)
    const uint32_t *key;   // KEY_CLIPPED marks sprites that are not deposited
    const uint32_t *id;
    const uint32_t *perm;  // particle slots in cell order
    const uint32_t *starts;
    Real *S;          // planar: channel q of cell (i,j) at S[q*plane + j*pitch + i]
    uint32_t *count;  // [ncell]
    int pitch;
    int64_t plane;
    uint32_t *heavy_list, *heavy_n;   // heavy_n[0]: cells for the block path, heavy_n[1]: for the warp path
    uint32_t *medium_list;
    int64_t ncell;
    int64_t n;        // live slots (upper bound), for the debug-build index checks
    int nr, nz, row0;
    // scratch of the block-per-cell path (the idle half of the particle double buffer)
    uint32_t *sid, *sidx;
    Real *scol[3];    // colours of a crowded cell in id order
};

// Register path for a cell with k <= KR particles: (id, slot) pairs are loaded once, ordered by id
// with a compile-time sorting network (padding = 0xffffffff sorts last), then the colours are
// added in that order.  All loads of a wave are independent, so they overlap.
template <typename Real, int KR>
__device__ __forceinline__ void cell_small(const CellSumArgs<Real> &a, uint32_t s, uint32_t k, Real (&acc)[4],
                                           uint32_t &cnt)
{
    uint32_t pp[KR], ii[KR];  // pp: slot | clipped flag (bit 31), as index_scatter stored it
#pragma unroll
    for (int j = 0; j < KR; ++j) pp[j] = (uint32_t)j < k ? a.perm[s + j] : KEY_CLIPPED;
#pragma unroll
    for (int j = 0; j < KR; ++j) FSIM_ASSERT((uint32_t)j >= k || ((int64_t)s + j < a.n && (int64_t)(pp[j] & KEY_MASK) < a.n));
#pragma unroll
    for (int j = 0; j < KR; ++j) ii[j] = (uint32_t)j < k ? a.id[pp[j] & KEY_MASK] : 0xffffffffu;
#pragma unroll
    for (int c = 0; c < SortNet<KR>::COUNT; ++c) {  // Batcher odd-even merge network (sortnet.cuh), ascending
        const int i = SortNet<KR>::A(c), l = SortNet<KR>::B(c);
        const bool sw = ii[i] > ii[l];
        const uint32_t ti = sw ? ii[l] : ii[i], tl = sw ? ii[i] : ii[l];
        const uint32_t qi = sw ? pp[l] : pp[i], ql = sw ? pp[i] : pp[l];
        ii[i] = ti; ii[l] = tl; pp[i] = qi; pp[l] = ql;
    }
    Real c0[KR], c1[KR], c2[KR];
#pragma unroll
    for (int j = 0; j < KR; ++j) {
        const bool on = !(pp[j] & KEY_CLIPPED);
        const size_t p = pp[j] & KEY_MASK;
        c0[j] = on ? a.dcol[0][p] : (Real)0;
        c1[j] = on ? a.dcol[1][p] : (Real)0;
        c2[j] = on ? (Real)FSIM_DEPOSIT_WEIGHT * a.vz[p] : (Real)0;
    }
#pragma unroll
    for (int j = 0; j < KR; ++j) {
        if (!(pp[j] & KEY_CLIPPED)) {
            acc[0] += c0[j]; acc[1] += c1[j]; acc[2] += c2[j];
            acc[3] += (Real)FSIM_DEPOSIT_WEIGHT * (Real)1.0;
            cnt++;
        }
    }
}

template <typename Real>
__global__ void __launch_bounds__(128) cellsum_kernel(const CellSumArgs<Real> a)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c < a.ncell;
    uint32_t s = 0, k = 0;
    if (live) {
        s = a.starts[c];
        k = a.starts[c + 1] - s;
    }
    Real acc[4] = {(Real)0, (Real)0, (Real)0, (Real)0};
    uint32_t cnt = 0;
    // warp-uniform choice of the network: the largest cell of the warp that stays on this path
    // (a warp of 32 cells with a mean of 4 particles needs 8 wires half of the time, 10 or 12 otherwise)
    const uint32_t ks = k <= THREAD_CELL_MAX ? k : 0;  // crowded cells go to the warp / block kernels
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, ks);
    if (kmax <= 4) cell_small<Real, 4>(a, s, ks, acc, cnt);
    else if (kmax <= 6) cell_small<Real, 6>(a, s, ks, acc, cnt);
    else if (kmax <= 8) cell_small<Real, 8>(a, s, ks, acc, cnt);
    else if (kmax <= 10) cell_small<Real, 10>(a, s, ks, acc, cnt);
    else if (kmax <= 12) cell_small<Real, 12>(a, s, ks, acc, cnt);
    else cell_small<Real, 16>(a, s, ks, acc, cnt);
    if (k <= THREAD_CELL_MAX) {
        // summed above
    } else if (k > WARP_CELL_MAX) {
        a.heavy_list[atomicAdd(a.heavy_n, 1u)] = (uint32_t)c;
        return;
    } else {
        a.medium_list[atomicAdd(a.heavy_n + 1, 1u)] = (uint32_t)c;
        return;
    }
    if (!live) return;
    Real *o = a.S + (size_t)(c / a.nr) * a.pitch + (size_t)(c % a.nr);
    o[0] = acc[0]; o[a.plane] = acc[1]; o[2 * a.plane] = acc[2]; o[3 * a.plane] = acc[3];
    a.count[c] = cnt;
}

// One warp per cell of 17..256 particles (the reference's own demo scene: 12-35 per occupied cell,
// up to 100).  Lanes hold the (id, slot) pairs; an element's position in id order is the number of
// smaller ids, counted with one shuffle per element; the colours are fetched in parallel and parked
// in shared memory in id order; lanes 0..3 then add one channel each, sequentially.
// NQ = 32-particle groups a lane holds: the counting loop costs k (1 + NQ) instructions per lane, so a cell of
// 17..32 particles (the common case of the demo scene) runs the NQ = 1 instance -- 64 instead of 288.
template <typename Real, int NQ>
__device__ __forceinline__ void cell_medium(const CellSumArgs<Real> &a, const uint32_t c, const uint32_t s, const uint32_t k,
                                            const int lane, Real (*s_col)[WARP_CELL_MAX], uint8_t *s_on)
{
    uint32_t pp[NQ], ii[NQ], rank[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const uint32_t j = q * 32 + lane;
        pp[q] = (q * 32 < k && j < k) ? a.perm[(size_t)s + j] : KEY_CLIPPED;
        FSIM_ASSERT(!(q * 32 < k && j < k) || ((int64_t)s + j < a.n && (int64_t)(pp[q] & KEY_MASK) < a.n));
        ii[q] = (q * 32 < k && j < k) ? a.id[pp[q] & KEY_MASK] : 0xffffffffu;
        rank[q] = 0;
    }
#pragma unroll
    for (int qt = 0; qt < NQ; ++qt) {
        if (qt * 32 >= k) break;  // warp-uniform
        const int nt = min(32u, k - qt * 32);
        for (int lt = 0; lt < nt; ++lt) {
            const uint32_t v = __shfl_sync(0xffffffffu, ii[qt], lt);
#pragma unroll
            for (int q = 0; q < NQ; ++q) rank[q] += (v < ii[q]) ? 1u : 0u;
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        if (q * 32 >= k) break;
        if (q * 32 + lane < k) {
            const bool on = !(pp[q] & KEY_CLIPPED);
            const size_t p = pp[q] & KEY_MASK;
            const uint32_t r = rank[q];
            s_on[r] = on ? 1 : 0;
            s_col[0][r] = on ? a.dcol[0][p] : (Real)0;
            s_col[1][r] = on ? a.dcol[1][p] : (Real)0;
            s_col[2][r] = on ? (Real)FSIM_DEPOSIT_WEIGHT * a.vz[p] : (Real)0;
        }
    }
    __syncwarp();
    if (lane < 4) {
        Real acc = (Real)0;
        uint32_t cnt = 0;
        for (uint32_t t = 0; t < k; ++t) {
            if (!s_on[t]) continue;
            acc += (lane < 3) ? s_col[lane < 3 ? lane : 0][t] : (Real)FSIM_DEPOSIT_WEIGHT * (Real)1.0;
            cnt++;
        }
        a.S[(size_t)lane * a.plane + (size_t)(c / a.nr) * a.pitch + (size_t)(c % a.nr)] = acc;
        if (lane == 3) a.count[c] = cnt;
    }
    __syncwarp();
}

template <typename Real>
__global__ void __launch_bounds__(128) cellsum_warp_kernel(const CellSumArgs<Real> a)
{
    __shared__ Real s_col[4][3][WARP_CELL_MAX];
    __shared__ uint8_t s_on[4][WARP_CELL_MAX];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t nmed = a.heavy_n[1];
    for (uint32_t h = blockIdx.x * 4 + w; h < nmed; h += gridDim.x * 4) {
        const uint32_t c = a.medium_list[h];
        const uint32_t s = a.starts[c], k = a.starts[c + 1] - s;
        if (k <= 32) cell_medium<Real, 1>(a, c, s, k, lane, s_col[w], s_on[w]);
        else if (k <= 64) cell_medium<Real, 2>(a, c, s, k, lane, s_col[w], s_on[w]);
        else if (k <= 128) cell_medium<Real, 4>(a, c, s, k, lane, s_col[w], s_on[w]);
        else cell_medium<Real, WARP_CELL_MAX / 32>(a, c, s, k, lane, s_col[w], s_on[w]);
    }
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
        const uint32_t pf = a.perm[(size_t)s + j];
        FSIM_ASSERT((int64_t)s + j < a.n && (int64_t)(pf & KEY_MASK) < a.n);
        sid[j] = a.id[pf & KEY_MASK];
        sidx[j] = j | (pf & KEY_CLIPPED);
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
    // the colours, gathered by all threads and parked in id order (a clipped sprite parks exact zeros: x + 0 = x
    // for every x a sum that started at +0 can hold, so the sequential additions below need no branch) ...
    Real *sc0 = a.scol[0] + s, *sc1 = a.scol[1] + s, *sc2 = a.scol[2] + s;
    for (uint32_t t = threadIdx.x; t < k; t += blockDim.x) {
        const uint32_t j = sidx[t];
        const bool on = !(j & KEY_CLIPPED);
        const size_t p = a.perm[(size_t)s + (j & KEY_MASK)] & KEY_MASK;
        sc0[t] = on ? a.dcol[0][p] : (Real)0;
        sc1[t] = on ? a.dcol[1][p] : (Real)0;
        sc2[t] = on ? (Real)FSIM_DEPOSIT_WEIGHT * a.vz[p] : (Real)0;
    }
    __syncthreads();
    // ... then one thread per channel adds them in that order: contiguous reads, nothing but the additions in the chain
    if (threadIdx.x < 4) {
        const int q = threadIdx.x;
        const Real *src = q == 0 ? sc0 : (q == 1 ? sc1 : sc2);
        Real acc = (Real)0;
        uint32_t cnt = 0;
        for (uint32_t t = 0; t < k; ++t) {
            const bool on = !(sidx[t] & KEY_CLIPPED);
            acc += (q < 3) ? src[t] : (on ? (Real)FSIM_DEPOSIT_WEIGHT * (Real)1.0 : (Real)0);
            cnt += on ? 1u : 0u;
        }
        a.S[(size_t)q * a.plane + (size_t)(c / a.nr) * a.pitch + (size_t)(c % a.nr)] = acc;
        if (q == 3) a.count[c] = cnt;
    }
    __syncthreads();
  }
}

// Measured alternative (FSIM_FLAG_ATOMIC_DEPOSIT): particle-parallel global atomics on the per-cell
// sums.  No sort needed, but the floating-point sum order is the arrival order of the atomics, so
// the result is NOT bit-reproducible; the id-ordered path above is the default.
template <typename Real>
__global__ void __launch_bounds__(256)
cellsum_atomic_kernel(const uint32_t *__restrict__ key, const Real *__restrict__ c0, const Real *__restrict__ c1,
                      const Real *__restrict__ vz, int64_t n, Real *__restrict__ S, uint32_t *__restrict__ count,
                      int nr, int pitch, int64_t plane)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t k = key[p];
    if (k & KEY_CLIPPED) return;
    Real *o = S + (size_t)(k / nr) * pitch + (size_t)(k % nr);
    atomicAdd(o, c0[p]);
    atomicAdd(o + plane, c1[p]);
    atomicAdd(o + 2 * plane, (Real)FSIM_DEPOSIT_WEIGHT * vz[p]);
    atomicAdd(o + 3 * plane, (Real)FSIM_DEPOSIT_WEIGHT * (Real)1.0);
    atomicAdd(count + k, 1u);
}

int launch_cellsum_atomic(fsim_sim *s)
{
    FSIM_TRY(post_join(s));  // the previous frame's stencil reads the sums
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        FSIM_CUDA(cudaMemsetAsync(s->cellsum, 0, sizeof(Real) * 4 * s->plane, s->stream));
        FSIM_CUDA(cudaMemsetAsync(s->cellcount, 0, sizeof(uint32_t) * s->ncell_local, s->stream));
        if (s->n == 0) return (int)FSIM_OK;
        Bracket b(s, "cellsum_atomic");
        cellsum_atomic_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
            s->key, (const Real *)s->dcol[0], (const Real *)s->dcol[1], (const Real *)s->part[s->cur][AVZ], s->n,
            (Real *)s->cellsum, s->cellcount, s->nr, s->pitch, s->plane);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
}

int launch_cellsum(fsim_sim *s)
{
    FSIM_TRY(post_join(s));  // the previous frame's stencil (second stream) reads the sums this pass overwrites
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int cur = s->cur, alt = s->cur ^ 1;
        CellSumArgs<Real> a;
        for (int q = 0; q < 2; ++q) a.dcol[q] = (const Real *)s->dcol[q];
        a.vz = (const Real *)s->part[cur][AVZ];
        a.key = s->key;
        a.id = s->pid[cur];
        a.perm = s->perm;
        a.starts = s->starts;
        a.S = (Real *)s->cellsum;
        a.count = s->cellcount;
        a.pitch = s->pitch; a.plane = s->plane;
        a.heavy_list = s->heavy_list; a.heavy_n = s->heavy_n;
        a.medium_list = s->medium_list;
        a.ncell = s->ncell_local;
        a.n = s->n;
        a.nr = s->nr; a.nz = s->nz; a.row0 = s->row0;
        a.sid = (uint32_t *)s->part[alt][4];
        a.sidx = (uint32_t *)s->part[alt][5];
        for (int q = 0; q < 3; ++q) a.scol[q] = (Real *)s->part[alt][6 + q];
        FSIM_CUDA(cudaMemsetAsync(s->heavy_n, 0, 2 * sizeof(uint32_t), s->stream));
        {
            Bracket b(s, "cellsum");
            cellsum_kernel<Real><<<grid_for(s->ncell_local, 128), 128, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        // crowded cells (list lengths are read on the device: no host round trip)
        {
            Bracket b(s, "cellsum_warp");
            cellsum_warp_kernel<Real><<<s->nsm * 8, 128, 0, s->stream>>>(a);  // 8 blocks of 25.6 KB shared memory fill an SM
            FSIM_CUDA(cudaGetLastError());
        }
        {
            Bracket b(s, "cellsum_heavy");
            cellsum_heavy_kernel<Real><<<s->nsm * 4, 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        return (int)FSIM_OK;
    });
}

// ---- 11x11 stencil + normalise + running average -------------------------------------------
// Tile: CT_I x CT_J output cells per block.  The per-cell sums are stored planar (one plane per
// channel, row pitch padded to 16 bytes), so ONE 3-D TMA box load (cp.async.bulk.tensor, mbarrier
// completion) brings the tile and its 5-cell halo of all four channels into shared memory;
// coordinates outside the tensor -- the grid edge -- are zero-filled by the TMA unit, which is
// exactly "sprites are clipped at the target edge": adding exact zeros equals skipping the source.
// A warp owns one channel of a 32-column x CSTRIP-row strip; a lane owns one column: per column
// offset di it reads two vertical windows of CSTRIP + 10 values (columns -di and +di; consecutive
// lanes read consecutive words, so no bank conflicts for any plane stride) and slides them over
// its CSTRIP outputs.  The footprint is mirror-symmetric, so the up to four mirror sources of a
// weight are added first and weighted once: classes di = 0..5 outer, dj = 0..5 inner, the two
// sources of a row first, then the two rows, (S[-di,-dj] + S[+di,-dj]) + (S[-di,+dj] + S[+di,+dj])
// -- the canonical order of the specification.  The row pairs do not depend on the output row: they
// are formed once per window (18 adds per column offset for 8 outputs) and shared, ~82 fp64
// operations per cell and channel instead of 107 (162 without the symmetry); the 40 taps that are
// exactly zero are removed at compile time.  With the two IEEE divisions of the normalisation the
// kernel is bound by the fp64 pipe.  Measured on B200 at C5 fp64 (tuning build, tools/tune.py): 8-row strips on
// 32 x 8 tiles 0.65 ms, 8-row strips on 32 x 16 tiles 0.69, 4-row strips on 32 x 16 tiles 0.86,
// 16-row strips 0.88-0.90, 2-row strips 1.50 -- longer strips need fewer shared-memory loads per
// output, small tiles put more blocks on an SM to hide the single TMA wait of each.
constexpr int CT_I = 32;                        // output tile width (one lane per column)
constexpr int CH = FSIM_SHAPE_MID;              // halo = 5
// TMA needs the box START (innermost coordinate x element size) 16-byte aligned, and the box width
// a multiple of 16 bytes: the box therefore begins CXOFF >= 5 cells left of the tile, CXOFF a
// multiple of 2 reals (fp64) / 4 reals (fp32), and is CBOXW >= CXOFF + 32 + 5 wide.  (A start at
// -5 cells is an "illegal instruction" fault: measured on B200.)
template <typename Real> struct ConvBox;
template <> struct ConvBox<double> { static constexpr int XOFF = 6, W = 44; };
template <> struct ConvBox<float> { static constexpr int XOFF = 8, W = 48; };

__constant__ double c_shape_f64[FSIM_NSHAPE * FSIM_NSHAPE];
__constant__ float c_shape_f32[FSIM_NSHAPE * FSIM_NSHAPE];
template <typename Real> __device__ __forceinline__ Real shape_w(int k);
template <> __device__ __forceinline__ double shape_w<double>(int k) { return c_shape_f64[k]; }
template <> __device__ __forceinline__ float shape_w<float>(int k) { return c_shape_f32[k]; }

// taps of the footprint that are not identically zero: cos^2(pi d / 10) with d <= 5 (empic.js:959)
__host__ __device__ constexpr bool tap_nonzero(int ti, int tj)
{
    return (ti - CH) * (ti - CH) + (tj - CH) * (tj - CH) <= CH * CH;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename Real>
struct ConvArgs {
    Real *mom, *norm, *avg;   // planar like S
    int nr, pitch;
    int64_t plane;
    int j0, j1;               // output rows [j0, j1) (owned rows, local index)
    int tile0;                // first tile row of this launch (blockIdx.y + tile0)
    int skip0, skip1;         // tile rows [skip0, skip1) are left out (boundary-only launch); empty when skip0 >= skip1
};

// CSTRIP = output rows per thread (window of CSTRIP + 10 values per column offset), CT_J = tile height
template <typename Real, int CSTRIP, int CT_J>
__global__ void __launch_bounds__(CT_J / CSTRIP * 4 * 32)
conv_kernel(const __grid_constant__ CUtensorMap tmS, const ConvArgs<Real> a)
{
    constexpr int CBOXW = ConvBox<Real>::W, CXOFF = ConvBox<Real>::XOFF;
    constexpr int CS_J = CT_J + 2 * CH, CWIN = CSTRIP + 2 * CH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Real *sm = reinterpret_cast<Real *>(smem_raw);  // [4][CS_J][CBOXW], written by the TMA unit
    __shared__ __align__(8) unsigned long long bar;
    int ty = (int)blockIdx.y + a.tile0;
    if (ty >= a.skip0) ty += a.skip1 - a.skip0;  // boundary launch: jump over the interior tile rows
    const int i0 = blockIdx.x * CT_I, jb = a.j0 + ty * CT_J;
    const int tid = threadIdx.x;
    constexpr uint32_t kBytes = 4u * CS_J * CBOXW * sizeof(Real);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(kBytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_u32(sm)), "l"(reinterpret_cast<unsigned long long>(&tmS)), "r"(i0 - CXOFF), "r"(jb - CH), "r"(0),
            "r"(smem_u32(&bar))
            : "memory");
    }
    {
        uint32_t done = 0;
        for (uint32_t spin = 0; !done; ++spin) {
            if (spin > (1u << 24)) __trap();  // a TMA that never completes must not hang the GPU
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0u)
                : "memory");
        }
    }

    const int lane = tid & 31, w = tid >> 5;
    const int c = w & 3, strip = w >> 2;  // channel; strip of CSTRIP output rows
    const Real *col = sm + c * (CS_J * CBOXW) + (strip * CSTRIP) * CBOXW + lane + CXOFF;
    Real acc[CSTRIP];
#pragma unroll
    for (int o = 0; o < CSTRIP; ++o) acc[o] = (Real)0;
    // A tile nothing was deposited near (the weight channel of the whole box is zero: every deposited
    // sprite adds 0.001 to it) convolves to exact zeros in all channels: skip the stencil.  The
    // reference's demo scene fills 4 % of its cells; a uniform plasma never takes this branch.
    bool any = false;
    for (int k = tid; k < CS_J * CBOXW; k += CT_J / CSTRIP * 4 * 32) any |= sm[3 * (CS_J * CBOXW) + k] != (Real)0;
    const bool occupied = __syncthreads_or(any);
    if (occupied) {
#pragma unroll
    for (int di = 0; di <= CH; ++di) {
        // the row pair S[-di] + S[+di] of every window row, formed ONCE and shared by the CSTRIP outputs
        // that slide over the window (the canonical order adds the two sources of a row first)
        Real hw[CWIN];
#pragma unroll
        for (int k = 0; k < CWIN; ++k) {
            hw[k] = col[k * CBOXW - di];
            if (di) hw[k] = hw[k] + col[k * CBOXW + di];
        }
#pragma unroll
        for (int o = 0; o < CSTRIP; ++o) {
#pragma unroll
            for (int dj = 0; dj <= CH; ++dj) {
                if (!tap_nonzero(CH + di, CH + dj)) continue;
                Real sum = hw[o + CH - dj];
                if (dj) sum = sum + hw[o + CH + dj];
                acc[o] = acc[o] + sum * shape_w<Real>((CH + di) + FSIM_NSHAPE * (CH + dj));
            }
        }
    }

    }  // occupied

    // the four channels of a cell sit in four warps: exchange through shared memory (the tile is dead)
    __syncthreads();
    Real *ex = sm;  // [4][CT_J][CT_I]
#pragma unroll
    for (int o = 0; o < CSTRIP; ++o) ex[(c * CT_J + strip * CSTRIP + o) * CT_I + lane] = acc[o];
    __syncthreads();
    const Real ratio = (Real)FSIM_EMA_RATIO;
    const int gi = i0 + lane;
#pragma unroll
    for (int o = 0; o < CSTRIP; ++o) {
        const int row = strip * CSTRIP + o, gj = jb + row;
        if (gi >= a.nr || gj >= a.j1) continue;
        const Real alpha = ex[(3 * CT_J + row) * CT_I + lane];
        const size_t idx = (size_t)c * a.plane + (size_t)gj * a.pitch + gi;
        if (a.mom) a.mom[idx] = acc[o];
        // programNormalizeMoments01, empic.js:1055-1056
        const Real u = ((Real)gi + (Real)0.5) / (Real)a.nr;
        Real M = (Real)0;
        if (alpha > (Real)0) M = (c == 3) ? alpha : acc[o] / alpha;
        const Real v = (Real)FSIM_NORM_SCALE * M * (Real)FSIM_NORM_HALF / u;
        if (a.norm) a.norm[idx] = v;
        a.avg[idx] = ratio * v + ((Real)1.0 - ratio) * a.avg[idx];  // avg_frag, empic.js:277
    }
}

// Tensor map of the planar per-cell sums: dims (i, j, channel), box (CBOXW, CS_J, 4), zero fill.
// cuTensorMapEncodeTiled is taken from the driver through the runtime (no link against libcuda).
int make_sums_tensor_map(fsim_sim *s, int box_rows)
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
    const cuuint64_t gdim[3] = {(cuuint64_t)s->nr, (cuuint64_t)s->rows, 4};
    const cuuint64_t gstride[2] = {(cuuint64_t)s->pitch * s->rs, (cuuint64_t)s->plane * s->rs};
    const cuuint32_t box[3] = {(cuuint32_t)(s->prec == FSIM_F64 ? ConvBox<double>::W : ConvBox<float>::W), (cuuint32_t)box_rows, 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    static_assert(sizeof(CUtensorMap) == sizeof(s->tm_sums), "tensor map storage");
    const CUresult r = ((EncodeFn)fn)(reinterpret_cast<CUtensorMap *>(s->tm_sums),
                                      s->prec == FSIM_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                      3, s->cellsum, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return FSIM_ERR_CUDA;
    }
    s->tm_sums_rows = box_rows;
    return FSIM_OK;
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

// part 0: every tile row; 1: the tile rows whose 5-row halo lies inside the owned rows (no received halo row
// is read); 2: the others (first tile row, last one or two).  Parts 1 and 2 are disjoint and cover part 0:
// the running average is updated in place, so every cell must be visited exactly once per density().
template <typename Real, int CSTRIP, int CT_J>
static int conv_launch(fsim_sim *s, ConvArgs<Real> a, int part, cudaStream_t st)
{
    constexpr int CS_J = CT_J + 2 * CH;
    if (s->tm_sums_rows != CS_J) FSIM_TRY(make_sums_tensor_map(s, CS_J));  // the box height is part of the map
    const size_t smem = sizeof(Real) * 4 * CS_J * ConvBox<Real>::W;
    // one bit per (strip, tile height) variant: strips 2, 4, 8, 16 x heights 8, 16, 32 -> bits 8..19
    constexpr uint32_t bit = 1u << (8 + (CSTRIP == 2 ? 0 : CSTRIP == 4 ? 1 : CSTRIP == 8 ? 2 : 3) * 3 + (CT_J == 8 ? 0 : CT_J == 16 ? 1 : 2));
    if (!(s->smem_opt_in & bit)) {  // per handle, hence per device: the attribute belongs to the device's context
        FSIM_CUDA(cudaFuncSetAttribute(conv_kernel<Real, CSTRIP, CT_J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s->smem_opt_in |= bit;
    }
    dim3 block(CT_J / CSTRIP * 4 * 32);
    const int ntile = (s->own_rows + CT_J - 1) / CT_J;
    // tile row t reads sums rows [t*CT_J - 5, t*CT_J + CT_J + 5) of the owned block: interior if that stays inside it
    const int in0 = std::min(ntile, (CH + CT_J - 1) / CT_J);
    const int in1 = std::max(in0, (s->own_rows - CH) / CT_J);  // tiles t < in1 end at (t+1)*CT_J <= own_rows - 5
    a.tile0 = 0; a.skip0 = a.skip1 = 1 << 30;
    int rows_of_tiles = ntile;
    if (part == 1) { a.tile0 = in0; rows_of_tiles = in1 - in0; }
    if (part == 2) { a.skip0 = in0; a.skip1 = in1; rows_of_tiles = ntile - (in1 - in0); }
    if (rows_of_tiles <= 0) return FSIM_OK;
    dim3 grid((s->nr + CT_I - 1) / CT_I, rows_of_tiles);
    conv_kernel<Real, CSTRIP, CT_J><<<grid, block, smem, st>>>(*reinterpret_cast<const CUtensorMap *>(s->tm_sums), a);
    FSIM_CUDA(cudaGetLastError());
    return FSIM_OK;
}

int launch_conv_rows(fsim_sim *s, int part)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        ConvArgs<Real> a;
        a.mom = (Real *)s->mom; a.norm = (Real *)s->norm; a.avg = (Real *)s->avg;
        a.nr = s->nr; a.pitch = s->pitch; a.plane = s->plane;
        a.j0 = s->own0 - s->row0; a.j1 = a.j0 + s->own_rows;
        // on the second stream: overlaps the next frame's sweep (DRAM-bound) with this fp64-bound stencil
        const cudaStream_t st = post_begin(s);
        struct End { fsim_sim *s; ~End() { post_end(s); } } end{s};
        Bracket b(s, "conv", st);
#ifdef FSIM_TUNE
        switch (g_conv_variant) {  // tuning build only (tools/tune.py)
        case 1: return conv_launch<Real, 8, 16>(s, a, part, st);
        case 2: return conv_launch<Real, 8, 32>(s, a, part, st);
        case 3: return conv_launch<Real, 4, 32>(s, a, part, st);
        case 4: return conv_launch<Real, 2, 16>(s, a, part, st);
        case 5: return conv_launch<Real, 16, 16>(s, a, part, st);
        case 6: return conv_launch<Real, 16, 32>(s, a, part, st);
        case 8: return conv_launch<Real, 4, 16>(s, a, part, st);
        default: break;
        }
#endif
        return conv_launch<Real, 8, 8>(s, a, part, st);  // 0.65 ms at C5 fp64 against 0.86 for <4,16>, 0.69 <8,16>, 0.88 <16,32>
    });
}

}  // namespace fsim
