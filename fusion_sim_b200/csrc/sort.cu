// sort.cu -- out-of-place counting sort of the particle storage by gather cell (sm_100a).
//
// Why: the reference gathers R1..R3,A at the particle's cell through four uncoalesced texture
// fetches per fragment (step_velocity_frag, empic.js:763-766) and deposits by 121 blended
// fragments per particle (programMoments01, empic.js:1473-1478).  Keeping the storage ordered
// by cell turns the gather into a streaming read of the cell table and gives every cell a
// contiguous particle segment for the deterministic per-cell sums (deposit.cu).
//
// Three launches: (1) key + histogram, (2) exclusive scan of the per-cell counts,
// (3) scatter of the 10 real arrays, the alive byte and the particle id.  The order INSIDE a
// cell segment is not fixed by this sort (atomic cursors); determinism of the deposit comes
// from summing each segment in ascending particle-id order (deposit.cu), and the accessors
// un-permute by id, so nothing observable depends on it.
#include "common.cuh"

namespace fsim {

template <typename Real>
__global__ void __launch_bounds__(256)
key_hist_kernel(const Real *__restrict__ x, const Real *__restrict__ y, const Real *__restrict__ z,
                int64_t n, int nr, int nz, int row0, int rows, uint32_t *__restrict__ key,
                uint32_t *__restrict__ counts)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < n;
    uint32_t c = 0xffffffffu;  // lanes past the end form their own (ignored) group
    if (valid) {
        const Real xx = x[p], yy = y[p];
        const Real r = fsqrt(xx * xx + yy * yy);
        int cj = tex_idx(z[p], nz) - row0;
        cj = cj < 0 ? 0 : (cj >= rows ? rows - 1 : cj);
        c = (uint32_t)tex_idx(r, nr) + (uint32_t)cj * (uint32_t)nr;
        key[p] = c;
    }
    // warp-aggregated histogram: sorted input puts a handful of distinct cells in a warp
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (valid && (__ffs(peers) - 1) == (int)(threadIdx.x & 31))
        atomicAdd(counts + c, (uint32_t)__popc(peers));
}

// ---- exclusive scan over ncell counts: per-block sums, scan of block sums, final pass ----
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t wsum[SCAN_BLOCK / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < SCAN_BLOCK / 32 ? wsum[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_BLOCK / 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane < SCAN_BLOCK / 32) wsum[lane] = s;
    }
    __syncthreads();
    const uint32_t base = w ? wsum[w - 1] : 0;
    if (total) *total = wsum[SCAN_BLOCK / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_tile_sums_kernel(const uint32_t *__restrict__ counts, int64_t m, uint32_t *__restrict__ tile_sums)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < m) s += counts[base + k];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_tile_offsets_kernel(uint32_t *tile_sums, int ntiles)
{
    // single block: sequential over chunks of SCAN_BLOCK tiles
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int t0 = 0; t0 < ntiles; t0 += SCAN_BLOCK) {
        const int t = t0 + threadIdx.x;
        const uint32_t v = t < ntiles ? tile_sums[t] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total);
        const uint32_t c = carry;
        if (t < ntiles) tile_sums[t] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}

// writes starts[c] (exclusive prefix), cursor[c] = starts[c], starts[m] = total, and zeroes
// counts[] for the next histogram.
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_final_kernel(uint32_t *__restrict__ counts, int64_t m, const uint32_t *__restrict__ tile_offsets,
                  uint32_t *__restrict__ starts, uint32_t *__restrict__ cursor)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < m) ? counts[base + k] : 0;
        s += v[k];
    }
    uint32_t run = tile_offsets[blockIdx.x] + block_exclusive_scan(s, nullptr);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < m) {
            starts[base + k] = run;
            cursor[base + k] = run;
            counts[base + k] = 0;
        }
        run += v[k];
        if (base + k == m - 1) starts[m] = run;
    }
}

template <typename Real>
struct ScatterArgs {
    const Real *src[NPART_ARRAYS];
    Real *dst[NPART_ARRAYS];
    const uint8_t *alive_src;
    uint8_t *alive_dst;
    const uint32_t *id_src;
    uint32_t *id_dst;
    const uint32_t *key;
    uint32_t *cursor;
    int64_t n;
};

template <typename Real>
__global__ void __launch_bounds__(256) scatter_kernel(const ScatterArgs<Real> a)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < a.n;
    const uint32_t c = valid ? a.key[p] : 0xffffffffu;
    // one atomic per distinct cell per warp; lanes of the same cell take consecutive slots
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(a.cursor + c, (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!valid) return;
    const size_t d = (size_t)base + (size_t)__popc(peers & ((1u << lane) - 1u));
#pragma unroll
    for (int k = 0; k < NPART_ARRAYS; ++k) a.dst[k][d] = __ldcs(a.src[k] + p);
    a.alive_dst[d] = a.alive_src[p];
    a.id_dst[d] = a.id_src[p];
}

int launch_sort(fsim_sim *s)
{
    if (s->n == 0) {
        FSIM_CUDA(cudaMemsetAsync(s->starts, 0, sizeof(uint32_t) * (s->ncell_local + 2), s->stream));
        s->sorted = true;
        return FSIM_OK;
    }
    const int64_t m = s->ncell_local;
    const int ntiles = (int)((m + SCAN_TILE - 1) / SCAN_TILE);
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int src = s->cur, dst = s->cur ^ 1;
        {
            Bracket b(s, "hist");
            key_hist_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
                (const Real *)s->part[src][AX], (const Real *)s->part[src][AY],
                (const Real *)s->part[src][AZ], s->n, s->nr, s->nz, s->row0, s->rows, s->key, s->counts);
            FSIM_CUDA(cudaGetLastError());
        }
        {
            Bracket b(s, "scan");
            scan_tile_sums_kernel<<<ntiles, SCAN_BLOCK, 0, s->stream>>>(s->counts, m, s->blocksums);
            scan_tile_offsets_kernel<<<1, SCAN_BLOCK, 0, s->stream>>>(s->blocksums, ntiles);
            scan_final_kernel<<<ntiles, SCAN_BLOCK, 0, s->stream>>>(s->counts, m, s->blocksums, s->starts,
                                                                   s->cursor);
            s->launches += 2;
            FSIM_CUDA(cudaGetLastError());
        }
        {
            ScatterArgs<Real> a;
            for (int k = 0; k < NPART_ARRAYS; ++k) {
                a.src[k] = (const Real *)s->part[src][k];
                a.dst[k] = (Real *)s->part[dst][k];
            }
            a.alive_src = s->alive[src]; a.alive_dst = s->alive[dst];
            a.id_src = s->pid[src]; a.id_dst = s->pid[dst];
            a.key = s->key; a.cursor = s->cursor; a.n = s->n;
            Bracket b(s, "scatter");
            scatter_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        s->cur = dst;
        return (int)FSIM_OK;
    });
    if (rc != FSIM_OK) return rc;
    s->sorted = true;
    s->ids_identity = false;
    s->steps_since_sort = 0;
    return FSIM_OK;
}

}  // namespace fsim
