// sort.cu -- out-of-place counting sort of the particle storage by gather cell (sm_100a).
//
// Why: the reference gathers R1..R3,A at the particle's cell through four uncoalesced texture
// fetches per fragment (step_velocity_frag, empic.js:763-766) and deposits by 121 blended
// fragments per particle (programMoments01, empic.js:1473-1478).  Keeping the storage ordered
// by cell turns the gather into a streaming read of the cell table and gives every cell a
// contiguous particle segment for the deterministic per-cell sums (deposit.cu).
//
// Every frame: (1) key + histogram (emitted by the second half-step's push, or by prepass_kernel),
// (2) exclusive scan of the per-cell counts, (3) a 4-byte INDEX scatter: perm[] lists the
// particle slots in cell order.  Every `sort_interval` frames the storage itself is permuted
// (apply_perm_kernel) so that the push's cell-table gather stays a near-streaming read.  The
// order INSIDE a cell segment is not fixed (atomic cursors); determinism of the deposit comes
// from summing each segment in ascending particle-id order (deposit.cu), and the accessors
// un-permute by id, so nothing observable depends on it.
#include "common.cuh"

namespace fsim {

// Deposit prepass from the stored state (used when the push did not emit it: first density(),
// after set()/half_step(), slab mode): sort key (+ clipped flag), sprite colour, histogram.
template <typename Real>
struct PrepassArgs {
    const Real *x, *y, *z, *vx, *vy, *vz;
    uint32_t *key, *counts;
    Real *dcol[2];
    int64_t n;
    const uint32_t *n_dev;
    int nr, nz, row0, rows, own_lo, own_hi;
};

template <typename Real>
__global__ void __launch_bounds__(256) prepass_kernel(const PrepassArgs<Real> a)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < live_count(a.n_dev, a.n);
    uint32_t c = 0xffffffffu;  // lanes past the end form their own (ignored) group
    if (valid) {
        const Real xx = a.x[p], yy = a.y[p];
        const Real r = fsqrt(xx * xx + yy * yy);
        Real c0, c1, c2;
        const uint32_t key = sprite_key_colour<Real>(xx, yy, a.z[p], r, a.vx[p], a.vy[p], a.vz[p], a.nr,
                                                     a.nz, a.row0, a.rows, a.own_lo, a.own_hi, c0, c1, c2);
        a.key[p] = key;
        a.dcol[0][p] = c0; a.dcol[1][p] = c1;  // c2 = 0.001 v_z: formed by the per-cell pass from v_z itself
        (void)c2;
        c = key & KEY_MASK;
    }
    // warp-aggregated histogram: one atomic per run of equal keys
    int leader;
    uint32_t len, rank;
    warp_runs(c, (int)(threadIdx.x & 31), leader, len, rank);
    if (valid && rank == 0) atomicAdd(a.counts + c, len);
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

// 8 consecutive counts per thread as two 128-bit loads (the arrays are padded to a multiple of 8)
__device__ __forceinline__ void load8(const uint32_t *p, int64_t base, int64_t m, uint32_t (&v)[SCAN_ITEMS])
{
    if (base + SCAN_ITEMS <= m) {
        const uint4 a = *reinterpret_cast<const uint4 *>(p + base);
        const uint4 b = *reinterpret_cast<const uint4 *>(p + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < m) ? p[base + k] : 0u;
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_tile_sums_kernel(const uint32_t *__restrict__ counts, int64_t m, uint32_t *__restrict__ tile_sums)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    load8(counts, base, m, v);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
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
    load8(counts, base, m, v);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
    uint32_t run = tile_offsets[blockIdx.x] + block_exclusive_scan(s, nullptr);
    uint32_t o[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        o[k] = run;
        run += v[k];
    }
    if (base + SCAN_ITEMS <= m) {
        const uint4 a = make_uint4(o[0], o[1], o[2], o[3]), b = make_uint4(o[4], o[5], o[6], o[7]);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(starts + base) = a;
        *reinterpret_cast<uint4 *>(starts + base + 4) = b;
        *reinterpret_cast<uint4 *>(cursor + base) = a;
        *reinterpret_cast<uint4 *>(cursor + base + 4) = b;
        *reinterpret_cast<uint4 *>(counts + base) = z;
        *reinterpret_cast<uint4 *>(counts + base + 4) = z;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < m) {
                starts[base + k] = o[k];
                cursor[base + k] = o[k];
                counts[base + k] = 0;
            }
    }
    if (base <= m - 1 && m - 1 < base + SCAN_ITEMS) starts[m] = run;  // total (run after the last item)
}

// physical re-sort: dst[j] = src[perm[j]] -- gathered reads (the storage is nearly ordered, so
// they stay within a few lines), fully coalesced writes, no atomics.
template <typename Real>
struct PermuteArgs {
    const Real *src[NPART_ARRAYS];
    Real *dst[NPART_ARRAYS];
    const uint8_t *alive_src;
    uint8_t *alive_dst;
    const uint32_t *id_src;
    uint32_t *id_dst;
    const uint32_t *perm;
    int64_t n;
    const uint32_t *n_dev;
};

template <typename Real>
__global__ void __launch_bounds__(256) apply_perm_kernel(const PermuteArgs<Real> a)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= live_count(a.n_dev, a.n)) return;
    const size_t p = a.perm[j] & KEY_MASK;
    FSIM_ASSERT((int64_t)p < live_count(a.n_dev, a.n));
    Real v[NPART_ARRAYS];
#pragma unroll
    for (int k = 0; k < NPART_ARRAYS; ++k) v[k] = a.src[k][p];
    const uint8_t al = a.alive_src[p];
    const uint32_t id = a.id_src[p];
#pragma unroll
    for (int k = 0; k < NPART_ARRAYS; ++k) __stcs(a.dst[k] + j, v[k]);
    a.alive_dst[j] = al;
    a.id_dst[j] = id;
}

// index sort: perm[slot in cell order] = storage slot; nothing but 4 bytes per particle moves.
// One cursor atomic per run of equal keys in a warp; the lanes of a run take consecutive slots.
constexpr int IDX_ITEMS = 4;  // 32-particle chunks per warp: their atomics' round trips overlap

__global__ void __launch_bounds__(256)
index_scatter_kernel(const uint32_t *__restrict__ key, uint32_t *__restrict__ cursor,
                     uint32_t *__restrict__ perm, int64_t n_host, const uint32_t *__restrict__ n_dev)
{
    const int64_t n = live_count(n_dev, n_host);
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t p0 = warp * (32 * IDX_ITEMS) + lane;
    uint32_t c[IDX_ITEMS], base[IDX_ITEMS], rank[IDX_ITEMS];
    int leader[IDX_ITEMS];
    uint32_t flag[IDX_ITEMS];  // the clipped bit travels in bit 31 of the perm entry
#pragma unroll
    for (int k = 0; k < IDX_ITEMS; ++k) {
        const int64_t p = p0 + 32 * k;
        const uint32_t kk = (p < n) ? key[p] : 0xffffffffu;
        c[k] = (p < n) ? (kk & KEY_MASK) : 0xffffffffu;
        flag[k] = kk & KEY_CLIPPED;
    }
#pragma unroll
    for (int k = 0; k < IDX_ITEMS; ++k) {
        uint32_t len;
        warp_runs(c[k], lane, leader[k], len, rank[k]);
        base[k] = 0;
        if (c[k] != 0xffffffffu && rank[k] == 0) base[k] = atomicAdd(cursor + c[k], len);
    }
#pragma unroll
    for (int k = 0; k < IDX_ITEMS; ++k) {
        const uint32_t b = __shfl_sync(0xffffffffu, base[k], leader[k]);
        FSIM_ASSERT(c[k] == 0xffffffffu || (int64_t)b + rank[k] < n);  // the cursor stays inside its cell's segment
        if (c[k] != 0xffffffffu) perm[(size_t)b + rank[k]] = (uint32_t)(p0 + 32 * k) | flag[k];
    }
}

// key[] + colour + histogram from the stored state
int launch_keys(fsim_sim *s)
{
    if (s->counts_dirty) {
        FSIM_CUDA(cudaMemsetAsync(s->counts, 0, sizeof(uint32_t) * (s->ncell_local + 1), s->stream));
        s->counts_dirty = false;
    }
    if (s->n) {
        int rc = dispatch(s, [&](auto tag) {
            using Real = decltype(tag);
            const int c = s->cur;
            PrepassArgs<Real> a;
            a.x = (const Real *)s->part[c][AX]; a.y = (const Real *)s->part[c][AY];
            a.z = (const Real *)s->part[c][AZ]; a.vx = (const Real *)s->part[c][AVX];
            a.vy = (const Real *)s->part[c][AVY]; a.vz = (const Real *)s->part[c][AVZ];
            a.key = s->key; a.counts = s->counts;
            for (int q = 0; q < 2; ++q) a.dcol[q] = (Real *)s->dcol[q];
            a.n = s->n; a.nr = s->nr; a.nz = s->nz; a.row0 = s->row0; a.rows = s->rows;
            a.n_dev = s->n_async ? s->mscratch + MC_NLIVE : nullptr;
            a.own_lo = s->own0 - s->row0; a.own_hi = a.own_lo + s->own_rows;
            Bracket b(s, "prepass");
            prepass_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
            return (int)FSIM_OK;
        });
        FSIM_TRY(rc);
    }
    s->keys_valid = true;
    s->counts_dirty = true;
    return FSIM_OK;
}

// scan of the histogram, then the 4-byte index scatter.  Requires keys_valid.
int launch_bin(fsim_sim *s)
{
    const int64_t m = s->ncell_local;
    const int ntiles = (int)((m + SCAN_TILE - 1) / SCAN_TILE);
    {
        Bracket b(s, "scan");
        scan_tile_sums_kernel<<<ntiles, SCAN_BLOCK, 0, s->stream>>>(s->counts, m, s->blocksums);
        scan_tile_offsets_kernel<<<1, SCAN_BLOCK, 0, s->stream>>>(s->blocksums, ntiles);
        scan_final_kernel<<<ntiles, SCAN_BLOCK, 0, s->stream>>>(s->counts, m, s->blocksums, s->starts, s->cursor);
        s->launches += 2;
        FSIM_CUDA(cudaGetLastError());
    }
    s->counts_dirty = false;  // scan_final zeroed counts[]
    if (s->n) {
        Bracket b(s, "index_scatter");
        index_scatter_kernel<<<grid_for((s->n + IDX_ITEMS - 1) / IDX_ITEMS, 256), 256, 0, s->stream>>>(
            s->key, s->cursor, s->perm, s->n, s->n_async ? s->mscratch + MC_NLIVE : nullptr);
        FSIM_CUDA(cudaGetLastError());
    }
    s->binned = true;
    return FSIM_OK;
}

// Physical re-sort of the particle storage into cell order.  Requires binned.  key[] / dcol[] are
// per-slot and become stale, so the binning is invalidated (the next push re-emits them).
int launch_apply_perm(fsim_sim *s)
{
    if (s->n) {
        int rc = dispatch(s, [&](auto tag) {
            using Real = decltype(tag);
            const int src = s->cur, dst = s->cur ^ 1;
            PermuteArgs<Real> a;
            for (int k = 0; k < NPART_ARRAYS; ++k) {
                a.src[k] = (const Real *)s->part[src][k];
                a.dst[k] = (Real *)s->part[dst][k];
            }
            a.alive_src = s->alive[src]; a.alive_dst = s->alive[dst];
            a.id_src = s->pid[src]; a.id_dst = s->pid[dst];
            a.perm = s->perm; a.n = s->n;
            a.n_dev = s->n_async ? s->mscratch + MC_NLIVE : nullptr;
            Bracket b(s, "permute");
            apply_perm_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
            s->cur = dst;
            return (int)FSIM_OK;
        });
        FSIM_TRY(rc);
    }
    s->binned = false;
    s->keys_valid = false;
    s->have_leavers = false;
    s->ever_sorted = true;
    s->ids_identity = false;
    s->steps_since_sort = 0;
    return FSIM_OK;
}

}  // namespace fsim
