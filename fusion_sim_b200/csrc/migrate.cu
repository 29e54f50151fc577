// migrate.cu -- multi-GPU slab exchange: one process per GPU, the grid is
// cut into slabs of rows along z (index i + j*nr, empic.js:1162, makes a slab contiguous).
//
//  * particle migration: a particle whose row floor(z*nz) is not owned by this rank (it drifted
//    across the slab boundary, or it respawned anywhere: the source pdf is global, empic.js:717)
//    is packed into a device buffer grouped by destination rank; the caller moves the groups with
//    an all-to-all-v (NCCL through torch.distributed) and hands the arrivals to
//    fsim_migrate_unpack, which fills the holes and compacts the storage.  Records carry the global
//    particle id, so the id-ordered deposit stays bit-identical to a single-GPU run.
//  * halo of the per-cell sums: 5 rows (footprint radius, empic.js:949-952) each side.
//
// Record layout: NPART_ARRAYS reals (x y z vx vy vz q0..q3), then u32 id, u32 alive.
#include <vector>

#include "common.cuh"

namespace fsim {

struct RankBounds {
    int n;
    int self;
    int lo[MAX_RANKS + 1];  // rank k owns rows [lo[k], lo[k+1])
};

template <typename Real>
__device__ __forceinline__ int dest_rank(const RankBounds &rb, Real z, int nz)
{
    const int row = tex_idx(z, nz);  // NaN -> row 0: same rule as every texture fetch
    int d = 0;
    while (d + 1 < rb.n && row >= rb.lo[d + 1]) ++d;
    return d;
}

// fallback when the push did not emit the list: scan all particles for rows that are not owned
template <typename Real>
__global__ void __launch_bounds__(256)
find_leavers_kernel(const Real *__restrict__ z, int64_t n, const uint32_t *__restrict__ n_dev, int nz, int own0,
                    int own_rows, uint32_t *__restrict__ list, uint32_t *__restrict__ nlist)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= live_count(n_dev, n)) return;
    const int gj = tex_idx(z[p], nz);
    if (gj < own0 || gj >= own0 + own_rows) list[atomicAdd(nlist, 1u)] = (uint32_t)p;
}

template <typename Real>
__global__ void __launch_bounds__(256)
migrate_count_kernel(const Real *__restrict__ z, const uint32_t *__restrict__ list,
                     const uint32_t *__restrict__ nlist_d, int nz, RankBounds rb, uint32_t *__restrict__ counts)
{
    // the list length stays on the device (no host round trip before this launch): grid-stride
    const uint32_t nlist = *nlist_d;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nlist; t += gridDim.x * blockDim.x)
        atomicAdd(counts + dest_rank(rb, z[list[t]], nz), 1u);
}

template <typename Real>
struct PackArgs {
    const Real *src[NPART_ARRAYS];
    const uint8_t *alive;
    const uint32_t *id;
    unsigned char *buf;
    uint32_t *cursor;        // [nranks] running offsets (records) into buf, initialised to the group starts
    const uint32_t *list;    // slots that leave (they become the holes)
    uint32_t nlist;
    int64_t n;               // live slots (debug-build index checks)
    uint8_t *hole_flag;
    const uint32_t *key;     // non-null: keep the sort histogram consistent
    uint32_t *counts;
    int nz;
    RankBounds rb;
};

template <typename Real>
__global__ void __launch_bounds__(256) migrate_pack_kernel(const PackArgs<Real> a)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nlist) return;
    const size_t p = a.list[t];
    FSIM_ASSERT((int64_t)p < a.n);
    const int d = dest_rank(a.rb, a.src[AZ][p], a.nz);
    const uint32_t slot = atomicAdd(a.cursor + d, 1u);
    FSIM_ASSERT(slot < a.nlist);  // the groups were sized by migrate_count_kernel
    unsigned char *rec = a.buf + (size_t)slot * (NPART_ARRAYS * sizeof(Real) + 8);
    Real *r = reinterpret_cast<Real *>(rec);
#pragma unroll
    for (int k = 0; k < NPART_ARRAYS; ++k) r[k] = a.src[k][p];
    uint32_t *u = reinterpret_cast<uint32_t *>(rec + NPART_ARRAYS * sizeof(Real));
    u[0] = a.id[p];
    u[1] = a.alive[p];
    a.hole_flag[p] = 1;
    if (a.key) atomicSub(a.counts + (a.key[p] & KEY_MASK), 1u);
}

template <typename Real>
struct UnpackArgs {
    Real *dst[NPART_ARRAYS];
    uint8_t *alive;
    uint32_t *id;
    const unsigned char *buf;
    const uint32_t *holes;
    uint8_t *hole_flag;
    uint32_t nholes;
    int64_t n_old;
    int64_t nrecv;
    // non-null key: arrivals get their deposit prepass here (key, colour, histogram)
    uint32_t *key, *counts;
    Real *dcol[2];
    int nr, nz, row0, rows, own_lo, own_hi;
};

// arrival i goes into hole i while holes last, then to the end of the storage
template <typename Real>
__global__ void __launch_bounds__(256) migrate_unpack_kernel(const UnpackArgs<Real> a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nrecv) return;
    size_t slot;
    if (i < (int64_t)a.nholes) {
        slot = a.holes[i];
        FSIM_ASSERT((int64_t)slot < a.n_old);
        a.hole_flag[slot] = 0;
    } else {
        slot = (size_t)(a.n_old + (i - (int64_t)a.nholes));
    }
    const unsigned char *rec = a.buf + (size_t)i * (NPART_ARRAYS * sizeof(Real) + 8);
    const Real *r = reinterpret_cast<const Real *>(rec);
    Real v[NPART_ARRAYS];
#pragma unroll
    for (int k = 0; k < NPART_ARRAYS; ++k) {
        v[k] = r[k];
        a.dst[k][slot] = v[k];
    }
    const uint32_t *u = reinterpret_cast<const uint32_t *>(rec + NPART_ARRAYS * sizeof(Real));
    a.id[slot] = u[0];
    a.alive[slot] = (uint8_t)u[1];
    if (a.key) {
        const Real rr = fsqrt(v[AX] * v[AX] + v[AY] * v[AY]);
        Real c0, c1, c2;
        const uint32_t key = sprite_key_colour<Real>(v[AX], v[AY], v[AZ], rr, v[AVX], v[AVY], v[AVZ], a.nr, a.nz,
                                                     a.row0, a.rows, a.own_lo, a.own_hi, c0, c1, c2);
        a.key[slot] = key;
        a.dcol[0][slot] = c0; a.dcol[1][slot] = c1;
        (void)c2;
        atomicAdd(a.counts + (key & KEY_MASK), 1u);
    }
}

// more leavers than arrivals: the storage shrinks to n_new; live particles in the tail
// [n_new, n_old) move into the unfilled holes below n_new.
__global__ void __launch_bounds__(256)
compact_lists_kernel(const uint32_t *__restrict__ holes, uint32_t nholes, uint32_t first_unfilled,
                     const uint8_t *__restrict__ hole_flag, int64_t n_new, int64_t n_old,
                     uint32_t *__restrict__ targets, uint32_t *__restrict__ ntargets,
                     uint32_t *__restrict__ sources, uint32_t *__restrict__ nsources)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nh = (int64_t)nholes - first_unfilled;
    if (t < nh) {
        const uint32_t slot = holes[first_unfilled + t];
        if ((int64_t)slot < n_new) targets[atomicAdd(ntargets, 1u)] = slot;
    }
    const int64_t q = n_new + t;
    if (q < n_old && !hole_flag[q]) sources[atomicAdd(nsources, 1u)] = (uint32_t)q;
}

template <typename Real>
struct MoveArgs {
    Real *a[NPART_ARRAYS];
    uint8_t *alive;
    uint32_t *id;
    const uint32_t *targets, *sources;
    const uint32_t *ntargets, *nsources;
    int64_t n_old;  // slots in use before the compaction (upper bound; debug-build index checks)
    uint32_t *key;  // non-null: the per-slot prepass data moves along
    Real *dcol[2];
};

template <typename Real>
__global__ void __launch_bounds__(256) compact_move_kernel(const MoveArgs<Real> m)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *m.ntargets) return;
    const size_t d = m.targets[i], s = m.sources[i];
    FSIM_ASSERT(i < *m.nsources && (int64_t)d < m.n_old && (int64_t)s < m.n_old && d < s);  // a tail particle moves DOWN into a hole
#pragma unroll
    for (int k = 0; k < NPART_ARRAYS; ++k) m.a[k][d] = m.a[k][s];
    m.alive[d] = m.alive[s];
    m.id[d] = m.id[s];
    if (m.key) {
        m.key[d] = m.key[s];
#pragma unroll
        for (int q = 0; q < 2; ++q) m.dcol[q][d] = m.dcol[q][s];
    }
}


// ---- asynchronous fixed-capacity exchange: every count stays on the device ----------------------------
struct PlanDev {  // the part of MigratePlan the kernels need (by value: ~1.3 KB of kernel parameters)
    int nranks, self;
    int lo[MAX_RANKS + 1];
    uint32_t send_cap[MAX_RANKS], recv_cap[MAX_RANKS], recv_slot0[MAX_RANKS + 1];
    uint64_t send_off[MAX_RANKS], recv_off[MAX_RANKS];
};

template <typename Real>
__device__ __forceinline__ int dest_rank_plan(const PlanDev &pl, Real z, int nz)
{
    const int row = tex_idx(z, nz);
    int d = 0;
    while (d + 1 < pl.nranks && row >= pl.lo[d + 1]) ++d;
    return d;
}

template <typename Real>
struct PackFixedArgs {
    const Real *src[NPART_ARRAYS];
    const uint8_t *alive;
    const uint32_t *id;
    unsigned char *send;     // regions of PlanDev::send_off, each: 16-byte header + records
    const uint32_t *list;    // leaver list (perm[]), length ctr[MC_NLEAVERS]
    uint32_t *holes;         // packed leavers' slots
    uint32_t holes_cap;      // entries `holes` can take (= sum of the send capacities)
    uint8_t *hole_flag;
    uint32_t *ctr;           // mscratch
    const uint32_t *key;     // non-null: keep the sort histogram consistent
    uint32_t *counts;
    int nz;
};

// grid-stride over the leaver list whose length is read on the device
template <typename Real>
__global__ void __launch_bounds__(256) migrate_pack_fixed_kernel(const PackFixedArgs<Real> a, const PlanDev pl)
{
    constexpr size_t REC = NPART_ARRAYS * sizeof(Real) + 8;
    const uint32_t nlist = a.ctr[MC_NLEAVERS];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nlist; t += gridDim.x * blockDim.x) {
        const size_t p = a.list[t];
        FSIM_ASSERT(p < a.ctr[MC_NLIVE]);
        const int d = dest_rank_plan(pl, a.src[AZ][p], a.nz);
        const uint32_t slot = atomicAdd(a.ctr + MC_CURSOR + d, 1u);
        if (d == pl.self || slot >= pl.send_cap[d]) {  // region full: the particle stays, the run is flagged invalid
            atomicOr(a.ctr + MC_ERR, MERR_SEND_OVERFLOW);
            continue;
        }
        unsigned char *rec = a.send + pl.send_off[d] + MIGRATE_HEADER_BYTES + (size_t)slot * REC;
        Real *r = reinterpret_cast<Real *>(rec);
#pragma unroll
        for (int k = 0; k < NPART_ARRAYS; ++k) r[k] = a.src[k][p];
        uint32_t *u = reinterpret_cast<uint32_t *>(rec + NPART_ARRAYS * sizeof(Real));
        u[0] = a.id[p];
        u[1] = a.alive[p];
        const uint32_t hslot = atomicAdd(a.ctr + MC_NHOLES, 1u);
        FSIM_ASSERT(hslot < a.holes_cap);
        a.holes[hslot] = (uint32_t)p;
        a.hole_flag[p] = 1;
        if (a.key) atomicSub(a.counts + (a.key[p] & KEY_MASK), 1u);
    }
}

// one block: record counts into the send headers, statistics
__global__ void __launch_bounds__(64) migrate_send_headers_kernel(unsigned char *send, uint32_t *ctr, const PlanDev pl)
{
    const int d = threadIdx.x;
    uint32_t c = 0;
    if (d < pl.nranks) {
        c = min(ctr[MC_CURSOR + d], pl.send_cap[d]);
        if (d == pl.self) c = 0;
        uint32_t *h = reinterpret_cast<uint32_t *>(send + pl.send_off[d]);
        h[0] = c; h[1] = 0; h[2] = 0; h[3] = 0;
    }
    // total packed (64-bit counter in two words; single block, so a plain read-modify-write)
    uint32_t tot = c;
    for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    __shared__ uint32_t part[2];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t old = ((uint64_t)ctr[MC_SENT_HI] << 32) | ctr[MC_SENT_LO];
        const uint64_t now = old + part[0] + part[1];
        ctr[MC_SENT_LO] = (uint32_t)now;
        ctr[MC_SENT_HI] = (uint32_t)(now >> 32);
    }
}

// one block: prefix of the received counts, new particle count, capacity check
__global__ void __launch_bounds__(64)
migrate_recv_headers_kernel(const unsigned char *recv, uint32_t *ctr, const PlanDev pl, uint32_t max_live)
{
    if (threadIdx.x != 0) return;
    uint32_t run = 0;
    for (int k = 0; k < pl.nranks; ++k) {
        ctr[MC_PREFIX + k] = run;
        uint32_t c = (k == pl.self) ? 0u : reinterpret_cast<const uint32_t *>(recv + pl.recv_off[k])[0];
        run += min(c, pl.recv_cap[k]);
    }
    ctr[MC_PREFIX + pl.nranks] = run;
    const uint32_t n_old = ctr[MC_NLIVE], nholes = ctr[MC_NHOLES];
    uint32_t nrecv = run;
    if ((uint64_t)n_old - nholes + nrecv > max_live) {  // does not fit: drop the excess, flag the run invalid
        atomicOr(ctr + MC_ERR, MERR_CAPACITY);
        nrecv = max_live - (n_old - nholes);
    }
    ctr[MC_NRECV] = nrecv;
    ctr[MC_NOLD] = n_old;
    ctr[MC_NNEW] = n_old - nholes + nrecv;
    ctr[MC_NTARGETS] = 0;
    ctr[MC_NSOURCES] = 0;
}

template <typename Real>
struct UnpackFixedArgs {
    Real *dst[NPART_ARRAYS];
    uint8_t *alive;
    uint32_t *id;
    const unsigned char *recv;
    const uint32_t *holes;
    uint8_t *hole_flag;
    const uint32_t *ctr;
    uint32_t *key, *counts;  // non-null key: arrivals get their deposit prepass here
    Real *dcol[2];
    int nr, nz, row0, rows, own_lo, own_hi;
};

// one thread per receive SLOT (sum of the capacities, known to the host); arrival i goes into hole i
// while holes last, then to the end of the storage
template <typename Real>
__global__ void __launch_bounds__(256) migrate_unpack_fixed_kernel(const UnpackFixedArgs<Real> a, const PlanDev pl)
{
    constexpr size_t REC = NPART_ARRAYS * sizeof(Real) + 8;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= pl.recv_slot0[pl.nranks]) return;
    int k = 0;
    while (t >= pl.recv_slot0[k + 1]) ++k;
    const uint32_t j = t - pl.recv_slot0[k];
    if (j >= a.ctr[MC_PREFIX + k + 1] - a.ctr[MC_PREFIX + k]) return;
    const uint32_t i = a.ctr[MC_PREFIX + k] + j;
    if (i >= a.ctr[MC_NRECV]) return;  // capacity overflow (flagged)
    const uint32_t nholes = a.ctr[MC_NHOLES];
    size_t slot;
    if (i < nholes) {
        slot = a.holes[i];
        a.hole_flag[slot] = 0;
    } else {
        slot = (size_t)a.ctr[MC_NOLD] + (i - nholes);
    }
    FSIM_ASSERT(slot < a.ctr[MC_NNEW] || slot < a.ctr[MC_NOLD]);
    const unsigned char *rec = a.recv + pl.recv_off[k] + MIGRATE_HEADER_BYTES + (size_t)j * REC;
    const Real *r = reinterpret_cast<const Real *>(rec);
    Real v[NPART_ARRAYS];
#pragma unroll
    for (int q = 0; q < NPART_ARRAYS; ++q) {
        v[q] = r[q];
        a.dst[q][slot] = v[q];
    }
    const uint32_t *u = reinterpret_cast<const uint32_t *>(rec + NPART_ARRAYS * sizeof(Real));
    a.id[slot] = u[0];
    a.alive[slot] = (uint8_t)u[1];
    if (a.key) {
        const Real rr = fsqrt(v[AX] * v[AX] + v[AY] * v[AY]);
        Real c0, c1, c2;
        const uint32_t key = sprite_key_colour<Real>(v[AX], v[AY], v[AZ], rr, v[AVX], v[AVY], v[AVZ], a.nr, a.nz,
                                                     a.row0, a.rows, a.own_lo, a.own_hi, c0, c1, c2);
        a.key[slot] = key;
        a.dcol[0][slot] = c0; a.dcol[1][slot] = c1;
        (void)c2;
        atomicAdd(a.counts + (key & KEY_MASK), 1u);
    }
}

// more leavers than arrivals: the storage shrinks to n_new; one thread per possible hole
__global__ void __launch_bounds__(256)
compact_lists_fixed_kernel(const uint32_t *__restrict__ holes, const uint8_t *__restrict__ hole_flag, uint32_t *ctr,
                           uint32_t *__restrict__ targets, uint32_t *__restrict__ sources, uint32_t span_max)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nholes = ctr[MC_NHOLES], nrecv = ctr[MC_NRECV];
    if (nholes <= nrecv || t >= span_max) return;
    const int64_t n_new = ctr[MC_NNEW], n_old = ctr[MC_NOLD];
    if (t < nholes - nrecv) {
        const uint32_t slot = holes[nrecv + t];
        FSIM_ASSERT((int64_t)slot < n_old);
        if ((int64_t)slot < n_new) {
            const uint32_t at = atomicAdd(ctr + MC_NTARGETS, 1u);
            FSIM_ASSERT(at < span_max);
            targets[at] = slot;
        }
    }
    const int64_t q = n_new + t;
    if (q < n_old && !hole_flag[q]) {
        const uint32_t at = atomicAdd(ctr + MC_NSOURCES, 1u);
        FSIM_ASSERT(at < span_max);
        sources[at] = (uint32_t)q;
    }
}

// clear the flags of the vacated slots, publish the new count, reset the per-frame counters
__global__ void __launch_bounds__(256)
migrate_finish_kernel(const uint32_t *__restrict__ holes, uint8_t *__restrict__ hole_flag, uint32_t *ctr, uint32_t span_max)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ctr[MC_NHOLES] && t < span_max) hole_flag[holes[t]] = 0;
}
__global__ void migrate_publish_kernel(uint32_t *ctr, int nranks)
{
    if (threadIdx.x == 0) {
        ctr[MC_NLIVE] = ctr[MC_NNEW];
        ctr[MC_NHOLES] = 0;
        ctr[MC_NLEAVERS] = 0;
    }
    if ((int)threadIdx.x < nranks) ctr[MC_CURSOR + threadIdx.x] = 0;
}

__global__ void set_words_kernel(uint32_t *p, uint32_t v, int n)
{
    if ((int)threadIdx.x < n) p[threadIdx.x] = v;
}

static int ensure_migr(fsim_sim *s, size_t bytes)
{
    if (bytes <= s->migr_bytes) return FSIM_OK;
    if (s->migr) FSIM_CUDA(cudaFree(s->migr));
    s->migr = nullptr;
    s->migr_bytes = 0;
    bytes = bytes + bytes / 2 + 4096;
    FSIM_CUDA(cudaMalloc(&s->migr, bytes));
    s->migr_bytes = bytes;
    return FSIM_OK;
}

}  // namespace fsim

using namespace fsim;

extern "C" {

int64_t fsim_migrate_record_bytes(const fsim_sim *s) { return s ? (int64_t)(NPART_ARRAYS * s->rs + 8) : -1; }

int fsim_migrate_pack(fsim_sim *s, const int64_t *row_bounds, int32_t nranks, int32_t self, int64_t *send_counts,
                      void **send_buf_dev)
{
    if (!s || !row_bounds || !send_counts || !send_buf_dev) {
        set_error("fsim_migrate_pack: null argument");
        return FSIM_ERR_INVALID;
    }
    if (nranks < 1 || nranks > MAX_RANKS || self < 0 || self >= nranks) {
        set_error("fsim_migrate_pack: bad rank arguments");
        return FSIM_ERR_INVALID;
    }
    if (row_bounds[self] != s->own0 || row_bounds[self + 1] != s->own0 + s->own_rows) {
        set_error("fsim_migrate_pack: row_bounds[self] does not match the slab of this handle");
        return FSIM_ERR_INVALID;
    }
    FSIM_TRY(check_handle(s));
    if (s->n_async) {
        set_error("fsim_migrate_pack: this handle uses the asynchronous exchange (fsim_migrate_setup)");
        return FSIM_ERR_STATE;
    }
    RankBounds rb;
    rb.n = nranks;
    rb.self = self;
    for (int k = 0; k <= nranks; ++k) rb.lo[k] = (int)row_bounds[k];
    uint32_t *scr = s->mscratch;  // layout: enum MC_* (common.cuh)
    uint32_t *nlist_d = scr + MC_NLEAVERS;
    FSIM_CUDA(cudaMemsetAsync(scr + MC_CURSOR, 0, sizeof(uint32_t) * (MC_WORDS - MC_CURSOR), s->stream));
    FSIM_CUDA(cudaMemsetAsync(scr + MC_NTARGETS, 0, sizeof(uint32_t) * 2, s->stream));
    for (int k = 0; k < nranks; ++k) send_counts[k] = 0;
    *send_buf_dev = nullptr;
    s->nholes_host = 0;
    s->binned = false;
    if (s->n == 0) return FSIM_OK;
    const bool keys = s->keys_valid;
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int c = s->cur;
        if (!s->have_leavers) {  // the push did not emit the list: scan the positions
            FSIM_CUDA(cudaMemsetAsync(nlist_d, 0, sizeof(uint32_t), s->stream));
            find_leavers_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
                (const Real *)s->part[c][AZ], s->n, nullptr, s->nz, s->own0, s->own_rows, s->leavers, nlist_d);
            FSIM_CUDA(cudaGetLastError());
            s->launches++;
        }
        s->have_leavers = false;
        // destination counts of the leavers and the list length: ONE read-back per frame
        migrate_count_kernel<Real><<<s->nsm * 2, 256, 0, s->stream>>>((const Real *)s->part[c][AZ], s->leavers, nlist_d, s->nz, rb,
                                                                   scr + MC_COUNTS);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        uint32_t nlist = 0;
        uint32_t hc[MAX_RANKS] = {}, off[MAX_RANKS] = {};
        FSIM_CUDA(cudaMemcpyAsync(&nlist, nlist_d, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        FSIM_CUDA(cudaMemcpyAsync(hc, scr + MC_COUNTS, sizeof(uint32_t) * nranks, cudaMemcpyDeviceToHost, s->stream));
        FSIM_CUDA(cudaStreamSynchronize(s->stream));
        if (nlist == 0) return (int)FSIM_OK;
        if ((int64_t)nlist > s->cap / 4) {  // the compaction lists of fsim_migrate_unpack hold cap/4 entries each
            set_error("fsim_migrate_pack: more than a quarter of the particle slots leave the slab at once");
            return (int)FSIM_ERR_RANGE;
        }
        int64_t total = 0;
        for (int k = 0; k < nranks; ++k) {
            send_counts[k] = hc[k];
            off[k] = (uint32_t)total;
            total += hc[k];
        }
        FSIM_TRY(ensure_migr(s, (size_t)total * (NPART_ARRAYS * sizeof(Real) + 8) + 16));
        FSIM_CUDA(cudaMemcpyAsync(scr + MC_CURSOR, off, sizeof(uint32_t) * nranks, cudaMemcpyHostToDevice, s->stream));
        PackArgs<Real> a;
        for (int k = 0; k < NPART_ARRAYS; ++k) a.src[k] = (const Real *)s->part[c][k];
        a.alive = s->alive[c];
        a.id = s->pid[c];
        a.buf = (unsigned char *)s->migr;
        a.cursor = scr + MC_CURSOR;
        a.list = s->leavers; a.nlist = nlist; a.n = s->n;
        a.hole_flag = s->hole_flag;
        a.key = keys ? s->key : nullptr;
        a.counts = s->counts;
        a.nz = s->nz; a.rb = rb;
        {
            Bracket b(s, "migrate_pack");
            migrate_pack_kernel<Real><<<grid_for(nlist, 256), 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        // an external (caller-owned) stream orders the collective after this kernel by itself
        if (!s->ext_stream) FSIM_CUDA(cudaStreamSynchronize(s->stream));
        s->nholes_host = nlist;
        *send_buf_dev = s->migr;
        return (int)FSIM_OK;
    });
}

int fsim_migrate_unpack(fsim_sim *s, const void *recv_buf_dev, int64_t nrecv)
{
    if (!s || nrecv < 0 || (nrecv > 0 && !recv_buf_dev)) {
        set_error("fsim_migrate_unpack: bad argument");
        return FSIM_ERR_INVALID;
    }
    FSIM_TRY(check_handle(s));
    const int64_t nholes = s->nholes_host;
    const int64_t n_old = s->n;
    const int64_t n_new = n_old - nholes + nrecv;
    if (n_new > s->cap - 1024) {
        set_error("fsim_migrate_unpack: arrivals exceed the particle capacity of this rank");
        return FSIM_ERR_RANGE;
    }
    uint32_t *scr = s->mscratch;
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int c = s->cur;
        if (nrecv) {
            UnpackArgs<Real> a;
            for (int k = 0; k < NPART_ARRAYS; ++k) a.dst[k] = (Real *)s->part[c][k];
            a.alive = s->alive[c]; a.id = s->pid[c];
            a.buf = (const unsigned char *)recv_buf_dev;
            a.holes = s->leavers; a.hole_flag = s->hole_flag;
            a.nholes = (uint32_t)nholes; a.n_old = n_old; a.nrecv = nrecv;
            a.key = s->keys_valid ? s->key : nullptr;
            a.counts = s->counts;
            for (int q = 0; q < 2; ++q) a.dcol[q] = (Real *)s->dcol[q];
            a.nr = s->nr; a.nz = s->nz; a.row0 = s->row0; a.rows = s->rows;
            a.own_lo = s->own0 - s->row0; a.own_hi = a.own_lo + s->own_rows;
            Bracket b(s, "migrate_unpack");
            migrate_unpack_kernel<Real><<<grid_for(nrecv, 256), 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        if (nholes > nrecv) {  // shrink: move tail particles into the remaining holes
            const int64_t span = std::max<int64_t>(nholes - nrecv, n_old - n_new);
            // target/source lists: the upper half of the leaver buffer, beyond the hole list (both are tiny)
            uint32_t *targets = s->leavers + s->cap / 2, *sources = s->leavers + s->cap / 2 + s->cap / 4;
            compact_lists_kernel<<<grid_for(span, 256), 256, 0, s->stream>>>(
                s->leavers, (uint32_t)nholes, (uint32_t)nrecv, s->hole_flag, n_new, n_old, targets,
                scr + MC_NTARGETS, sources, scr + MC_NSOURCES);
            MoveArgs<Real> m;
            for (int k = 0; k < NPART_ARRAYS; ++k) m.a[k] = (Real *)s->part[c][k];
            m.alive = s->alive[c]; m.id = s->pid[c];
            m.targets = targets; m.sources = sources; m.ntargets = scr + MC_NTARGETS; m.nsources = scr + MC_NSOURCES;
            m.n_old = n_old;
            m.key = s->keys_valid ? s->key : nullptr;
            for (int q = 0; q < 2; ++q) m.dcol[q] = (Real *)s->dcol[q];
            compact_move_kernel<Real><<<grid_for(nholes - nrecv, 256), 256, 0, s->stream>>>(m);
            FSIM_CUDA(cudaGetLastError());
            s->launches += 2;
        }
        // clear the hole flags of everything that was vacated (cheap: the flag array is 1 B/slot)
        if (nholes) FSIM_CUDA(cudaMemsetAsync(s->hole_flag, 0, (size_t)std::max(n_old, n_new), s->stream));
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    s->n = n_new;
    s->nholes_host = 0;
    s->ids_identity = false;
    s->binned = false;  // keys_valid is kept: the kernels above maintained key[], dcol[] and counts[]
    return FSIM_OK;
}


// ---- asynchronous exchange: fixed-capacity regions, counts in the region headers ----------------------
static PlanDev plan_dev(const MigratePlan &p)
{
    PlanDev d;
    d.nranks = p.nranks; d.self = p.self;
    for (int k = 0; k <= MAX_RANKS; ++k) { d.lo[k] = p.lo[k]; d.recv_slot0[k] = p.recv_slot0[k]; }
    for (int k = 0; k < MAX_RANKS; ++k) {
        d.send_cap[k] = p.send_cap[k]; d.recv_cap[k] = p.recv_cap[k];
        d.send_off[k] = p.send_off[k]; d.recv_off[k] = p.recv_off[k];
    }
    return d;
}

int fsim_migrate_setup(fsim_sim *s, const int64_t *row_bounds, int32_t nranks, int32_t self, const int64_t *send_caps,
                       const int64_t *recv_caps, void **send_buf_dev, void **recv_buf_dev, int64_t *send_region_bytes,
                       int64_t *recv_region_bytes)
{
    FSIM_TRY(check_handle(s));
    if (!row_bounds || !send_caps || !recv_caps || !send_buf_dev || !recv_buf_dev || !send_region_bytes || !recv_region_bytes) {
        set_error("fsim_migrate_setup: null argument");
        return FSIM_ERR_INVALID;
    }
    if (!s->slab || nranks < 1 || nranks > MAX_RANKS || self < 0 || self >= nranks ||
        row_bounds[self] != s->own0 || row_bounds[self + 1] != s->own0 + s->own_rows) {
        set_error("fsim_migrate_setup: rank arguments do not match the slab of this handle");
        return FSIM_ERR_INVALID;
    }
    FSIM_TRY(settle_count(s));
    MigratePlan &p = s->plan;
    for (void *q : {(void *)p.send, (void *)p.recv, (void *)p.holes, (void *)p.targets, (void *)p.sources})
        if (q) FSIM_CUDA(cudaFree(q));
    p = MigratePlan();
    p.nranks = nranks; p.self = self;
    for (int k = 0; k <= nranks; ++k) p.lo[k] = (int)row_bounds[k];
    const size_t rec = NPART_ARRAYS * s->rs + 8;
    auto region = [&](int64_t cap) { return (uint64_t)((MIGRATE_HEADER_BYTES + (size_t)cap * rec + 15) / 16 * 16); };
    for (int k = 0; k < nranks; ++k) {
        if (send_caps[k] < 0 || recv_caps[k] < 0 || send_caps[k] > (1 << 28) || recv_caps[k] > (1 << 28)) {
            set_error("fsim_migrate_setup: capacity out of range");
            return FSIM_ERR_INVALID;
        }
        p.send_cap[k] = k == self ? 0u : (uint32_t)send_caps[k];
        p.recv_cap[k] = k == self ? 0u : (uint32_t)recv_caps[k];
        p.send_off[k + 1] = p.send_off[k] + region(p.send_cap[k]);
        p.recv_off[k + 1] = p.recv_off[k] + region(p.recv_cap[k]);
        p.recv_slot0[k + 1] = p.recv_slot0[k] + p.recv_cap[k];
        p.send_total += p.send_cap[k];
        p.recv_total += p.recv_cap[k];
        send_region_bytes[k] = (int64_t)(p.send_off[k + 1] - p.send_off[k]);
        recv_region_bytes[k] = (int64_t)(p.recv_off[k + 1] - p.recv_off[k]);
    }
    FSIM_CUDA(cudaMalloc((void **)&p.send, p.send_off[nranks]));
    FSIM_CUDA(cudaMalloc((void **)&p.recv, p.recv_off[nranks]));
    FSIM_CUDA(cudaMemsetAsync(p.send, 0, p.send_off[nranks], s->stream));
    FSIM_CUDA(cudaMemsetAsync(p.recv, 0, p.recv_off[nranks], s->stream));
    const size_t nl = std::max<size_t>(p.send_total, 1);
    FSIM_CUDA(cudaMalloc((void **)&p.holes, sizeof(uint32_t) * nl));
    FSIM_CUDA(cudaMalloc((void **)&p.targets, sizeof(uint32_t) * nl));
    FSIM_CUDA(cudaMalloc((void **)&p.sources, sizeof(uint32_t) * nl));
    if (!s->n_pinned) {
        FSIM_CUDA(cudaMallocHost((void **)&s->n_pinned, 2 * sizeof(uint32_t)));
        for (int k = 0; k < 2; ++k) FSIM_CUDA(cudaEventCreateWithFlags(&s->n_event[k], cudaEventDisableTiming));
    }
    // from here on the exact particle count lives on the device
    FSIM_CUDA(cudaMemsetAsync(s->mscratch, 0, sizeof(uint32_t) * MC_WORDS, s->stream));
    set_words_kernel<<<1, 32, 0, s->stream>>>(s->mscratch + MC_NLIVE, (uint32_t)s->n, 1);
    FSIM_CUDA(cudaGetLastError());
    s->n_async = true;
    s->n_inflight[0] = s->n_inflight[1] = false;
    s->have_leavers = false;
    *send_buf_dev = p.send;
    *recv_buf_dev = p.recv;
    return FSIM_OK;
}

// Leavers -> send regions (headers carry the counts).  Nothing is read back: the caller enqueues its
// all-to-all of the (fixed-size) regions on the same stream and then calls fsim_migrate_end.
int fsim_migrate_begin(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    if (!s->n_async) {
        set_error("fsim_migrate_begin: call fsim_migrate_setup first");
        return FSIM_ERR_STATE;
    }
    const MigratePlan &p = s->plan;
    uint32_t *ctr = s->mscratch;
    s->binned = false;
    const bool keys = s->keys_valid;
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int c = s->cur;
        if (!s->have_leavers) {  // the push did not emit the list: scan the positions
            FSIM_CUDA(cudaMemsetAsync(ctr + MC_NLEAVERS, 0, sizeof(uint32_t), s->stream));
            if (s->n) {
                find_leavers_kernel<Real><<<grid_for(s->n, 256), 256, 0, s->stream>>>(
                    (const Real *)s->part[c][AZ], s->n, ctr + MC_NLIVE, s->nz, s->own0, s->own_rows, s->leavers, ctr + MC_NLEAVERS);
                FSIM_CUDA(cudaGetLastError());
                s->launches++;
            }
        }
        s->have_leavers = false;
        const PlanDev pd = plan_dev(p);
        PackFixedArgs<Real> a;
        for (int k = 0; k < NPART_ARRAYS; ++k) a.src[k] = (const Real *)s->part[c][k];
        a.alive = s->alive[c]; a.id = s->pid[c];
        a.send = p.send; a.list = s->leavers; a.holes = p.holes; a.holes_cap = p.send_total; a.hole_flag = s->hole_flag;
        a.ctr = ctr;
        a.key = keys ? s->key : nullptr;
        a.counts = s->counts;
        a.nz = s->nz;
        {
            Bracket b(s, "migrate_pack");
            migrate_pack_fixed_kernel<Real><<<s->nsm * 2, 256, 0, s->stream>>>(a, pd);
            migrate_send_headers_kernel<<<1, 64, 0, s->stream>>>(p.send, ctr, pd);
            FSIM_CUDA(cudaGetLastError());
            s->launches++;
        }
        return (int)FSIM_OK;
    });
}

// Arrivals (receive regions) -> holes / end of the storage, compaction, new count -- all from device-side counts.
int fsim_migrate_end(fsim_sim *s)
{
    FSIM_TRY(check_handle(s));
    if (!s->n_async) {
        set_error("fsim_migrate_end: call fsim_migrate_setup first");
        return FSIM_ERR_STATE;
    }
    const MigratePlan &p = s->plan;
    uint32_t *ctr = s->mscratch;
    const uint32_t max_live = (uint32_t)(s->cap - 1024);
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int c = s->cur;
        const PlanDev pd = plan_dev(p);
        Bracket b(s, "migrate_unpack");
        migrate_recv_headers_kernel<<<1, 64, 0, s->stream>>>(p.recv, ctr, pd, max_live);
        if (p.recv_total) {
            UnpackFixedArgs<Real> a;
            for (int k = 0; k < NPART_ARRAYS; ++k) a.dst[k] = (Real *)s->part[c][k];
            a.alive = s->alive[c]; a.id = s->pid[c];
            a.recv = p.recv; a.holes = p.holes; a.hole_flag = s->hole_flag; a.ctr = ctr;
            a.key = s->keys_valid ? s->key : nullptr;
            a.counts = s->counts;
            for (int q = 0; q < 2; ++q) a.dcol[q] = (Real *)s->dcol[q];
            a.nr = s->nr; a.nz = s->nz; a.row0 = s->row0; a.rows = s->rows;
            a.own_lo = s->own0 - s->row0; a.own_hi = a.own_lo + s->own_rows;
            migrate_unpack_fixed_kernel<Real><<<grid_for(p.recv_total, 256), 256, 0, s->stream>>>(a, pd);
        }
        if (p.send_total) {  // shrink when more left than arrived (decided on the device)
            compact_lists_fixed_kernel<<<grid_for(p.send_total, 256), 256, 0, s->stream>>>(
                p.holes, s->hole_flag, ctr, p.targets, p.sources, p.send_total);
            MoveArgs<Real> m;
            for (int k = 0; k < NPART_ARRAYS; ++k) m.a[k] = (Real *)s->part[c][k];
            m.alive = s->alive[c]; m.id = s->pid[c];
            m.targets = p.targets; m.sources = p.sources; m.ntargets = ctr + MC_NTARGETS; m.nsources = ctr + MC_NSOURCES;
            m.n_old = s->cap;
            m.key = s->keys_valid ? s->key : nullptr;
            for (int q = 0; q < 2; ++q) m.dcol[q] = (Real *)s->dcol[q];
            compact_move_kernel<Real><<<grid_for(p.send_total, 256), 256, 0, s->stream>>>(m);
            migrate_finish_kernel<<<grid_for(p.send_total, 256), 256, 0, s->stream>>>(p.holes, s->hole_flag, ctr, p.send_total);
        }
        migrate_publish_kernel<<<1, 64, 0, s->stream>>>(ctr, p.nranks);
        FSIM_CUDA(cudaGetLastError());
        s->launches += 5;
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    // Host-side UPPER bound of the live slots (grid sizes): the newest read-back that has arrived plus
    // everything that may have come in since; the read-back of this frame's count is queued, never awaited.
    for (int k = 0; k < 2; ++k)
        if (s->n_inflight[k]) s->n_pending_bound[k] += p.recv_total;
    int64_t bound = std::min<int64_t>(s->n + (int64_t)p.recv_total, s->cap - 1024);
    for (int k = 0; k < 2; ++k)
        if (s->n_inflight[k] && cudaEventQuery(s->n_event[k]) == cudaSuccess) {
            bound = std::min<int64_t>(bound, (int64_t)s->n_pinned[k] + s->n_pending_bound[k]);
            s->n_inflight[k] = false;
        }
    s->n = bound;
    const int slot = s->n_slot ^= 1;
    if (!s->n_inflight[slot]) {
        FSIM_CUDA(cudaMemcpyAsync(s->n_pinned + slot, ctr + MC_NLIVE, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        FSIM_CUDA(cudaEventRecord(s->n_event[slot], s->stream));
        s->n_inflight[slot] = true;
        s->n_pending_bound[slot] = 0;
    }
    s->ids_identity = false;
    s->binned = false;  // keys_valid is kept: the kernels above maintained key[], dcol[] and counts[]
    return FSIM_OK;
}

// records packed so far (statistics; synchronises)
int fsim_migrate_stats(fsim_sim *s, int64_t *sent_total)
{
    FSIM_TRY(check_handle(s));
    if (!sent_total) {
        set_error("fsim_migrate_stats: null argument");
        return FSIM_ERR_INVALID;
    }
    uint32_t w[2] = {0, 0};
    FSIM_CUDA(cudaMemcpyAsync(w, s->mscratch + MC_SENT_LO, sizeof w, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    *sent_total = (int64_t)(((uint64_t)w[1] << 32) | w[0]);
    return FSIM_OK;
}

int fsim_halo_ptrs(fsim_sim *s, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, int64_t *bytes_each)
{
    if (!s || !send_lo || !send_hi || !recv_lo || !recv_hi || !bytes_each) {
        set_error("fsim_halo_ptrs: null argument");
        return FSIM_ERR_INVALID;
    }
    FSIM_TRY(check_handle(s));
    const size_t each = (size_t)4 * FSIM_SHAPE_MID * s->nr * s->rs;
    if (!s->halo_buf) {
        FSIM_CUDA(cudaMalloc(&s->halo_buf, 4 * each));
        FSIM_CUDA(cudaMemsetAsync(s->halo_buf, 0, 4 * each, s->stream));  // ordered before the first halo pack by the stream
    }
    unsigned char *b = (unsigned char *)s->halo_buf;
    const int o0 = s->own0 - s->row0;  // first owned row, local index
    *bytes_each = (int64_t)each;
    *send_lo = b;
    *send_hi = b + each;
    *recv_lo = (o0 >= FSIM_SHAPE_MID) ? b + 2 * each : nullptr;                           // a lower neighbour exists
    *recv_hi = (o0 + s->own_rows + FSIM_SHAPE_MID <= s->rows) ? b + 3 * each : nullptr;   // an upper neighbour exists
    return FSIM_OK;
}

}  // extern "C"

namespace fsim {

// boundary rows of the planar per-cell sums <-> contiguous exchange buffers [channel][5 rows][nr]
template <typename Real>
__global__ void __launch_bounds__(256)
halo_copy_kernel(Real *__restrict__ S, Real *__restrict__ buf, int nr, int pitch, int64_t plane, int row_first,
                 int to_buffer)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)FSIM_SHAPE_MID * nr;
    if (t >= 4 * per) return;
    const int q = (int)(t / per), r = (int)((t % per) / nr), i = (int)(t % nr);
    Real *cell = S + q * plane + (size_t)(row_first + r) * pitch + i;
    if (to_buffer) buf[t] = *cell;
    else *cell = buf[t];
}

static int halo_copy(fsim_sim *s, int which, int row_first, int to_buffer)
{
    const size_t each = (size_t)4 * FSIM_SHAPE_MID * s->nr * s->rs;
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const int64_t n = 4ll * FSIM_SHAPE_MID * s->nr;
        halo_copy_kernel<Real><<<grid_for(n, 256), 256, 0, s->stream>>>(
            (Real *)s->cellsum, (Real *)((unsigned char *)s->halo_buf + which * each), s->nr, s->pitch, s->plane,
            row_first, to_buffer);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
}

int launch_halo_pack(fsim_sim *s)
{
    void *p[4];
    int64_t n;
    FSIM_TRY(fsim_halo_ptrs(s, &p[0], &p[1], &p[2], &p[3], &n));
    const int o0 = s->own0 - s->row0;
    FSIM_TRY(halo_copy(s, 0, o0, 1));                                   // first 5 owned rows -> send_lo
    return halo_copy(s, 1, o0 + s->own_rows - FSIM_SHAPE_MID, 1);       // last 5 owned rows  -> send_hi
}

int launch_halo_unpack(fsim_sim *s)
{
    void *p[4];
    int64_t n;
    FSIM_TRY(fsim_halo_ptrs(s, &p[0], &p[1], &p[2], &p[3], &n));
    const int o0 = s->own0 - s->row0;
    if (p[2]) FSIM_TRY(halo_copy(s, 2, o0 - FSIM_SHAPE_MID, 0));        // recv_lo -> 5 rows below the slab
    if (p[3]) FSIM_TRY(halo_copy(s, 3, o0 + s->own_rows, 0));           // recv_hi -> 5 rows above the slab
    return FSIM_OK;
}

// asynchronous exchange: the host's n is an upper bound between synchronisation points; make it exact
int settle_count(fsim_sim *s)
{
    if (!s->n_async) return FSIM_OK;
    uint32_t n = 0;
    FSIM_CUDA(cudaMemcpyAsync(&n, s->mscratch + MC_NLIVE, sizeof n, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));
    s->n = (int64_t)n;
    s->n_inflight[0] = s->n_inflight[1] = false;
    return FSIM_OK;
}

}  // namespace fsim
