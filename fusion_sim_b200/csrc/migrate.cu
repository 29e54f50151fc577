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

constexpr int MAX_RANKS = 64;

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
find_leavers_kernel(const Real *__restrict__ z, int64_t n, int nz, int own0, int own_rows,
                    uint32_t *__restrict__ list, uint32_t *__restrict__ nlist)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
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
    const int d = dest_rank(a.rb, a.src[AZ][p], a.nz);
    const uint32_t slot = atomicAdd(a.cursor + d, 1u);
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
    const uint32_t *ntargets;
    uint32_t *key;  // non-null: the per-slot prepass data moves along
    Real *dcol[2];
};

template <typename Real>
__global__ void __launch_bounds__(256) compact_move_kernel(const MoveArgs<Real> m)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *m.ntargets) return;
    const size_t d = m.targets[i], s = m.sources[i];
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
    FSIM_CUDA(cudaSetDevice(s->device));
    RankBounds rb;
    rb.n = nranks;
    rb.self = self;
    for (int k = 0; k <= nranks; ++k) rb.lo[k] = (int)row_bounds[k];
    // scratch: counts[MAX_RANKS] | cursor[MAX_RANKS] | nleavers | ntargets | nsources
    uint32_t *scr = s->mscratch;
    uint32_t *nlist_d = scr + 2 * MAX_RANKS;
    FSIM_CUDA(cudaMemsetAsync(scr, 0, sizeof(uint32_t) * 2 * MAX_RANKS, s->stream));
    FSIM_CUDA(cudaMemsetAsync(scr + 2 * MAX_RANKS + 1, 0, sizeof(uint32_t) * 2, s->stream));
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
                (const Real *)s->part[c][AZ], s->n, s->nz, s->own0, s->own_rows, s->perm, nlist_d);
            FSIM_CUDA(cudaGetLastError());
            s->launches++;
        }
        s->have_leavers = false;
        // destination counts of the leavers and the list length: ONE read-back per frame
        migrate_count_kernel<Real><<<s->nsm * 2, 256, 0, s->stream>>>((const Real *)s->part[c][AZ], s->perm, nlist_d, s->nz, rb,
                                                                   scr);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        uint32_t nlist = 0;
        uint32_t hc[MAX_RANKS] = {}, off[MAX_RANKS] = {};
        FSIM_CUDA(cudaMemcpyAsync(&nlist, nlist_d, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        FSIM_CUDA(cudaMemcpyAsync(hc, scr, sizeof(uint32_t) * nranks, cudaMemcpyDeviceToHost, s->stream));
        FSIM_CUDA(cudaStreamSynchronize(s->stream));
        if (nlist == 0) return (int)FSIM_OK;
        if ((int64_t)nlist > s->cap / 2) {
            set_error("fsim_migrate_pack: more than half of the particle slots leave the slab at once");
            return (int)FSIM_ERR_RANGE;
        }
        int64_t total = 0;
        for (int k = 0; k < nranks; ++k) {
            send_counts[k] = hc[k];
            off[k] = (uint32_t)total;
            total += hc[k];
        }
        FSIM_TRY(ensure_migr(s, (size_t)total * (NPART_ARRAYS * sizeof(Real) + 8) + 16));
        FSIM_CUDA(cudaMemcpyAsync(scr + MAX_RANKS, off, sizeof(uint32_t) * nranks, cudaMemcpyHostToDevice, s->stream));
        PackArgs<Real> a;
        for (int k = 0; k < NPART_ARRAYS; ++k) a.src[k] = (const Real *)s->part[c][k];
        a.alive = s->alive[c];
        a.id = s->pid[c];
        a.buf = (unsigned char *)s->migr;
        a.cursor = scr + MAX_RANKS;
        a.list = s->perm; a.nlist = nlist;
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
    FSIM_CUDA(cudaSetDevice(s->device));
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
            a.holes = s->perm; a.hole_flag = s->hole_flag;
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
            // target/source lists: the tail of perm[] beyond the hole list (both are tiny)
            uint32_t *targets = s->perm + s->cap / 2, *sources = s->perm + s->cap / 2 + s->cap / 4;
            compact_lists_kernel<<<grid_for(span, 256), 256, 0, s->stream>>>(
                s->perm, (uint32_t)nholes, (uint32_t)nrecv, s->hole_flag, n_new, n_old, targets,
                scr + 2 * MAX_RANKS + 1, sources, scr + 2 * MAX_RANKS + 2);
            MoveArgs<Real> m;
            for (int k = 0; k < NPART_ARRAYS; ++k) m.a[k] = (Real *)s->part[c][k];
            m.alive = s->alive[c]; m.id = s->pid[c];
            m.targets = targets; m.sources = sources; m.ntargets = scr + 2 * MAX_RANKS + 1;
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

int fsim_halo_ptrs(fsim_sim *s, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, int64_t *bytes_each)
{
    if (!s || !send_lo || !send_hi || !recv_lo || !recv_hi || !bytes_each) {
        set_error("fsim_halo_ptrs: null argument");
        return FSIM_ERR_INVALID;
    }
    FSIM_CUDA(cudaSetDevice(s->device));
    const size_t each = (size_t)4 * FSIM_SHAPE_MID * s->nr * s->rs;
    if (!s->halo_buf) {
        FSIM_CUDA(cudaMalloc(&s->halo_buf, 4 * each));
        FSIM_CUDA(cudaMemset(s->halo_buf, 0, 4 * each));
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

}  // namespace fsim
