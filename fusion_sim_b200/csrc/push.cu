// push.cu -- the fused leap-frog half-step kernel (sm_100a).
//
// One launch = programStepRand + programStepVelocity + programStepPosition of the reference
// (empic.js:1438-1451 or :1454-1467) for every particle: per-particle RNG update, nearest-grid-
// point gather of the Boris rows R1..R3 and the half-kick A, rotation in cylindrical components,
// push, sink-mask absorption and inverse-cdf respawn.  The three shaders of a half-step all read
// the rand texture from BEFORE the update (bindings empic.js:819,832,849 / :894,907,924) and the
// position shader reads the old position with the new velocity (:847-848), so the fusion is
// exact and the update can be done in place (a particle touches only its own state).
//
// HBM-bound: 10 reals + 1 byte read and written per particle, streamed with ld.global.cs /
// st.global.cs (they keep the 126 MB L2 for the entropy and cell tables).  ONE particle per thread in both
// precisions: a warp's loads of consecutive particles are whole lines either way, and what the sweep needs is
// resident warps to cover its dependent gathers -- fp64 at 62 registers, 32 warps per SM (12 % faster than
// 128-bit loads = 2 particles per thread = 124 registers), fp32 at 40 registers, 48 warps (24 % faster than
// 128-bit loads = 4 particles per thread = 125 registers; ncu pair in profiles/r2_kernel_rooflines.md).
// Tables come through the read-only path.  Arithmetic is IEEE, left to right as the GLSL is
// written, compiled with -fmad=false so it rounds exactly like the CPU oracle.
#include <algorithm>
#include <map>

#include "common.cuh"

namespace fsim {

template <typename Real>
struct PushArgs {
    Real *a[NPART_ARRAYS];
    uint8_t *alive;
    // fused re-sort (PERM): slot j of the output takes the particle in slot perm[j] of `in` (the other copy)
    const Real *in[NPART_ARRAYS];
    const uint8_t *alive_in;
    const uint32_t *perm, *id_in;
    uint32_t *id_out;
    const Real *__restrict__ ent;
    const Real *__restrict__ cellrec;
#ifdef FSIM_TUNE
    const Real *__restrict__ cellrec12;  // measured alternative (OPT bit 2): nine Boris entries + A per cell, 12 reals
#endif
    const uint32_t *__restrict__ sink;  // 1 bit per GLOBAL cell (2 MB at 8192 x 2048: stays in L2)
    const Real *__restrict__ invcdf;
    uint32_t *key;      // optional deposit prepass: sort key, sprite colour, histogram
    Real *dcol[2];
    uint32_t *counts;
    uint32_t *oob;
    uint32_t *leavers, *nleavers;  // slab mode: slots whose new row is not owned (migration list)
    int own0, own_rows;
    int64_t n;              // live slots (an upper bound when n_dev is set)
    const uint32_t *n_dev;  // asynchronous slab exchange: the exact count lives on the device
    int nr, nz, row0, rows, own_lo, own_hi;
    int periodic;           // EXTENSION: z wraps (FSIM_FLAG_PERIODIC_Z)
    Real sf, h, k13, k31;
};

template <typename Real, int V> struct Vec;
template <> struct Vec<double, 2> { using T = double2; };
template <> struct Vec<float, 4> { using T = float4; };
template <> struct Vec<double, 1> { using T = double; };
template <> struct Vec<float, 1> { using T = float; };
template <> struct Vec<float, 2> { using T = float2; };

template <typename Real, int V>
__device__ __forceinline__ void ld_stream(const Real *p, Real (&o)[V])
{
    typename Vec<Real, V>::T t = __ldcs(reinterpret_cast<const typename Vec<Real, V>::T *>(p));
    const Real *q = reinterpret_cast<const Real *>(&t);
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = q[k];
}
template <typename Real, int V>
__device__ __forceinline__ void st_stream(Real *p, const Real (&o)[V])
{
    typename Vec<Real, V>::T t;
    Real *q = reinterpret_cast<Real *>(&t);
#pragma unroll
    for (int k = 0; k < V; ++k) q[k] = o[k];
    __stcs(reinterpret_cast<typename Vec<Real, V>::T *>(p), t);
}

// 4 reals of one entropy texel / 12 reals of one cell record through the read-only path
// 256-bit read-only load (sm_100: LDG.E.256): one request, and for a gather one L1 wavefront per
// lane, where two 128-bit loads take two.  p must be 32-byte aligned.
__device__ __forceinline__ void ld_ro256(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ld_ro4(const double *p, double (&o)[4])
{
    ld_ro256(p, o[0], o[1], o[2], o[3]);
}
// entropy texel: a random 32-byte gather from a 33.5 MB table never hits L1 again -- do not let it evict the
// cell records and sink words that do (L1::no_allocate)
__device__ __forceinline__ void ld_ent_na(const double *p, double (&o)[4])
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
}
__device__ __forceinline__ void ld_ent_na(const float *p, float (&o)[4])
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p));
}
__device__ __forceinline__ void ld_ro4(const float *p, float (&o)[4])
{
    float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}
// the 8-real cell record (common.cuh): two 256-bit loads in fp64, two 128-bit loads in fp32
__device__ __forceinline__ void ld_rec(const double *p, double (&o)[RECSTRIDE])
{
    ld_ro256(p, o[0], o[1], o[2], o[3]);
    ld_ro256(p + 4, o[4], o[5], o[6], o[7]);
}
__device__ __forceinline__ void ld_rec(const float *p, float (&o)[RECSTRIDE])
{
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(p) + k);
        o[4 * k] = a.x; o[4 * k + 1] = a.y; o[4 * k + 2] = a.z; o[4 * k + 3] = a.w;
    }
}
#ifdef FSIM_TUNE
__device__ __forceinline__ void ld_rec12(const double *p, double (&o)[12])
{
    ld_ro256(p, o[0], o[1], o[2], o[3]);
    ld_ro256(p + 4, o[4], o[5], o[6], o[7]);
    ld_ro256(p + 8, o[8], o[9], o[10], o[11]);
}
__device__ __forceinline__ void ld_rec12(const float *p, float (&o)[12])
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(p) + k);
        o[4 * k] = a.x; o[4 * k + 1] = a.y; o[4 * k + 2] = a.z; o[4 * k + 3] = a.w;
    }
}
#endif
__device__ __forceinline__ void ld_ro2(const double *p, double &a, double &b)
{
    double2 t = __ldg(reinterpret_cast<const double2 *>(p));
    a = t.x; b = t.y;
}
__device__ __forceinline__ void ld_ro2(const float *p, float &a, float &b)
{
    float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    a = t.x; b = t.y;
}

// The particle state a thread carries: V consecutive slots of every array.
template <typename Real, int V>
struct Slots {
    Real x[V], y[V], z[V], vx[V], vy[V], vz[V], q0[V], q1[V], q2[V], q3[V];
    Real rcur[V];  // sqrt(x*x + y*y) of the current position (:755; reused by the next half-step)
    uint8_t al[V];
};

// NH leap-frog half-steps of the V particles in `t`, entirely in registers.
// out.step() is two half-steps with nothing in between that couples particles (static fields), so
// NH = 2 performs the B-pass and the A-pass of empic.js:1438-1467 in one sweep over HBM: state
// read once, written once.  Same operations in the same order, hence the same bits.
// OPT bit 0: entropy gathers bypass L1 allocation; bit 1: the entropy texel of the NEXT half-step is fetched as soon
// as the new RNG state exists (it depends on nothing else), under the Boris arithmetic of this one.
template <typename Real, int V, int NH, int OPT = 0>
__device__ __forceinline__ void advance(const PushArgs<Real> &a, const int64_t p0, const int64_t n, Slots<Real, V> &t)
{
    Real (&x)[V] = t.x, (&y)[V] = t.y, (&z)[V] = t.z, (&vx)[V] = t.vx, (&vy)[V] = t.vy, (&vz)[V] = t.vz;
    Real (&q0)[V] = t.q0, (&q1)[V] = t.q1, (&q2)[V] = t.q2, (&q3)[V] = t.q3, (&rcur)[V] = t.rcur;
    uint8_t (&al)[V] = t.al;
#pragma unroll
    for (int k = 0; k < V; ++k) rcur[k] = fsqrt(x[k] * x[k] + y[k] * y[k]);

    auto gather_entropy = [&](Real (&e)[V][4]) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int ie = tex_idx(q2[k], FSIM_N_ENTROPY) + FSIM_N_ENTROPY * tex_idx(q3[k], FSIM_N_ENTROPY);
            if constexpr (OPT & 1) ld_ent_na(a.ent + 4 * (size_t)ie, e[k]);
            else ld_ro4(a.ent + 4 * (size_t)ie, e[k]);
        }
    };
    Real e[V][4];
    if constexpr (OPT & 2) gather_entropy(e);

#pragma unroll 1
    for (int hs = 0; hs < NH; ++hs) {
        // dependent gathers: entropy texel (empic.js:802) and cell record (:763-766)
        constexpr int RW = (OPT & 4) ? 12 : RECSTRIDE;
        Real rec[V][RW], dx[V], dy[V];
        if constexpr (!(OPT & 2)) gather_entropy(e);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const Real r = rcur[k];
            dx[k] = x[k] / r;  // :756
            dy[k] = y[k] / r;
            const int ci = tex_idx(r, a.nr);
            int cj = tex_idx(z[k], a.nz) - a.row0;
            if (cj < 0 || cj >= a.rows) {  // slab mode: particle outside the local table
                // counted only if the record is USED: a particle respawned in the previous half-step
                // (al = 0) gets a fresh random velocity (:772) and may sit anywhere in the global domain
                if (al[k] && p0 + k < n) atomicAdd(a.oob, 1u);
                cj = cj < 0 ? 0 : a.rows - 1;
            }
#ifdef FSIM_TUNE
            if constexpr ((OPT & 4) != 0) ld_rec12(a.cellrec12 + 12 * ((size_t)ci + (size_t)cj * a.nr), rec[k]);
            else
#endif
            ld_rec(a.cellrec + RECSTRIDE * ((size_t)ci + (size_t)cj * a.nr), rec[k]);
        }

#pragma unroll
        for (int k = 0; k < V; ++k) {
            // ---- programStepRandA/B, empic.js:800-807 ----
            const Real x0 = (Real)FSIM_RNG_KEEP * q2[k] + (Real)FSIM_RNG_MIX * e[k][2];
            const Real x1 = (Real)FSIM_RNG_KEEP * q3[k] + (Real)FSIM_RNG_MIX * e[k][3];
            const Real m0 = q0[k] + e[k][0];
            const Real m1 = q1[k] + e[k][1];
            const Real o0 = q0[k], o1 = q1[k], o2 = q2[k];  // shaders below read the OLD rand
            q0[k] = (m0 > (Real)1.0) ? m0 - (Real)1.0 : m0;
            q1[k] = (m1 > (Real)1.0) ? m1 - (Real)1.0 : m1;
            q2[k] = (Real)4.0 * x0 * ((Real)1.0 - x0);
            q3[k] = (Real)4.0 * x1 * ((Real)1.0 - x1);
            if constexpr ((OPT & 2) != 0) {  // next half-step's texel, as early as its address exists
                if (hs + 1 < NH) {
                    const int ie = tex_idx(q2[k], FSIM_N_ENTROPY) + FSIM_N_ENTROPY * tex_idx(q3[k], FSIM_N_ENTROPY);
                    if constexpr (OPT & 1) ld_ent_na(a.ent + 4 * (size_t)ie, e[k]);
                    else ld_ro4(a.ent + 4 * (size_t)ie, e[k]);
                }
            }

            // ---- step_velocity_frag, empic.js:758-772 ----
            const Real vr = vx[k] * dx[k] + vy[k] * dy[k];
            const Real va = vy[k] * dx[k] - vx[k] * dy[k];
            Real R[9];  // rows of the Boris matrix, rebuilt from the record (programPre1/2/3)
            Real Ax, Ay, Az;
            if constexpr ((OPT & 4) != 0) {
#pragma unroll
                for (int i = 0; i < 9; ++i) R[i] = rec[k][i];
                Ax = rec[k][9]; Ay = rec[k][10]; Az = rec[k][11];
            } else {
                boris_rows<Real>(rec[k], a.h, a.k13, a.k31, R);
                Ax = rec[k][REC_AX]; Ay = rec[k][REC_AY]; Az = rec[k][REC_AZ];
            }
            const Real c0 = (R[0] * vr + R[1] * va + R[2] * vz[k]) + Ax;
            const Real c1 = (R[3] * vr + R[4] * va + R[5] * vz[k]) + Ay;
            const Real c2 = (R[6] * vr + R[7] * va + R[8] * vz[k]) + Az;
            Real nvx, nvy, nvz;
            if (al[k]) {
                nvx = c0 * dx[k] - c1 * dy[k];
                nvy = c0 * dy[k] + c1 * dx[k];
                nvz = c2;
            } else {  // just respawned: fresh random velocity (:772)
                nvx = (Real)FSIM_RESPAWN_SPEED * ((Real)2.0 * o0 - (Real)1.0);
                nvy = (Real)FSIM_RESPAWN_SPEED * ((Real)2.0 * o1 - (Real)1.0);
                nvz = (Real)FSIM_RESPAWN_SPEED * ((Real)2.0 * o2 - (Real)1.0);
            }
            vx[k] = nvx; vy[k] = nvy; vz[k] = nvz;

            // ---- step_position_frag, empic.js:714-719 ----
            const Real nx = x[k] + a.sf * nvx;
            const Real ny = y[k] + a.sf * nvy;
            Real nzp = z[k] + a.sf * nvz;
            if (a.periodic) {  // EXTENSION (no reference counterpart): periodic in z
                nzp = nzp - ffloor(nzp);
                if (nzp >= (Real)1.0) nzp = (Real)0.0;
            }
            const Real rn = fsqrt(nx * nx + ny * ny);
            bool keep = false;
            if (rn == rn && nzp == nzp) {  // NaN position => absorbed (documented rule)
                const uint32_t gc = (uint32_t)tex_idx(rn, a.nr) + (uint32_t)tex_idx(nzp, a.nz) * (uint32_t)a.nr;
                keep = (__ldg(a.sink + (gc >> 5)) >> (gc & 31u)) & 1u;
            }
            if (keep) {
                x[k] = nx; y[k] = ny; z[k] = nzp; al[k] = 1;
                rcur[k] = rn;
            } else {
                const int it = tex_idx(o0, FSIM_N_INVCDF) + FSIM_N_INVCDF * tex_idx(o1, FSIM_N_INVCDF);
                Real sx, sz;
                ld_ro2(a.invcdf + 2 * (size_t)it, sx, sz);
                x[k] = sx; y[k] = (Real)0.0; z[k] = sz; al[k] = 0;
                rcur[k] = fsqrt(sx * sx + (Real)0.0 * (Real)0.0);
            }
        }
    }

}

// Deposit prepass on the NEW state (what density() will see): sort key, sprite colour and the
// warp-aggregated histogram of the counting sort.  Whole warps must call this converged.
template <typename Real, int V>
__device__ __forceinline__ void emit_prepass(const PushArgs<Real> &a, const int64_t p0, const int64_t n, const Slots<Real, V> &t)
{
    const int lane = threadIdx.x & 31;
    uint32_t newcell[V];
    Real col[3][V];
#pragma unroll
    for (int k = 0; k < V; ++k)
        newcell[k] = sprite_key_colour<Real>(t.x[k], t.y[k], t.z[k], t.rcur[k], t.vx[k], t.vy[k], t.vz[k], a.nr, a.nz,
                                             a.row0, a.rows, a.own_lo, a.own_hi, col[0][k], col[1][k], col[2][k]);
#pragma unroll
    for (int q = 0; q < 2; ++q) st_stream<Real, V>(a.dcol[q] + p0, col[q]);  // 0.001 v_z is formed by the per-cell pass from v_z itself
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const bool valid = p0 + k < n;
        const uint32_t c = valid ? (newcell[k] & KEY_MASK) : 0xffffffffu;
        if (valid) a.key[p0 + k] = newcell[k];
        if (valid && a.leavers) {
            const int gj = tex_idx(t.z[k], a.nz);
            if (gj < a.own0 || gj >= a.own0 + a.own_rows) {
                const uint32_t at = atomicAdd(a.nleavers, 1u);
                FSIM_ASSERT((int64_t)at < n);
                a.leavers[at] = (uint32_t)(p0 + k);
            }
        }
        int leader;
        uint32_t len, rank;
        warp_runs(c, lane, leader, len, rank);
        if (valid && rank == 0) atomicAdd(a.counts + c, len);
    }
}

template <typename Real, int V>
__device__ __forceinline__ void store_slots(const PushArgs<Real> &a, const int64_t p0, const Slots<Real, V> &t)
{
    st_stream<Real, V>(a.a[AX] + p0, t.x);
    st_stream<Real, V>(a.a[AY] + p0, t.y);
    st_stream<Real, V>(a.a[AZ] + p0, t.z);
    st_stream<Real, V>(a.a[AVX] + p0, t.vx);
    st_stream<Real, V>(a.a[AVY] + p0, t.vy);
    st_stream<Real, V>(a.a[AVZ] + p0, t.vz);
    st_stream<Real, V>(a.a[AQ0] + p0, t.q0);
    st_stream<Real, V>(a.a[AQ1] + p0, t.q1);
    st_stream<Real, V>(a.a[AQ2] + p0, t.q2);
    st_stream<Real, V>(a.a[AQ3] + p0, t.q3);
#pragma unroll
    for (int k = 0; k < V; ++k) a.alive[p0 + k] = t.al[k];
}

// One thread = V consecutive particles, state loaded straight into registers.  (A persistent
// variant that prefetched the next tile's state into shared memory with cp.async -- no registers
// held while the streaming loads fly -- was measured 18 % SLOWER on B200: the sweep is not bound by
// the latency of the streaming loads but by L1TEX wavefronts and dependent fp64 chains, and the
// detour through shared memory adds to both.  DESIGN.md section 4.)
template <typename Real, int V, int BLOCK, int MINB, int NH, bool PERM, int OPT = 0>
__global__ void __launch_bounds__(BLOCK, MINB) push_kernel(const PushArgs<Real> a)
{
    // no early exit: the arrays are padded past n (common.cuh), the whole warp stays converged for
    // the warp-aggregated histogram; side effects of slots >= n are masked.
    const int64_t p0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    const int64_t n = live_count(a.n_dev, a.n);
    Slots<Real, V> t;
    if constexpr (PERM) {
        // Fused physical re-sort: the sweep READS through the index list of the last binning (gathered, but
        // the storage is nearly ordered, so neighbouring lanes stay within a few lines) and WRITES the
        // other copy in cell order, coalesced -- the storage is re-sorted without a pass of its own.
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int64_t j = p0 + k;
            const size_t p = j < n ? (size_t)(a.perm[j] & KEY_MASK) : (size_t)j;
            FSIM_ASSERT(j >= n || (int64_t)p < n);
            t.q2[k] = __ldcs(a.in[AQ2] + p); t.q3[k] = __ldcs(a.in[AQ3] + p);
            t.x[k] = __ldcs(a.in[AX] + p); t.y[k] = __ldcs(a.in[AY] + p); t.z[k] = __ldcs(a.in[AZ] + p);
            t.q0[k] = __ldcs(a.in[AQ0] + p); t.q1[k] = __ldcs(a.in[AQ1] + p);
            t.vx[k] = __ldcs(a.in[AVX] + p); t.vy[k] = __ldcs(a.in[AVY] + p); t.vz[k] = __ldcs(a.in[AVZ] + p);
            t.al[k] = a.alive_in[p];
            a.id_out[j] = a.id_in[p];
        }
    } else {
        ld_stream<Real, V>(a.a[AQ2] + p0, t.q2);
        ld_stream<Real, V>(a.a[AQ3] + p0, t.q3);
        ld_stream<Real, V>(a.a[AX] + p0, t.x);
        ld_stream<Real, V>(a.a[AY] + p0, t.y);
        ld_stream<Real, V>(a.a[AZ] + p0, t.z);
        ld_stream<Real, V>(a.a[AQ0] + p0, t.q0);
        ld_stream<Real, V>(a.a[AQ1] + p0, t.q1);
        ld_stream<Real, V>(a.a[AVX] + p0, t.vx);
        ld_stream<Real, V>(a.a[AVY] + p0, t.vy);
        ld_stream<Real, V>(a.a[AVZ] + p0, t.vz);
#pragma unroll
        for (int k = 0; k < V; ++k) t.al[k] = a.alive[p0 + k];
    }
    advance<Real, V, NH, OPT>(a, p0, n, t);
    store_slots<Real, V>(a, p0, t);
    if (a.key) emit_prepass<Real, V>(a, p0, n, t);
}


#ifdef FSIM_TUNE
// ---- MEASURED ALTERNATIVE (tuning build only): the same sweep with the particle state staged by the TMA unit ----
// Persistent blocks, one tile of 256 consecutive particles at a time.  The ten state arrays (and the alive
// bytes) of a tile are brought into shared memory by eleven 1-D bulk copies (cp.async.bulk, SASS UBLKCP,
// completion counted in bytes on an mbarrier), STAGES tiles ahead of the arithmetic.  The idea: this access
// pattern by itself streams at 6.86 TB/s (tools/stream_layout_bench.cu) while the per-thread-load sweep moves
// its bytes at 5.1 TB/s, so keep more bytes in flight.  The measurement (profiles/r2_push_variants.md): SLOWER,
// 3.31 ms (2 stages x 3 blocks per SM) to 4.05 ms (4 stages) against 2.94 ms -- the more shared memory the
// stages take, the slower: shared memory and L1 are one array on the SM, and the sweep lives on the L1 hits of
// its cell-record and sink gathers.  Not in the product (and its multi-tile path has an unresolved
// data hazard at > 3 tiles per block; it is kept for the measurement only).
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        if (spin > (1u << 26)) __trap();  // a copy that never lands must not hang the GPU
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

template <typename Real, int NH, int STAGES, int MINB>
__global__ void __launch_bounds__(256, MINB) push_tma_kernel(const PushArgs<Real> a, const int64_t ntiles)
{
    constexpr int T = 256;
    constexpr uint32_t ARR_BYTES = T * sizeof(Real);
    constexpr uint32_t STAGE_BYTES = NPART_ARRAYS * ARR_BYTES + T;  // + the alive bytes
    extern __shared__ __align__(128) unsigned char tile_smem[];
    __shared__ __align__(8) unsigned long long full[STAGES];
    const int tid = threadIdx.x;
    const int64_t n = live_count(a.n_dev, a.n);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < STAGES; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full[k])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long policy = 0;  // streamed once: first out of L2, the tables stay
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    auto issue = [&](int stage, int64_t tile) {  // one thread: arm the barrier, launch the eleven copies of `tile`
        const uint32_t bar = smem_addr(&full[stage]);
        unsigned char *dst = tile_smem + (size_t)stage * STAGE_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE_BYTES) : "memory");
#pragma unroll
        for (int k = 0; k < NPART_ARRAYS; ++k)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_addr(dst + k * ARR_BYTES)), "l"(a.a[k] + tile * T), "r"(ARR_BYTES), "r"(bar), "l"(policy) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(smem_addr(dst + NPART_ARRAYS * ARR_BYTES)), "l"(a.alive + tile * T), "r"((uint32_t)T), "r"(bar), "l"(policy) : "memory");
    };
    const int64_t first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < STAGES; ++k)
            if (first + k * stride < ntiles) issue(k, first + k * stride);
    }
    int it = 0;
    for (int64_t tile = first; tile < ntiles; tile += stride, ++it) {
        const int stage = it % STAGES;
        mbar_wait(smem_addr(&full[stage]), (uint32_t)(it / STAGES) & 1u);
        const Real *sm = reinterpret_cast<const Real *>(tile_smem + (size_t)stage * STAGE_BYTES);
        Slots<Real, 1> t;
        t.x[0] = sm[AX * T + tid]; t.y[0] = sm[AY * T + tid]; t.z[0] = sm[AZ * T + tid];
        t.vx[0] = sm[AVX * T + tid]; t.vy[0] = sm[AVY * T + tid]; t.vz[0] = sm[AVZ * T + tid];
        t.q0[0] = sm[AQ0 * T + tid]; t.q1[0] = sm[AQ1 * T + tid]; t.q2[0] = sm[AQ2 * T + tid]; t.q3[0] = sm[AQ3 * T + tid];
        t.al[0] = tile_smem[(size_t)stage * STAGE_BYTES + NPART_ARRAYS * ARR_BYTES + tid];
        __syncthreads();  // every thread holds its particle in registers: the stage is free for the tile STAGES ahead
        if (tid == 0 && tile + STAGES * stride < ntiles) issue(stage, tile + STAGES * stride);
        const int64_t p0 = tile * T + tid;
        advance<Real, 1, NH>(a, p0, n, t);
        store_slots<Real, 1>(a, p0, t);
        if (a.key) emit_prepass<Real, 1>(a, p0, n, t);
    }
}

#endif  // FSIM_TUNE

template <typename Real>
static PushArgs<Real> make_args(fsim_sim *s, bool with_hist, bool resort)
{
    PushArgs<Real> a;
    const int out = resort ? s->cur ^ 1 : s->cur;  // fused re-sort: read the current copy through perm[], write the other one
    for (int k = 0; k < NPART_ARRAYS; ++k) {
        a.a[k] = (Real *)s->part[out][k];
        a.in[k] = (const Real *)s->part[s->cur][k];
    }
    a.alive = s->alive[out];
    a.alive_in = s->alive[s->cur];
    a.perm = s->perm;
    a.id_in = s->pid[s->cur];
    a.id_out = s->pid[out];
    a.ent = (const Real *)s->entropy;
    a.cellrec = (const Real *)s->cellrec;
    a.sink = s->sink;
    a.invcdf = (const Real *)s->invcdf;
    a.key = with_hist ? s->key : nullptr;
    for (int q = 0; q < 2; ++q) a.dcol[q] = (Real *)s->dcol[q];
    a.counts = s->counts;
    a.oob = s->oob;
    a.leavers = (with_hist && s->slab) ? s->leavers : nullptr;
    a.nleavers = s->mscratch + MC_NLEAVERS;
    a.own0 = s->own0; a.own_rows = s->own_rows;
    a.n = s->n;
    a.n_dev = s->n_async ? s->mscratch + MC_NLIVE : nullptr;
    a.nr = s->nr; a.nz = s->nz; a.row0 = s->row0; a.rows = s->rows;
    a.own_lo = s->own0 - s->row0; a.own_hi = a.own_lo + s->own_rows;
    a.periodic = s->ring ? 1 : 0;
    a.sf = (Real)s->step_factor;
    a.h = (Real)s->h; a.k13 = (Real)s->k13; a.k31 = (Real)s->k31;
    return a;
}

template <typename Real, int V, int BLOCK, int MINB, int OPT = 0>
static int push_impl(fsim_sim *s, const PushArgs<Real> &a, int nhalf, bool resort = false)
{
    const int64_t nvec = (s->n + V - 1) / V;
    if (resort)  // only step() re-sorts: two half-steps
        push_kernel<Real, V, BLOCK, MINB, 2, true, OPT><<<grid_for(nvec, BLOCK), BLOCK, 0, s->stream>>>(a);
    else if (nhalf == 2)
        push_kernel<Real, V, BLOCK, MINB, 2, false, OPT><<<grid_for(nvec, BLOCK), BLOCK, 0, s->stream>>>(a);
    else  // the single half-step keeps more values live per particle: it spills under a cap below 64 registers
        push_kernel<Real, V, BLOCK, (MINB * BLOCK > 1024 ? 1024 / BLOCK : MINB), 1, false><<<grid_for(nvec, BLOCK), BLOCK, 0, s->stream>>>(a);
    FSIM_CUDA(cudaGetLastError());
    return FSIM_OK;
}

#ifdef FSIM_TUNE
// The TMA-staged sweep: persistent grid, STAGES tiles in flight per block, MINB blocks per SM.
template <typename Real, int STAGES, int MINB>
static int push_tma_impl(fsim_sim *s, const PushArgs<Real> &a)
{
    constexpr size_t smem = (size_t)STAGES * (NPART_ARRAYS * 256 * sizeof(Real) + 256);
    constexpr uint32_t bit = 1u << (sizeof(Real) == 8 ? 30 : 31);  // bits 30, 31 of smem_opt_in
#ifdef FSIM_TUNE
    s->smem_opt_in &= ~bit;  // several instantiations share the bit in the tuning build
#endif
    if (!(s->smem_opt_in & bit)) {
        FSIM_CUDA(cudaFuncSetAttribute(push_tma_kernel<Real, 2, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s->smem_opt_in |= bit;
    }
    const int64_t ntiles = (s->n + 255) / 256;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)s->nsm * MINB);
    push_tma_kernel<Real, 2, STAGES, MINB><<<grid, 256, smem, s->stream>>>(a, ntiles);
    FSIM_CUDA(cudaGetLastError());
    return FSIM_OK;
}

// Tuning build only (make EXTRA=-DFSIM_TUNE, tools/tune.py): vector width / block size / register cap
// variants selectable at run time through fsim_tune_set().  The product compiles ONE variant per precision.
int g_push_variant = 0;

// measured alternative (variants 40-42): a 12-real cell table (nine Boris entries + A) expanded from the records
template <typename Real>
__global__ void __launch_bounds__(256) expand12_kernel(const Real *__restrict__ rec, Real *__restrict__ out, int64_t ncell, Real h, Real k13, Real k31)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    Real R[9];
    boris_rows<Real>(rec + RECSTRIDE * c, h, k13, k31, R);
#pragma unroll
    for (int i = 0; i < 9; ++i) out[12 * c + i] = R[i];
    out[12 * c + 9] = rec[RECSTRIDE * c + REC_AX];
    out[12 * c + 10] = rec[RECSTRIDE * c + REC_AY];
    out[12 * c + 11] = rec[RECSTRIDE * c + REC_AZ];
}
static std::map<const fsim_sim *, void *> g_rec12;
template <typename Real>
static const Real *rec12_of(fsim_sim *s)
{
    auto it = g_rec12.find(s);
    if (it != g_rec12.end()) return (const Real *)it->second;
    void *p = nullptr;
    if (cudaMalloc(&p, sizeof(Real) * 12 * (size_t)s->ncell_local) != cudaSuccess) return nullptr;
    expand12_kernel<Real><<<grid_for(s->ncell_local, 256), 256, 0, s->stream>>>((const Real *)s->cellrec, (Real *)p, s->ncell_local,
                                                                               (Real)s->h, (Real)s->k13, (Real)s->k31);
    g_rec12[s] = p;
    return (const Real *)p;
}
#endif

// resort: the sweep also performs the physical re-sort the last binning prepared (perm[] must match the
// current positions: s->binned); the particle storage flips to the other copy.
int launch_push(fsim_sim *s, bool with_hist, int nhalf, bool resort)
{
    if (resort && !(s->binned && nhalf == 2 && with_hist)) resort = false;
    if (with_hist && s->counts_dirty) {  // an unconsumed histogram: start from zero
        FSIM_CUDA(cudaMemsetAsync(s->counts, 0, sizeof(uint32_t) * (s->ncell_local + 1), s->stream));
        s->counts_dirty = false;
    }
    if (with_hist && s->slab) FSIM_CUDA(cudaMemsetAsync(s->mscratch + MC_NLEAVERS, 0, sizeof(uint32_t), s->stream));
    int rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        if (s->n == 0) return (int)FSIM_OK;
        const PushArgs<Real> a = make_args<Real>(s, with_hist, resort);
        Bracket b(s, resort ? "push2_resort" : (nhalf == 2 ? "push2" : "push"));
#ifdef FSIM_TUNE
        constexpr int V = 16 / sizeof(Real);  // particles per 128 bits
        switch (g_push_variant) {
        case 1: return push_impl<Real, V, 128, 4>(s, a, nhalf, resort);
        case 2: return push_impl<Real, V, 256, 3>(s, a, nhalf, resort);
        case 3: return push_impl<Real, V / 2, 256, 3>(s, a, nhalf, resort);
        case 4: return push_impl<Real, V / 2, 256, 4>(s, a, nhalf, resort);
        case 5: return push_impl<Real, V / 2, 128, 6>(s, a, nhalf, resort);
        case 6: return push_impl<Real, V / 2, 512, 2>(s, a, nhalf, resort);
        case 7: return push_impl<Real, V, 256, 2>(s, a, nhalf, resort);
        case 10: if (nhalf == 2 && !resort) return push_tma_impl<Real, 2, 3>(s, a); break;
        case 11: if (nhalf == 2 && !resort) return push_tma_impl<Real, 2, 4>(s, a); break;
        case 12: if (nhalf == 2 && !resort) return push_tma_impl<Real, 3, 3>(s, a); break;
        case 13: if (nhalf == 2 && !resort) return push_tma_impl<Real, 4, 2>(s, a); break;
        case 14: if (nhalf == 2 && !resort) return push_tma_impl<Real, 3, 2>(s, a); break;
        case 15: if (nhalf == 2 && !resort) return push_tma_impl<Real, 2, 5>(s, a); break;
        case 20: return push_impl<Real, V / 2, 256, 4, 1>(s, a, nhalf, resort);
        case 21: return push_impl<Real, V / 2, 256, 4, 2>(s, a, nhalf, resort);
        case 22: return push_impl<Real, V / 2, 256, 4, 3>(s, a, nhalf, resort);
        case 23: return push_impl<Real, V / 2, 256, 3, 3>(s, a, nhalf, resort);
        case 24: return push_impl<Real, V, 256, 2, 3>(s, a, nhalf, resort);
        // more resident warps at a tighter register cap (one particle per thread; fp32: also two)
        case 30: return push_impl<Real, 1, 256, 5>(s, a, nhalf, resort);
        case 31: return push_impl<Real, 1, 256, 6>(s, a, nhalf, resort);
        case 32: return push_impl<Real, 1, 256, 8>(s, a, nhalf, resort);
        case 33: return push_impl<Real, 1, 256, 4>(s, a, nhalf, resort);
        case 34: return push_impl<Real, V / 2, 256, 5>(s, a, nhalf, resort);
        case 35: return push_impl<Real, V / 2, 256, 6>(s, a, nhalf, resort);
        case 36: return push_impl<Real, 1, 256, 6, 2>(s, a, nhalf, resort);
        case 37: return push_impl<Real, 1, 512, 3>(s, a, nhalf, resort);
        case 38: return push_impl<Real, 1, 128, 12>(s, a, nhalf, resort);
        case 39: return push_impl<Real, 1, 256, 7>(s, a, nhalf, resort);
        case 40: { PushArgs<Real> b2 = a; b2.cellrec12 = rec12_of<Real>(s); return push_impl<Real, 1, 256, (sizeof(Real) == 4 ? 6 : 4), 4>(s, b2, nhalf, resort); }
        case 41: { PushArgs<Real> b2 = a; b2.cellrec12 = rec12_of<Real>(s); return push_impl<Real, 1, 256, (sizeof(Real) == 4 ? 5 : 3), 4>(s, b2, nhalf, resort); }
        case 42: { PushArgs<Real> b2 = a; b2.cellrec12 = rec12_of<Real>(s); return push_impl<Real, 1, 256, (sizeof(Real) == 4 ? 8 : 4), 4>(s, b2, nhalf, resort); }
        default: break;
        }
#endif
        // measured fastest on B200 (profiles/r2_push_variants.md): one particle per thread; fp64 256 threads x 4 blocks
        // per SM (64 registers, 32 warps), fp32 x 6 blocks (40 registers, 48 warps, no spills): more particles in
        // flight under the latency of the dependent gathers (fp64 spills below 62 registers and loses 30 %)
        return push_impl<Real, 1, 256, (sizeof(Real) == 4 ? 6 : 4)>(s, a, nhalf, resort);
    });
    if (resort && rc == FSIM_OK && s->n) {
        s->cur ^= 1;
        s->ever_sorted = true;
        s->ids_identity = false;
        s->steps_since_sort = 0;
        s->resort_due = false;
    }
    s->binned = false;
    s->keys_valid = with_hist;
    s->have_leavers = with_hist && s->slab;
    if (with_hist) s->counts_dirty = true;
    return rc;
}

}  // namespace fsim
