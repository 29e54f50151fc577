// common.cuh -- shared state and helpers of libfusionsim.so (sm_100a only).
//
// The engine behind the C ABI of include/fusionsim.h.  Data layout in HBM:
//   particles : structure of arrays, 10 reals + 1 alive byte + 1 u32 id per particle,
//               two copies (the counting sort is out of place);
//   cell table: array of 8-real records B.xyz f one_m A.xyz (below), index i + j*nr
//               (empic.js:1162), rows [row0, row0+rows) of the global grid;
//   sink mask : 1 BIT per GLOBAL cell; entropy 1024^2 x 4 reals; inv_cdf 512^2 x 2 reals.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/fsim_constants.h"
#include "../../include/fusionsim.h"

// Debug build (make EXTRA=-DFSIM_DEBUG_BOUNDS; tools/ab_build.sh dbg WORK -DFSIM_DEBUG_BOUNDS): every index that
// is READ FROM MEMORY before it addresses the particle storage or a list (perm[], holes, leavers, targets,
// sources, cursors) is checked on the device; a violation traps the kernel and the next API call fails with
// FSIM_ERR_CUDA.  The pool forbids compute-sanitizer, this is its substitute: the GPU test-suite is run once
// per round against this build (profiles/).  Compiled out of the product.
#ifdef FSIM_DEBUG_BOUNDS
#include <assert.h>
#define FSIM_ASSERT(cond) assert(cond)
#else
#define FSIM_ASSERT(cond) ((void)0)
#endif

namespace fsim {

constexpr int NPART_ARRAYS = 10;  // x y z vx vy vz q0 q1 q2 q3
// Device cell record: 8 reals = 64 bytes in fp64 (two 256-bit loads, half a cache line):
//   Bx By Bz f one_m Ax Ay Az,   f = 2/(1+h^2|B|^2), one_m = 1 - h^2|B|^2 f   (empic.js:521-524)
// The nine Boris entries R1..R3 of programPre1/2/3 are products of these (no sqrt, no divide);
// the push rebuilds them per particle with the reference's own expressions (boris_rows below) --
// ~33 cheap fp64 operations instead of gathering 96 bytes: the gather, not the fp64 pipe, is what
// limits the step kernel on B200.  Same expressions, same order => same bits as a stored table.
constexpr int RECSTRIDE = 8;

// ---- multi-GPU slab exchange (migrate.cu) ---------------------------------------------------------
constexpr int MAX_RANKS = 64;
// words of the small device counter block fsim_sim::mscratch
enum {
    MC_NLEAVERS = 0,  // length of the leaver list the push (or find_leavers) wrote into perm[]
    MC_NHOLES,        // slots vacated by packed leavers (list in MigratePlan::holes / perm[])
    MC_NTARGETS, MC_NSOURCES,  // compaction lists
    MC_NRECV,         // arrivals of this frame (sum of the received region headers)
    MC_NOLD, MC_NNEW, // live slots before / after this frame's migration
    MC_ERR,           // sticky error bits MERR_* (read by fsim_sync)
    MC_SENT_LO, MC_SENT_HI,  // 64-bit running total of packed records (statistics)
    MC_NLIVE,         // device-resident particle count (asynchronous exchange: fsim_sim::n is an upper bound)
    MC_CURSOR = 16,                        // [MAX_RANKS] records packed for each destination
    MC_PREFIX = MC_CURSOR + MAX_RANKS,     // [MAX_RANKS + 1] exclusive prefix of the received counts
    MC_COUNTS = MC_PREFIX + MAX_RANKS + 1, // [MAX_RANKS] destination counts (exact exchange)
    MC_WORDS = MC_COUNTS + MAX_RANKS
};
constexpr uint32_t MERR_SEND_OVERFLOW = 1u;  // more leavers for one destination than its send region holds
constexpr uint32_t MERR_CAPACITY = 2u;       // arrivals exceed the particle capacity of this rank
constexpr int MIGRATE_HEADER_BYTES = 16;     // every exchange region starts with its record count (u32) + padding

// Fixed-capacity exchange plan (fsim_migrate_setup): region k of the send buffer goes to rank k, region k
// of the receive buffer comes from rank k; sizes are known to the host of both sides, the record counts
// travel in the region headers -- no host round trip in the frame.
struct MigratePlan {
    int nranks = 0, self = 0;
    int lo[MAX_RANKS + 1] = {};            // rank k owns rows [lo[k], lo[k+1])
    uint32_t send_cap[MAX_RANKS] = {}, recv_cap[MAX_RANKS] = {};  // records
    uint64_t send_off[MAX_RANKS + 1] = {}, recv_off[MAX_RANKS + 1] = {};  // byte offsets of the regions
    uint32_t recv_slot0[MAX_RANKS + 1] = {};  // exclusive prefix of recv_cap (thread -> region mapping)
    uint32_t send_total = 0, recv_total = 0;  // sums of the capacities
    unsigned char *send = nullptr, *recv = nullptr;
    uint32_t *holes = nullptr, *targets = nullptr, *sources = nullptr;  // [send_total] each
};
enum { REC_BX = 0, REC_BY, REC_BZ, REC_F, REC_ONEM, REC_AX, REC_AY, REC_AZ };
enum { AX = 0, AY, AZ, AVX, AVY, AVZ, AQ0, AQ1, AQ2, AQ3 };

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FSIM_CUDA(call)                                                         \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return fsim::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define FSIM_TRY(call)                  \
    do {                                \
        int rc__ = (call);              \
        if (rc__ != FSIM_OK) return rc__; \
    } while (0)

struct KernelTimer {
    double ms = 0.0;
    int64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

}  // namespace fsim

// The opaque handle of include/fusionsim.h.
struct fsim_sim {
    fsim_spec spec{};
    int prec = FSIM_F64;
    size_t rs = 8;  // sizeof(real)
    int device = 0;
    int nsm = 148;            // multiprocessors of the device (grid sizes of the grid-stride kernels)
    uint32_t smem_opt_in = 0; // kernels whose dynamic shared-memory limit was raised ON THIS DEVICE (one bit each)
    cudaStream_t stream = nullptr;
    // MEASURED ALTERNATIVE (FSIM_FLAG_POST_STREAM): stencil + canvas draws ("post" work: reads the per-cell sums,
    // writes the running average and the canvas) on a second, higher-priority stream so that they overlap the NEXT
    // frame's sweep.  post_begin forks it off the main stream, post_join makes the main stream wait for it (before
    // anything that reads the average / writes the sums or B).  Off by default: on B200 the overlap doubles the
    // sweep's time (the stencil's shared memory shrinks the L1 the sweep's gathers live on).
    cudaStream_t post_stream = nullptr;
    cudaEvent_t post_fork = nullptr, post_done = nullptr;
    bool post_pending = false;
    bool ext_stream = false;  // `stream` belongs to the caller (fsim_set_stream): collectives are stream-ordered
    bool sticky_error = false;

    // geometry
    int nr = 0, nz = 0;       // global grid
    int row0 = 0, rows = 0;   // local table rows (owned + halo, clipped to the grid)
    int own0 = 0, own_rows = 0;  // owned rows (== whole grid on one GPU)
    int64_t ncell_local = 0, ncell_global = 0;
    // planar grid fields (per-cell sums, running average, moments): channel q of local cell (i, j) at
    // base[q*plane + j*pitch + i]; the row pitch is padded to 16 bytes for the TMA tensor map
    int pitch = 0;
    int64_t plane = 0;
    alignas(64) unsigned char tm_sums[128] = {};  // CUtensorMap of the per-cell sums (deposit.cu)
    int tm_sums_rows = 0;                         // box height the map was encoded with
    bool slab = false;
    bool ring = false;        // FSIM_FLAG_PERIODIC_Z: nz owned rows + ghost rows either side that hold the wrapped rows

    // physical constants (host doubles, empic.js:44-46, :852)
    double h = 0, factor_r = 0, factor_z = 0, step_factor = 0;
    double k13 = 0, k31 = 0, kr = 0, kz = 0;  // N(...) literals of empic.js:527,606,647

    // particles
    int64_t n = 0, cap = 0;
    void *part[2][fsim::NPART_ARRAYS] = {};
    uint8_t *alive[2] = {};
    uint32_t *pid[2] = {};
    int cur = 0;
    bool ids_identity = true;   // storage order == id order
    bool fresh = true;          // nothing per particle has been set or computed since create / fsim_set_particle_count
    bool rand_default = true;   // the RNG state is the engine's default draw (a function of the particle id)
    uint32_t id_base = 0;

    // sort scratch
    uint32_t *key = nullptr;      // [cap] gather cell of each particle (local index)
    uint32_t *counts = nullptr;   // [ncell_local + 1]
    uint32_t *starts = nullptr;   // [ncell_local + 2]
    uint32_t *cursor = nullptr;   // [ncell_local + 1]
    uint32_t *blocksums = nullptr;
    uint32_t *perm = nullptr;     // [cap] particle slots ordered by cell (index sort)
    void *dcol[2] = {};           // [cap] sprite colour 0.001*(v_r, v_a) of each slot (deposit prepass); 0.001*v_z is
                                  // formed by the per-cell pass from the stored v_z (8 bytes less written per particle)
    bool keys_valid = false;      // key[] and the histogram in counts[] match the current positions
    bool counts_dirty = false;    // counts[] holds a histogram that no scan has consumed yet
    bool binned = false;          // starts[] and perm[] match the current positions
    bool ever_sorted = false;
    int steps_since_sort = 0;     // step() calls since the last physical sort
    bool resort_due = false;      // density() found the storage due for a re-sort: the next step() sweeps through perm[]

    // tables
    void *cellrec = nullptr;   // [ncell_local][RECSTRIDE]
    void *E = nullptr, *B = nullptr;  // [ncell_local][3]
    uint32_t *sink = nullptr;  // [(ncell_global + 31) / 32] bit c = 1: cell c keeps the particle (sink_mask.r > 0.5)
    void *entropy = nullptr;   // [1024*1024][4]
    void *invcdf = nullptr;    // [512*512][2]
    bool have_precalc = false;

    // deposit
    void *cellsum = nullptr;     // [ncell_local][4]
    uint32_t *cellcount = nullptr;  // [ncell_local]
    void *mom = nullptr, *norm = nullptr;  // [ncell_local][4] (FSIM_FLAG_KEEP_MOMENTS)
    void *avg = nullptr;         // [ncell_local][4]
    uint32_t *heavy_list = nullptr;   // cells summed by one block each (> 256 particles)
    uint32_t *medium_list = nullptr;  // cells summed by one warp each (17..256 particles)
    uint32_t *heavy_n = nullptr;      // [2]: lengths of heavy_list, medium_list
    uint32_t *oob = nullptr;     // particles whose gather row fell outside the local table

    // EXTENSION: self-consistent field solve (fieldsolve.cu), allocated at the first fsim_solve_fields()
    void *phi[2] = {};           // potential, planar [rows][pitch], ping-pong
    int phi_cur = 0;
    void *rho_src = nullptr;     // rho/eps0, planar
    void *background = nullptr;  // neutralising background in density units, planar (zero until fsim_set_field)
    void *relax_coef = nullptr;  // [nr][4] Jacobi coefficients cE cW cZ cB
    alignas(64) unsigned char tm_phi[2][128] = {};  // CUtensorMaps of phi[0], phi[1], rho_src
    alignas(64) unsigned char tm_src[128] = {};

    // EXTENSION: electromagnetic (Yee) field update (em.cu), allocated by fsim_em_init()
    void *em[6] = {};            // E_r E_z B_t E_t B_r B_z on their own lattices, row-major [j][i]
    void *em_B0 = nullptr;       // [ncell_local][3] the static B underneath
    void *em_coef = nullptr;     // [nr+1][6] per-column coefficients
    double em_weight = 0.0;      // macro weight the coefficients were formed with (NaN: not uploaded)
    bool em_on = false;

    // launch-bound scenes: a captured CUDA graph of one cycle of frames (fsim_run_frames, api.cu)
    uint64_t config_epoch = 0;        // bumped by every entry point that can change what a frame launches or works on
    cudaGraphExec_t frame_graph = nullptr;
    uint64_t graph_epoch = 0;         // config_epoch the graph was captured under
    int graph_frames = 0;             // frames per replay (2 x sort interval: the storage copies alternate)
    int64_t graph_launches = 0;       // kernel launches per replay
    int graph_phase[12] = {};         // host-side frame state at the start of the captured cycle
    bool graph_failed = false;        // capture was tried and did not work out: frames are launched one by one
    int64_t frames_run = 0;           // frames fsim_run_frames launched one by one (lazy initialisation is done after a cycle)
    int64_t graph_replays = 0;

    // staging
    void *stage = nullptr;
    size_t stage_bytes = 0;
    void *migr = nullptr;          // packed migration records (send side)
    size_t migr_bytes = 0;
    uint32_t *leavers = nullptr;   // [cap], slab mode: slots whose row left the slab (written by the sweep), then the hole list
    uint32_t *mscratch = nullptr;  // [MC_WORDS] small counters of the migration kernels (enum MC_*)
    uint8_t *hole_flag = nullptr;  // [cap] 1 = slot vacated by a leaver
    uint32_t nholes_host = 0;
    fsim::MigratePlan plan;        // asynchronous fixed-capacity exchange (fsim_migrate_setup)
    bool n_async = false;          // the exact particle count lives in mscratch[MC_NLIVE]; `n` is an upper bound
    uint32_t *n_pinned = nullptr;  // [2] pinned host copies of MC_NLIVE, read back without waiting
    cudaEvent_t n_event[2] = {};
    int64_t n_pending_bound[2] = {};  // arrivals that may have been added after read-back k was taken
    bool n_inflight[2] = {false, false};
    int n_slot = 0;
    void *halo_buf = nullptr;      // slab mode: [send_lo | send_hi | recv_lo | recv_hi], each 4 x 5 x nr reals
    bool conv_interior_done = false;  // fsim_density_interior ran: fsim_density_end convolves only the boundary tiles
    bool have_leavers = false;     // leavers[0..*nleavers) lists the slots whose row left the slab (emitted by the push)

    // spindle-cusp boundary solve (spindle.cu): what the last addSpindleCuspPlasmaField() found
    std::vector<double> spindle_x, spindle_currents, spindle_A, spindle_rhs;
    int spindle_iterations = 0;
    double spindle_diff = 0.0;

    // measurement
    bool timing = false;
    std::map<std::string, fsim::KernelTimer> timers;
    int64_t launches = 0;
    cudaEvent_t marks[16] = {};

    void *bmag = nullptr;          // [nr * own_rows] uchar4: the |B| layer of the canvas (first draw), kept until B changes
    bool bmag_valid = false;

    // asynchronous canvas read-back (fsim_render_rgba8_async): two device images, copy stream
    cudaStream_t copy_stream = nullptr;
    uint8_t *canvas_dev[2] = {};
    cudaEvent_t render_done[2] = {}, copy_done[2] = {};
    bool copy_pending[2] = {false, false};
    int canvas_slot = 0;
};

namespace fsim {

int ensure_stage(fsim_sim *s, size_t bytes);

// RAII kernel bracket: counts the launch and, when timing is on, records CUDA events on the
// handle's stream (the stream the kernel is launched on).
struct Bracket {
    fsim_sim *s;
    KernelTimer *t = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    Bracket(fsim_sim *sim, const char *name, cudaStream_t stream = nullptr) : s(sim), st(stream ? stream : sim->stream)
    {
        s->launches++;
        KernelTimer &kt = s->timers[name];
        kt.launches++;
        if (s->timing) {
            t = &kt;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, st);
        }
    }
    ~Bracket()
    {
        if (t) {
            cudaEventRecord(e1, st);
            t->pending.emplace_back(e0, e1);
        }
    }
};

// precision dispatch: f(double{}) or f(float{})
template <typename F>
inline int dispatch(const fsim_sim *s, F &&f)
{
    if (s->prec == FSIM_F64) return f(double{});
    return f(float{});
}

// ---- device helpers --------------------------------------------------------------------
// NEAREST + CLAMP_TO_EDGE texel index (utilities.js:528-531); NaN samples texel 0.
// Four instructions: the conversion itself maps NaN to 0 and saturates +-inf / out-of-range values
// (cvt.rzi.s32), so the clamp to [0, n-1] is done on the INTEGER (two IMNMX) -- an fp64 min/max
// clamp costs ~25 instructions per call, which was a third of the step kernel's instruction stream.
template <typename Real>
__device__ __forceinline__ int tex_idx(Real u, int n)
{
    if constexpr (sizeof(Real) == 8)
        return min(max(__double2int_rz(u * (Real)n), 0), n - 1);
    else
        return min(max(__float2int_rz(u * (Real)n), 0), n - 1);
}

__device__ __forceinline__ double fsqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float fsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double ffloor(double x) { return floor(x); }
__device__ __forceinline__ float ffloor(float x) { return floorf(x); }

// programPre1/2/3 (empic.js:524-527, :563-566, :603-606) from a cell record: R[0..8] row-major.
template <typename Real>
__device__ __forceinline__ void boris_rows(const Real *rec, Real h, Real k13, Real k31, Real (&R)[9])
{
    const Real Bx = rec[REC_BX], By = rec[REC_BY], Bz = rec[REC_BZ], f = rec[REC_F], one_m = rec[REC_ONEM];
    R[0] = one_m + f * h * h * Bx * Bx;
    R[1] = f * h * (Bz + h * Bx * By);
    R[2] = (f * h * (-By + h * Bx * Bz)) * k13;
    R[3] = f * h * (-Bz + h * By * Bx);
    R[4] = one_m + f * h * h * By * By;
    R[5] = (f * h * (Bx + h * By * Bz)) * k13;
    R[6] = (f * h * (By + h * Bz * Bx)) * k31;
    R[7] = (f * h * (-Bx + h * Bz * By)) * k31;
    R[8] = one_m + f * h * h * Bz * Bz;
}

// bit 31 of a sort key: the particle's sprite is clipped (not deposited)
constexpr uint32_t KEY_CLIPPED = 0x80000000u;
constexpr uint32_t KEY_MASK = 0x7fffffffu;

// Vertex shader of programMoments01 (empic.js:994-1006) for a particle at (x,y,z) with velocity v:
// colour 0.001*(v_r, v_a, v_z) (the alpha channel is the constant 0.001*1.0) and the sort key =
// gather cell (clamped like the texture fetch), with KEY_CLIPPED when the sprite centre lies
// outside the target (GLES2 point clipping), is NaN, or is not in a row this rank owns.
// `r` = sqrt(x*x + y*y) is passed in because the push kernel already has it.
template <typename Real>
__device__ __forceinline__ uint32_t sprite_key_colour(Real x, Real y, Real z, Real r, Real vx, Real vy,
                                                      Real vz, int nr, int nz, int row0, int rows,
                                                      int own_lo, int own_hi, Real &c0, Real &c1, Real &c2)
{
    const Real dx = x / r, dy = y / r;
    const Real vr = vx * dx + vy * dy;
    const Real va = vy * dx - vx * dy;
    c0 = (Real)FSIM_DEPOSIT_WEIGHT * vr;
    c1 = (Real)FSIM_DEPOSIT_WEIGHT * va;
    c2 = (Real)FSIM_DEPOSIT_WEIGHT * vz;
    const int gj = tex_idx(z, nz);
    int cj = gj - row0;
    const bool local = cj >= own_lo && cj < own_hi;  // deposited only by the rank that owns the row
    cj = cj < 0 ? 0 : (cj >= rows ? rows - 1 : cj);
    uint32_t key = (uint32_t)tex_idx(r, nr) + (uint32_t)cj * (uint32_t)nr;
    const Real xw = r * (Real)nr, yw = z * (Real)nz;
    const bool inside = (xw >= (Real)0) && (xw < (Real)nr) && (yw >= (Real)0) && (yw < (Real)nz);
    if (!inside || !local) key |= KEY_CLIPPED;
    return key;
}

// Warp-level aggregation over RUNS of equal keys in consecutive lanes (the storage is nearly in cell
// order, so equal keys sit next to each other).  Returns the lane that leads this lane's run, the
// run length and this lane's rank in it -- from one shuffle and one ballot, no MATCH.ANY.
// Equal keys in non-adjacent lanes simply form separate runs (one atomic each): still correct.
__device__ __forceinline__ void warp_runs(uint32_t key, int lane, int &leader, uint32_t &len, uint32_t &rank)
{
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    leader = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));
    const unsigned after = heads & ~(0xffffffffu >> (31 - lane));  // heads strictly above this lane
    const int next = after ? (__ffs((int)after) - 1) : 32;
    len = (uint32_t)(next - leader);
    rank = (uint32_t)(lane - leader);
}

#ifdef FSIM_TUNE
extern int g_push_variant, g_conv_variant;  // tuning build only (fsim_tune_set)
#endif

// empic.js:317: the quadrature angle is formed in the shader's working precision, its cosine is the host libm's
inline void host_cos_tables(double *c64, float *c32)
{
    for (int k = 0; k < FSIM_NQUAD; ++k) {
        c64[k] = cos(FSIM_PI_GLSL * ((double)k + 0.5) / 1000.0);
        const float a = (float)FSIM_PI_GLSL * ((float)k + 0.5f) / 1000.0f;
        c32[k] = (float)cos((double)a);
    }
}

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// exact particle count inside a kernel: the device word when the exchange is asynchronous, else the host's n
__device__ __forceinline__ int64_t live_count(const uint32_t *n_dev, int64_t n_host)
{
    if (!n_dev) return n_host;
    const int64_t d = (int64_t)*n_dev;
    return d < n_host ? d : n_host;
}

// kernels / host stages implemented in the other translation units
int launch_push(fsim_sim *s, bool with_hist, int nhalf, bool resort = false);  // nhalf half-steps in one sweep (+ fused re-sort)
int launch_keys(fsim_sim *s);      // deposit prepass from the stored state: key, colour, histogram
int launch_bin(fsim_sim *s);       // scan + index scatter -> starts[], perm[]
int launch_apply_perm(fsim_sim *s);  // physical re-sort: storage <- storage[perm]
int launch_cellsum(fsim_sim *s);
int launch_cellsum_atomic(fsim_sim *s);
int make_sums_tensor_map(fsim_sim *s, int box_rows);
int launch_halo_pack(fsim_sim *s);
int launch_halo_unpack(fsim_sim *s);
int launch_conv_rows(fsim_sim *s, int part);  // part 0: all owned rows, 1: rows that need no halo, 2: the rest
int settle_count(fsim_sim *s);               // asynchronous exchange: make fsim_sim::n exact again (synchronises)
int check_handle(fsim_sim *s);               // sticky-error / device check of every entry point (does not join the post stream)
cudaStream_t post_begin(fsim_sim *s);        // fork: the post stream sees everything enqueued on the main stream so far
int post_end(fsim_sim *s);                   // marks the end of the post work enqueued since post_begin
int post_join(fsim_sim *s);                  // the main stream waits for pending post work
int launch_precalc(fsim_sim *s);
int launch_expand_records(fsim_sim *s, double *dev_out);  // [cells][12] R1 R2 R3 A as doubles
int launch_add_loop(fsim_sim *s, double R, double Z, double I);
int launch_add_uniform(fsim_sim *s, int kind, double val);
int spindle_solve(fsim_sim *s, double coil_r, double B_c, double beta_c);  // spindle.cu
int launch_render(fsim_sim *s, uint8_t *dev_rgba, cudaStream_t st);
int ensure_fieldsolve(fsim_sim *s);
// em.cu
int em_init(fsim_sim *s);
int em_step(fsim_sim *s, double macro_weight, bool with_current);
void em_free(fsim_sim *s);
int em_field_index(const std::string &name);
int64_t em_field_count(const fsim_sim *s, int field);
int launch_charge_source(fsim_sim *s, const void *dens_a, double rho_scale);
int ring_wrap_rows(fsim_sim *s, void *plane_base, int nrows);  // periodic z: owned boundary rows -> ghost rows (planar field)
int launch_relax(fsim_sim *s, int sweeps, double omega);  // 1..4 sweeps, one launch
int launch_efield(fsim_sim *s);
int launch_plane_out(fsim_sim *s, const void *plane, double *dev_out);

}  // namespace fsim
