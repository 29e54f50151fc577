/* fusionsim.h -- C ABI of libfusionsim.so, the B200-native particle-step engine that
 * replaces the WebGL back end of kcdodd/fusion-sim's
 *     empic.makeCylindricalParticlePusher(spec)      (public/javascripts/empic.js:30)
 * One fsim_* entry point per member of the object that constructor returns
 * (empic.js:60, :1157-:1526).  Plain pointers and sizes only; all array arguments are
 * HOST pointers (doubles = JS Numbers) copied during the call; the handle owns every
 * device buffer.  Every function returns 0 on success or an FSIM_ERR_* code;
 * fsim_last_error() gives the message the JS shim turns into `throw new Error(msg)`
 * (the reference reports every failure by a synchronous throw: utilities.js:118-127,
 * :213-260, :678-680).  There is no CPU fallback: without a CUDA device fsim_create
 * fails with FSIM_ERR_CUDA.
 *
 * Array conventions (flattened JS nesting, row-major):
 *   field  value.E / value.B          [nr][nz][3]  -> (i*nz + j)*3 + k   empic.js:1160-1166
 *   grid   value.sink_mask/source_pdf [nr][nz]     -> i*nz + j           empic.js:1247-1250
 *   particle value.position/velocity  [N][3]       -> 3*p + k            empic.js:1201-1205
 * Units as in the reference: positions in metres, velocities in units of c, fields in
 * tesla and V/m; fsim_set_* normalises exactly as empic.js:1202-1204, :1226-1228.
 * Accessors return the NORMALISED device state (r in [0,1], z in [0,1]).
 * Device textures are addressed texel (i,j) -> i + j*nr (empic.js:1162): cell-indexed
 * accessors use that order.
 *
 * Lifetime of input arrays: every setter enqueues its host->device copy on the handle's stream and
 * returns.  Arrays in pageable memory may be reused as soon as the call returns (the driver stages
 * them); arrays in page-locked (pinned) memory are read asynchronously and must stay unchanged until
 * the next fsim_sync() or accessor call.  Limits: particle slots per handle < 2^31, id_base + slots
 * < 2^32, nr*nz < 2^31.
 */
#ifndef FUSIONSIM_H
#define FUSIONSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSIM_ABI_VERSION 1

enum {
    FSIM_OK = 0,
    FSIM_ERR_INVALID = 1,     /* bad argument / spec validation (utilities.js:118-127)      */
    FSIM_ERR_CUDA = 2,        /* CUDA runtime failure; sticky on the handle                 */
    FSIM_ERR_UNSUPPORTED = 3, /* not available in this mode (e.g. solveFields on a slab handle)  */
    FSIM_ERR_STATE = 4,       /* call order violated (e.g. step before precalc)             */
    FSIM_ERR_RANGE = 5        /* "function out of range" (empic.js:1294-1296), capacity     */
};

enum { FSIM_F64 = 0, FSIM_F32 = 1 };

/* flags */
#define FSIM_FLAG_CORRECTED_PREA  1u /* textbook h(E.B)B instead of the scalar add of empic.js:645 */
#define FSIM_FLAG_KEEP_MOMENTS    2u /* density() also stores moments01 and moments01_norm          */
#define FSIM_FLAG_ATOMIC_DEPOSIT  4u /* measured alternative: global-atomic per-cell sums           */
#define FSIM_FLAG_UNFUSED_SORT   16u /* measurement: physical re-sort as a pass of its own in density()
                                      * (round-1 behaviour) instead of fused into the next step()'s sweep  */
#define FSIM_FLAG_PERIODIC_Z    128u /* EXTENSION (the reference has no periodic
                                      * boundary: CLAMP_TO_EDGE textures and absorbing sink rows): z is periodic --
                                      * a pushed position wraps (z <- z - floor(z), exactly 1 -> 0) before the sink
                                      * lookup, the 11x11 deposit footprint wraps in z, and the field solve's Poisson
                                      * stencil and gradient wrap in z instead of the grounded end walls.  Single GPU.
                                      * The local tables carry 8 ghost rows either side of the nz owned rows
                                      * (fsim_local_cells = nr (nz + 16)): cell-indexed accessors return them too  */
#define FSIM_FLAG_POST_STREAM     8u /* measured alternative: stencil and canvas draws on a second stream, under the
                                      * next frame's sweep (slower on B200: the stencil's shared memory is L1 the
                                      * sweep's gathers lose while the two share an SM)      */
#define FSIM_FLAG_POST_NO_PRIORITY 64u /* ... that stream at the main stream's priority instead of above it       */

typedef struct fsim_sim fsim_sim;

/* spec of empic.js:31-41 (first 8 fields, same names) plus the extension fields the
 * headless build needs.                       */
typedef struct fsim_spec {
    double radius;          /* metres                                                     */
    double height;          /* metres                                                     */
    int64_t nr;             /* grid cells along r                                         */
    int64_t nz;             /* grid cells along z (GLOBAL grid)                           */
    double dt;              /* seconds                                                    */
    int64_t nparticles;     /* SIDE of the particle texture: N = nparticles^2 (:107-109)  */
    double particle_mass;   /* kg                                                         */
    double particle_charge; /* C                                                          */
    /* ---- extensions ---- */
    int32_t precision;      /* FSIM_F64 (default) | FSIM_F32 (mirrors RGBA32F storage)    */
    int32_t device;         /* CUDA device ordinal                                        */
    uint32_t flags;         /* FSIM_FLAG_*                                                */
    int32_t sort_interval;  /* physical re-sort of the particle storage by cell: density() does it every
                             * k-th frame (0 = the default, 8); a loop of step() calls without density()
                             * re-sorts after 4*k steps                                                */
    int64_t nparticles_total; /* if > 0: particle count, overrides nparticles^2           */
    int64_t capacity;       /* particle slots to allocate (>= count; 0 = count)           */
    int64_t slab_row0;      /* multi-GPU slab: first grid row (z index) owned             */
    int64_t slab_rows;      /* rows owned; 0 = whole grid (single GPU)                    */
    int64_t halo_rows;      /* extra table rows kept either side of the slab              */
    uint64_t id_base;       /* global index of local particle 0 (multi-GPU)               */
} fsim_spec;

const char *fsim_last_error(void);
int fsim_abi_version(void);

/* ---- constructor / destructor : empic.js:30 (the reference never frees) ---------------- */
int fsim_create(const fsim_spec *spec, fsim_sim **out);
int fsim_destroy(fsim_sim *sim);

/* ---- out.set(value), empic.js:1157-1350; one call per optional property ----------------- */
int fsim_set_E(fsim_sim *sim, const double *E);                 /* value.E         :1159-1177 */
int fsim_set_B(fsim_sim *sim, const double *B);                 /* value.B         :1179-1197 */
int fsim_set_position(fsim_sim *sim, const double *pos);        /* value.position  :1199-1221 */
int fsim_set_velocity(fsim_sim *sim, const double *vel);        /* value.velocity  :1223-1244 */
int fsim_set_sink_mask(fsim_sim *sim, const double *mask);      /* value.sink_mask :1246-1260 */
int fsim_set_source_pdf(fsim_sim *sim, const double *pdf, int64_t n0, int64_t n1); /* :1263-1349 */
/* extensions: the reference draws these from Math.random / window.crypto (empic.js:148-173) */
int fsim_set_rand(fsim_sim *sim, const double *rnd);            /* [N][4] in [0,1]            */
int fsim_set_entropy(fsim_sim *sim, const double *entropy);     /* [1024*1024][4] in [0,1]    */
int fsim_set_inv_cdf(fsim_sim *sim, const double *table);       /* [512*512][2], texel i+j*512 */
int fsim_set_particle_count(fsim_sim *sim, int64_t n);          /* multi-GPU: live count <= capacity; ids := id_base + slot */
int fsim_set_ids(fsim_sim *sim, const uint64_t *ids);           /* multi-GPU: global particle ids */

/* ---- static field builders, blended ONE,ONE into B ---------------------------------------- */
int fsim_add_current_loop(fsim_sim *sim, double r, double z, double I); /* addCurrentLoop :1352 */
int fsim_add_current_z(fsim_sim *sim, double I);                        /* addCurrentZ    :1380 */
int fsim_add_bz(fsim_sim *sim, double Bz);                              /* addBZ          :1391 */
int fsim_add_btheta(fsim_sim *sim, double Btheta);                      /* addBTheta      :1402 */
/* out.addSpindleCuspPlasmaField(r, B_c, beta_c), empic.js:1369.  spindle.makeSpindleCuspPlasmaField does not run
 * in the reference, so this entry point implements the INTENT of spindle.js:27-30 and
 * :632-654 -- "solves the boundary conditions for a perfect conductor in center of a spindle cusp magnetic field",
 * then superposes the solution on B -- with this specification (parity unpinned; oracle:
 * oracle/fsim_oracle_spindle_impl.h, bit-identical):
 *   field added = two opposing coils of radius r [m], +I_c at z = 0 and -I_c at z = height (the demo's cusp,
 *     fusionsim.js:137-138), I_c = 2 r B_c / mu0 (B_c [T] = field of one coil at its own centre), PLUS the surface
 *     currents that make B.n = 0 on the plasma surface, scaled by 1 - sqrt(1 - beta_c) (1: full exclusion, 0: vacuum);
 *   surface (lower half, spindle.js:138-147; the upper half is its mirror image about z = height/2 carrying the
 *     opposite current, :357-366): x = R cos(-phi) + radius, z = s R sin(-phi), R = radius sqrt(1 + a^2), a = 0.4,
 *     phi in [theta, theta + arc], theta = atan(a) + pi, arc = pi/2 - 2 atan(a), s = height / (2 radius)
 *     (s = 1 in the demo: from the axis point (0, a radius) to the ring cusp (radius (1-a), height/2));
 *   L = 256 elements (makeSORIterative n_power 3, :64): nodes at phi_l = theta + l arc / L, l = 0..L (the
 *     reference divides by 1000, a slip that covers a quarter of the arc); node 0 is put on the axis (x = 0);
 *     collocation point p at l = p + 1/2 with unit normal (-s cos(-phi), -sin(-phi)) / |.|;
 *   unknown x_e = strength of element e: a loop of current +x_e through node e and one of -x_e through node e+1
 *     (:161-176, signs :351-391), each with its mirror image of opposite sign;  A[p][e] = n_p . (field of element e
 *     at point p), rhs[p] = -n_p . (field of the coils at p);  the constant vector is a null vector of A (node l
 *     carries x_l - x_{l-1}): gauge x_{L-1} = 0, the last collocation point (the ring-cusp tip) is dropped, row and
 *     column L-1 of the solved system are those of the identity;
 *   loop field [T/A] at (x, z) of a loop of radius Rl at height Zl -- the quadrature of programCurrentLoopShape
 *     (empic.js:308-326) at the exact relative position, in metres, midpoint-rule weight W = 2 pi / 1000:
 *       K = Rl W mu0 / (4 pi);  for k < 1000: c = cos(pi (k + .5) / 1000), rho = sqrt(Rl^2 + x^2 + dz^2 - 2 x Rl c),
 *       f = rho > 0 ? K / (rho rho rho) : 0,  B_r += dz f c,  B_z += f (Rl - x c),  dz = z - Zl   (left to right, no FMA);
 *   solver: fsim_jacobi_* (matrix_webgl.makeSORIterative), fp64, relaxation 1, tolerance 1e-9, substep 64, at
 *     most 4000 checks (the Jacobi spectral radius of this system is 0.999: the reference's tolerance 1e-3 and
 *     10 iterations cannot converge);  FSIM_ERR_RANGE if it does not converge;
 *   B(cell) += sum over loops, in the order coil z=0, coil z=height, then node 0, mirror 0, node 1, mirror 1, ...,
 *     of I_l x (loop field at the cell centre ((i+.5) radius/nr, (j+.5) height/nz)), in the engine's precision.   */
int fsim_add_spindle_cusp_plasma_field(fsim_sim *sim, double r, double B_c, double beta_c);
/* what the last call found (any pointer may be NULL): element strengths x [256], node currents [257] in amperes,
 * the solved system A [256][256] row-major and rhs [256], solver checks and final diff (matrix_webgl.js:687)   */
int fsim_get_spindle(fsim_sim *sim, double *x, double *currents, double *A, double *rhs, int32_t *iterations,
                     double *diff);

/* ---- precalc / step / density / canvas ----------------------------------------------------- */
int fsim_precalc(fsim_sim *sim);   /* out.precalc :1413-1434: (E,B) -> R1,R2,R3,A per cell      */
int fsim_step(fsim_sim *sim);      /* out.step    :1436-1469: TWO leap-frog half-steps         */
int fsim_half_step(fsim_sim *sim); /* extension: one half-step (rand, velocity, position)      */
int fsim_density(fsim_sim *sim);   /* out.density :1471-1495: deposit, normalise, running avg  */
int fsim_render_rgba8(fsim_sim *sim, uint8_t *rgba); /* out.canvas :60 after :1497-1504; [nz][nr][4], top row first */
/* extension: same image, but the device->host copy runs on a second stream and overlaps the next
 * frame; `rgba` (pinned memory for a truly asynchronous copy) is complete after the next fsim_sync()
 * or after two further fsim_render_rgba8_async() calls (two device buffers alternate).           */
int fsim_render_rgba8_async(fsim_sim *sim, uint8_t *rgba);
int fsim_render_rows_async(fsim_sim *sim, uint8_t *rows); /* same, but `rows` holds only the owned rows
                                                            * [slab_rows][nr][4] (a slab rank's share)      */
int fsim_draw_canvas(fsim_sim *sim); /* the two canvas draws of out.density (:1497-1504) into the device-
                                      * resident canvas, no read-back (the reference's canvas stays on the GPU) */
/* extension: `nframes` iterations of the page loop (fusionsim.js:170-178) = nframes x (fsim_step, fsim_density,
 * fsim_draw_canvas), bit-identical to calling them one by one.  Scenes whose frame is bound by launch latency (the
 * reference's demo scene: ~14 launches of 3-30 us) gain from it: after one cycle of 2 x sort_interval frames launched
 * one by one, the next cycle is captured into a CUDA graph and replayed while whole cycles remain; any entry point
 * that can change what a frame launches drops the graph.  One GPU (a slab's frame has exchanges between its parts).
 * fsim_frame_graph_info: frames and kernel launches per captured cycle (0 = no graph in use), replays so far.      */
int fsim_run_frames(fsim_sim *sim, int64_t nframes);
int fsim_frame_graph_info(fsim_sim *sim, int32_t *frames_per_cycle, int64_t *launches_per_cycle, int64_t *replays);
int fsim_sort(fsim_sim *sim);      /* extension: re-sort particle storage by cell now          */
int fsim_sync(fsim_sim *sim);      /* wait for the handle's stream                             */

/* ---- EXTENSION, no reference counterpart: self-consistent
 * electrostatic field solve.  The reference pushes test particles in static fields; this closes
 * the loop: rho = q * macro_weight * n from the deposited density (source 0: the running average
 * moments01_avg.a, 1: moments01_norm.a of the last density(), needs FSIM_FLAG_KEEP_MOMENTS),
 * `sweeps` weighted-Jacobi sweeps (the iteration of matrix_webgl.makeSORIterative,
 * matrix_webgl.js:224-300) on the 5-point cylindrical Poisson operator with grounded walls, warm
 * started from the previous potential, E = -grad(phi) (overwrites E), then precalc().
 *   cells (i,j) centred at r = (i+.5) dr, z = (j+.5) dz, dr = radius/nr, dz = height/nz;
 *   src = (q macro_weight / (pi radius dr dz eps0)) * (density.a - background)   (background: fsim_set_field, default 0);
 *   per-column coefficients in host fp64, rounded to the engine's real type:
 *     aE = (i+1)/((i+.5) dr^2), aW = i/((i+.5) dr^2), aZ = 1/dz^2, aC = aE + aW + 2 aZ,
 *     cE = aE/aC, cW = aW/aC, cZ = aZ/aC, cB = 1/aC     (finite volumes on rings; no axis ghost);
 *   one sweep, in exactly this order, no fused multiply-add, ghost cells beyond r = radius,
 *   z = 0 and z = height hold phi = 0:
 *     t = ((cE phi_E + cW phi_W) + cZ (phi_N + phi_S)) + cB src;  phi' = omega t + (1 - omega) phi;
 *   E_r = -((phi_E - phi_W) / (2 dr)) with phi_W := phi at i = 0, E_z = -((phi_N - phi_S) / (2 dz))
 *   (both as a multiplication by the host-computed reciprocal), E_theta = 0.            Single GPU. */
int fsim_solve_fields(fsim_sim *sim, double macro_weight, int32_t sweeps, double omega, int32_t source);
/* The same solve on a slab (one process per GPU): stage 0 = charge source, 1 = `sweeps` (1..4)
 * Jacobi sweeps in one launch, 2 = E = -grad(phi), 3 = precalc().  Every stage works on all local
 * rows; between the stages the caller copies boundary rows to the neighbouring ranks' halo rows:
 * 4 rows of "rho_src" after stage 0, 4 rows of "phi" before every stage 1 and before stage 2,
 * halo_rows rows of "E" before stage 3.  fsim_field_rows gives the device address of rows
 * [first_row, first_row + nrows) of the local table (rows are contiguous).  Bit-identical to the
 * single-GPU solve: each rank recomputes the halo cells it needs with the same arithmetic.       */
int fsim_solve_fields_stage(fsim_sim *sim, int32_t stage, double macro_weight, int32_t sweeps, double omega,
                            int32_t source);
int fsim_field_rows(fsim_sim *sim, const char *name, int64_t first_row, int64_t nrows, void **ptr, int64_t *nbytes);

/* ---- EXTENSION, NO REFERENCE COUNTERPART (BASELINE.json configs[2] "Boris + Yee
 * FDTD"): electromagnetic field update.  The reference's E and B never change (value.E / value.B, empic.js:1159-1197);
 * this advances them with Maxwell's equations on an axisymmetric Yee mesh laid over the cells of the field
 * textures, with the current the deposit measures as the source.  PARITY UNPINNED by construction.
 *   cell (i,j) = [i dr, (i+1) dr] x [j dz, (j+1) dz];  all arrays row-major [j][i]:
 *     "Er" (i+1/2, j) [nz+1][nr]     "Ez" (i, j+1/2) [nz][nr+1]     "Bt" (i+1/2, j+1/2) [nz][nr]      TM set
 *     "Et" (i, j)     [nz+1][nr+1]   "Br" (i, j+1/2) [nz][nr+1]     "Bz" (i+1/2, j)     [nz+1][nr]    TE set
 *   perfectly conducting wall at r = radius and end plates at z = 0, height: the tangential E there (Ez at i = nr,
 *   Et at i = nr and j = 0, nz, Er at j = 0, nz) is never updated; Et on the axis likewise.
 * fsim_em_init: zero Yee fields; the B present at the call stays underneath as the static field B0 (changes made to
 *   E or B through the other entry points afterwards are overwritten by the next fsim_em_step; call fsim_em_init
 *   again after changing the static field).  Fails unless c dt sqrt(1/dr^2 + 1/dz^2) < 1.  One GPU, not periodic.
 * fsim_em_set / fsim_em_get: one of the six arrays above, host doubles.
 * fsim_em_step: ONE leap-frog step of spec.dt -- pair it with fsim_half_step() + fsim_density(), which advance the
 *   particles by the same dt and leave the moments.  Every product formed as written, left to right, no fused
 *   multiply-add; scalar and per-column coefficients in host fp64, rounded to the engine's real type:
 *     kz = dt/dz, kr = dt/dr, cz = c^2 dt/dz, cr = c^2 dt/dr, cj = dt/eps0, ax = 4 c^2 dt/dr,
 *     a1_i = dt (i+1) dr/((i+1/2) dr dr), a0_i = dt i dr/((i+1/2) dr dr),
 *     b1_i = c^2 dt (i+1/2) dr/(i dr dr), b0_i = c^2 dt (i-1/2) dr/(i dr dr)   (i >= 1)
 *   B from E:
 *     Br += kz (Et[j+1][i] - Et[j][i])
 *     Bt -= (kz (Er[j+1][i] - Er[j][i])) - (kr (Ez[j][i+1] - Ez[j][i]))
 *     Bz -= (a1_i Et[j][i+1]) - (a0_i Et[j][i])
 *   E from the new B and the current density J at the cell centres:
 *     Er += (-(cz (Bt[j][i] - Bt[j-1][i]))) - cj (0.5 (Jr[j-1][i] + Jr[j][i]))                       1 <= j <= nz-1
 *     Et += ((cz (Br[j][i] - Br[j-1][i])) - (cr (Bz[j][i] - Bz[j][i-1])))
 *           - cj (0.25 (((Jt[j-1][i-1] + Jt[j-1][i]) + Jt[j][i-1]) + Jt[j][i]))                      1 <= i <= nr-1, 1 <= j <= nz-1
 *     Ez += ((b1_i Bt[j][i]) - (b0_i Bt[j][i-1])) - cj (0.5 (Jz[j][i-1] + Jz[j][i]))                 1 <= i <= nr-1
 *     Ez[j][0] += (ax Bt[j][0]) - cj Jz[j][0]                                                        (axis: 4 B/dr)
 *   J_q[j][i] = g_q,i * moments01.q of cell (i,j) (the raw deposit of the last fsim_density(), needs
 *   FSIM_FLAG_KEEP_MOMENTS; J = 0 when with_current == 0):  g_r,i = g_t,i = G_i radius, g_z,i = G_i height,
 *     G_i = particle_charge macro_weight 1000 c / (2 pi u_i radius dr dz),  u_i = (i + 1/2)/nr
 *   (the sprites deposit 0.001 v in the normalised units v/c (1/radius, 1/radius, 1/height); 2 pi r dr dz is the
 *   volume of the cell's ring);
 *   then the fields the push gathers, at the cell centres, and precalc():
 *     E = (0.5 (Er[j][i] + Er[j+1][i]), 0.25 (((Et[j][i] + Et[j][i+1]) + Et[j+1][i]) + Et[j+1][i+1]), 0.5 (Ez[j][i] + Ez[j][i+1]))
 *     B = B0 + (0.5 (Br[j][i] + Br[j][i+1]), Bt[j][i], 0.5 (Bz[j][i] + Bz[j+1][i]))                              */
int fsim_em_init(fsim_sim *sim);
int fsim_em_set(fsim_sim *sim, const char *name, const double *data);
int fsim_em_get(fsim_sim *sim, const char *name, double *out);
int fsim_em_step(fsim_sim *sim, double macro_weight, int32_t with_current);

/* ---- checkpoint restore (extension; the reference can neither read nor restore its state) ---------
 * fsim_set_state is the exact inverse of fsim_get_position / _velocity / _rand: normalised units,
 * particle-id order, alive flag in position[..][3]; any pointer may be NULL.  fsim_set_field restores
 * "moments01_avg" [cells][4] (the running average of density()), "phi" [cells], or "background" [cells]:
 * the neutralising background the field solve subtracts from the density, in the units of the density
 * texture (e.g. the species' own density at t = 0 = immobile ions; zero until set).               */
int fsim_set_state(fsim_sim *sim, const double *position4, const double *velocity3, const double *rand4);
int fsim_set_field(fsim_sim *sim, const char *name, const double *data);

/* ---- accessors (extension; the reference exposes none) ---------- */
int64_t fsim_particle_count(const fsim_sim *sim);
int64_t fsim_local_cells(const fsim_sim *sim); /* nr * (slab_rows + 2*halo_rows) or nr*nz      */
int fsim_get_position(fsim_sim *sim, double *out);  /* [N][4]: x,y,z,alive in particle-id order */
int fsim_get_velocity(fsim_sim *sim, double *out);  /* [N][3]                                   */
int fsim_get_rand(fsim_sim *sim, double *out);      /* [N][4]                                   */
int fsim_get_ids(fsim_sim *sim, uint64_t *out);     /* [N] ids in STORAGE order                 */
int fsim_get_cells(fsim_sim *sim, int64_t *out);    /* [N] gather cell i+j*nr of each particle  */
/* name: "E","B","R1","R2","R3","A" -> [cells][3]; "cell_sums","moments01","moments01_norm",
 * "moments01_avg" -> [cells][4]; "inv_cdf" -> [512*512][2]; "entropy" -> [1024*1024][4];
 * "phi", "rho_src" -> [cells] (after fsim_solve_fields)                                          */
int fsim_get_field(fsim_sim *sim, const char *name, double *out);
int fsim_get_cell_count(fsim_sim *sim, uint32_t *out); /* [cells] particles deposited per cell   */
int fsim_get_sink_mask(fsim_sim *sim, uint8_t *out);   /* [nr*nz] 1 = keep, 0 = absorb           */

/* ---- run invariants (extension): reduced on the device, printed by bench.py with every line ---------
 * out[0] live particles, out[1] XOR of their ids, out[2] sum of their ids mod 2^64, out[3] particles the
 * last density() deposited on the owned rows; *sum_alpha (may be NULL) = sum of the weight channel of the
 * per-cell sums over the owned rows.  Over all ranks of a slab run out[0..2] must equal those of the ids
 * 0..N-1: the migration neither lost nor duplicated a particle.                                          */
int fsim_check_digest(fsim_sim *sim, uint64_t *out /* [4] */, double *sum_alpha);

/* ---- measurement hooks (extension) ----------------------------------------------------------- */
/* Per-kernel device time (ms, CUDA events on the handle's stream) accumulated since the last
 * reset, and launch counts.  names: "push","hist","scan","scatter","cellsum","conv","render".   */
int fsim_timing_enable(fsim_sim *sim, int on);
int fsim_timing_reset(fsim_sim *sim);
int fsim_timing_get(fsim_sim *sim, const char *name, double *ms, int64_t *launches);
int64_t fsim_launch_count(const fsim_sim *sim); /* kernels launched by this handle so far          */
/* CUDA-event stopwatch on the handle's stream (the stream every kernel above is launched on):
 * fsim_mark records event `slot` (0..15); fsim_elapsed_ms waits for slot b and returns b - a.  */
int fsim_mark(fsim_sim *sim, int slot);
int fsim_elapsed_ms(fsim_sim *sim, int slot_a, int slot_b, double *ms);

/* ---- multi-GPU slab exchange (extension) -------------------------------- */
/* Particles leaving the slab are packed into a device buffer grouped by destination rank;
 * row_bounds[k]..row_bounds[k+1] are the grid rows rank k owns.  The caller (one process per
 * GPU) moves the packed records with its own collective and hands them to fsim_migrate_unpack.  */
/* Put the handle on a stream the caller owns (cudaStream_t), so that kernels and the caller's
 * collectives are ordered by the stream instead of by host synchronisation.                        */
int fsim_set_stream(fsim_sim *sim, void *cuda_stream);
int64_t fsim_migrate_record_bytes(const fsim_sim *sim);
int fsim_migrate_pack(fsim_sim *sim, const int64_t *row_bounds, int32_t nranks, int32_t self,
                      int64_t *send_counts /* host [nranks] */, void **send_buf_dev);
int fsim_migrate_unpack(fsim_sim *sim, const void *recv_buf_dev, int64_t nrecv);
/* ASYNCHRONOUS exchange (no host round trip in the frame).  fsim_migrate_setup fixes, once, the capacity in
 * records of the region this rank sends to / receives from every other rank and returns the two device
 * buffers and the byte size of every region (16-byte header holding the record count, then the records);
 * region k of the send buffer goes to rank k, region k of the receive buffer comes from rank k.  From then
 * on the handle keeps the exact particle count on the device.  Per frame: fsim_migrate_begin (pack; nothing
 * is read back), the caller's all-to-all of the fixed-size regions on the handle's stream,
 * fsim_migrate_end (arrivals fill the holes, compaction, new count: all from device-side counts).
 * A region too small for a frame's leavers, or arrivals beyond the particle capacity, are reported by the
 * next fsim_sync() as FSIM_ERR_RANGE (the state is then invalid; fsim_migrate_pack/_unpack is the exact,
 * host-synchronising alternative for arbitrary volumes).                                               */
int fsim_migrate_setup(fsim_sim *sim, const int64_t *row_bounds, int32_t nranks, int32_t self,
                       const int64_t *send_caps /* [nranks] records */, const int64_t *recv_caps,
                       void **send_buf_dev, void **recv_buf_dev,
                       int64_t *send_region_bytes /* out [nranks] */, int64_t *recv_region_bytes);
int fsim_migrate_begin(fsim_sim *sim);
int fsim_migrate_end(fsim_sim *sim);
int fsim_migrate_stats(fsim_sim *sim, int64_t *sent_total); /* records packed so far (synchronises) */
/* Halo of the per-cell sums (5 rows each side) for the slab convolution. */
int fsim_halo_ptrs(fsim_sim *sim, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi,
                   int64_t *bytes_each);
/* measured alternative (replicated tables, index-sharded particles): the per-cell sums and counts of
 * every rank are added between fsim_density_begin and fsim_density_end; these are their addresses. */
int fsim_cellsum_ptrs(fsim_sim *sim, void **sums, int64_t *sum_bytes, void **counts, int64_t *count_bytes);
int fsim_density_begin(fsim_sim *sim); /* sort + per-cell sums + halo pack (before the halo exchange) */
int fsim_density_interior(fsim_sim *sim); /* optional, slab mode: stencil + normalise + running average on the
                                        * rows that read no halo row -- runs while the exchange is in flight */
int fsim_density_end(fsim_sim *sim);   /* halo unpack, then the stencil on the remaining rows (all owned rows
                                        * if fsim_density_interior was not called)                        */

/* ---- dense weighted-Jacobi solver: matrix_webgl.makeSORIterative (matrix_webgl.js:35-711) ------------
 * "next" row N3 of SURVEY.md section 8f.  vec_length L = 4 (2^n_power)^2; A is [L][L] row-major.
 * FSIM_JACOBI_LITERAL reproduces two defects of the reference (row gather of programResult
 * :408-411, never-reset statistics :628-634); without it the solver does what was evidently meant. */
#define FSIM_JACOBI_LITERAL 1u
typedef struct fsim_jacobi fsim_jacobi;
int fsim_jacobi_create(int32_t n_power, double relaxation, int32_t precision, int32_t device, uint32_t flags,
                       fsim_jacobi **out);                                   /* makeSORIterative :35 */
int fsim_jacobi_destroy(fsim_jacobi *j);
int64_t fsim_jacobi_vec_length(const fsim_jacobi *j);                        /* out.vec_length :52  */
int fsim_jacobi_set_matrix(fsim_jacobi *j, const double *A);                 /* out.set_matrix :455 */
int fsim_jacobi_set_b(fsim_jacobi *j, const double *b);                      /* out.set_b :482      */
int fsim_jacobi_init_vector(fsim_jacobi *j, const double *x);                /* out.init_vector :503 */
int fsim_jacobi_solve(fsim_jacobi *j, double tolerance, int32_t substep, int32_t max_iterations,
                      double *correlation, double *diff, int32_t *iterations, double *result); /* out.solve :576 */
int fsim_jacobi_get_result(fsim_jacobi *j, double *x);                       /* out.x_result_tex :703 */
int64_t fsim_jacobi_launch_count(const fsim_jacobi *j);
int fsim_jacobi_timing(fsim_jacobi *j, double *mv_ms, int64_t *mv_launches); /* device time of the mat-vec kernel */

#ifdef __cplusplus
}
#endif
#endif /* FUSIONSIM_H */
