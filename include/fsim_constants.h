/* fsim_constants.h -- numeric constants of the fusion-sim particle-step path.
 *
 * Every number here is a fact read out of the reference's shaders
 * (public/javascripts/empic.js); the file:line of each is cited.  The header is
 * included by the product (fusion_sim_b200/csrc) and by the test oracle
 * (oracle/), so both sides use the same literals.  Plain C, no dependencies.
 */
#ifndef FSIM_CONSTANTS_H
#define FSIM_CONSTANTS_H

#define FSIM_C_LIGHT          2.998e8          /* empic.js:27, :645 (m/s)                 */
#define FSIM_MU0              1.25663706e-6    /* empic.js:313, :404                      */
#define FSIM_PI_GLSL          3.14159265359    /* empic.js:313, :317, :404                */
#define FSIM_NQUAD            1000             /* empic.js:315 loop-quadrature points     */
#define FSIM_QUAD_SCALE       0.001            /* empic.js:313 (= 1/NQUAD, literal)       */
#define FSIM_N_ENTROPY        1024             /* empic.js:142 entropy table side         */
#define FSIM_N_INVCDF         512              /* empic.js:229-230 inverse-cdf table side */
#define FSIM_RNG_KEEP         0.999            /* empic.js:804                            */
#define FSIM_RNG_MIX          0.001            /* empic.js:804                            */
#define FSIM_RESPAWN_SPEED    0.001            /* empic.js:772                            */
#define FSIM_NSHAPE           11               /* empic.js:949 deposit footprint side     */
#define FSIM_SHAPE_MID        5                /* empic.js:952 (nshape-1)/2               */
#define FSIM_DEPOSIT_WEIGHT   0.001            /* empic.js:1006                           */
#define FSIM_NORM_SCALE       1000.0           /* empic.js:1056                           */
#define FSIM_NORM_HALF        0.5              /* empic.js:1056                           */
#define FSIM_EMA_RATIO        0.01             /* empic.js:1083                           */
#define FSIM_LOOP_FAR         2.0              /* empic.js:371 near/far table switch      */
#define FSIM_RENDER_DENSITY   0.5              /* empic.js:1105                           */

/* EXTENSION (no reference counterpart): self-consistent field solve */
#define FSIM_EPS0             8.8541878128e-12 /* vacuum permittivity (F/m), CODATA 2018  */
#define FSIM_PI               3.14159265358979323846

/* "next" row N3, second half: spindle-cusp boundary solve written from the intent of spindle.js
 * (it does not run in the reference); specification in include/fusionsim.h                       */
#define FSIM_SPINDLE_A        0.4              /* spindle.js:138 shape parameter of the plasma surface    */
#define FSIM_SPINDLE_NPOWER   3                /* spindle.js:64  makeSORIterative({n_power: 3})           */
#define FSIM_SPINDLE_L        256              /* 4 (2^n_power)^2 surface elements                        */
#define FSIM_SPINDLE_QW       0.00628318530718 /* 2 pi / 1000: midpoint-rule weight of the loop quadrature */
#define FSIM_SPINDLE_TOL      1e-9             /* solver tolerance (the reference's 1e-3 cannot converge)  */
#define FSIM_SPINDLE_SUBSTEP  64               /* mat-vecs between convergence checks                      */
#define FSIM_SPINDLE_MAXCHECK 4000             /* at most this many checks                                 */

/* per-cell record produced by precalc(): rows of the Boris matrix and the
 * half-kick constant, 12 reals per cell: R1.xyz R2.xyz R3.xyz A.xyz
 * (empic.js:499-502 keeps them in four RGBA textures).                        */
#define FSIM_CELLREC          12

#endif /* FSIM_CONSTANTS_H */
