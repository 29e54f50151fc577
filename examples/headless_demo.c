/* headless_demo.c -- the reference's page loop (public/javascripts/fusionsim.js:72-178) as a plain C
 * host of libfusionsim.so: the default demo scene, K frames of step() + density(), the canvas written
 * as a PPM.  No Python, no torch: only include/fusionsim.h.  Math.random() of the page is replaced by
 * a counter-based generator (splitmix64) so that a run can be reproduced.
 *
 *   make -C fusion_sim_b200/csrc demo
 *   fusion_sim_b200/csrc/headless_demo [frames] [canvas.ppm]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../include/fusionsim.h"

static double uniform01(uint64_t k) /* splitmix64 of the draw index -> [0,1) */
{
    uint64_t x = k + 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

#define CHECK(call)                                                          \
    do {                                                                     \
        if ((call) != FSIM_OK) {                                             \
            fprintf(stderr, "%s failed: %s\n", #call, fsim_last_error());    \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(int argc, char **argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 100;
    const char *ppm = argc > 2 ? argv[2] : NULL;
    /* fusionsim.js:74-83 */
    fsim_spec spec = {0};
    spec.radius = 1; spec.height = 2; spec.nr = 400; spec.nz = 800; spec.dt = 2e-9;
    spec.nparticles = 400; spec.particle_mass = 1.67e-27; spec.particle_charge = 1.602e-19;
    const int64_t nr = spec.nr, nz = spec.nz, n = spec.nparticles * spec.nparticles;
    fsim_sim *sim = NULL;
    CHECK(fsim_create(&spec, &sim));

    double *sink = malloc(sizeof(double) * nr * nz), *source = malloc(sizeof(double) * nr * nz);
    double *pos = malloc(sizeof(double) * 3 * n), *vel = malloc(sizeof(double) * 3 * n);
    uint8_t *canvas = malloc((size_t)4 * nr * nz);
    if (!sink || !source || !pos || !vel || !canvas) return 2;
    /* :94-122: sink[i][j], source[i][j] with i along r */
    for (int64_t i = 0; i < nr; ++i)
        for (int64_t j = 0; j < nz; ++j) { sink[i * nz + j] = 1.0; source[i * nz + j] = 0.0; }
    for (int64_t j = 0; j < nz; ++j) sink[(nr - 1) * nz + j] = 0;
    for (int64_t i = 1; i < nr - 1; ++i) { sink[i * nz] = 0; sink[i * nz + nz - 1] = 0; }
    for (int64_t i = 0; i < 50; ++i)
        for (int64_t j = 350; j < 450; ++j) source[i * nz + j] = 1.0;
    /* :125-128: six draws per particle, in this order */
    for (int64_t p = 0; p < n; ++p) {
        const uint64_t k = 6 * (uint64_t)p;
        pos[3 * p] = 0.2 * (uniform01(k) - 0.5);
        pos[3 * p + 1] = 0.2 * (uniform01(k + 1) - 0.5);
        pos[3 * p + 2] = 0.2 * (uniform01(k + 2) - 0.5) + 1;
        vel[3 * p] = 0.002 * (uniform01(k + 3) - 0.5);
        vel[3 * p + 1] = 0.002 * (uniform01(k + 4) - 0.5);
        vel[3 * p + 2] = 0.002 * (uniform01(k + 5) - 0.5);
    }
    /* :130-148 */
    CHECK(fsim_set_position(sim, pos));
    CHECK(fsim_set_velocity(sim, vel));
    CHECK(fsim_set_sink_mask(sim, sink));
    CHECK(fsim_set_source_pdf(sim, source, nr, nz));
    CHECK(fsim_add_current_loop(sim, 0.8, 2.0, -10000000));
    CHECK(fsim_add_current_loop(sim, 0.8, 0.0, 10000000));
    CHECK(fsim_precalc(sim));
    /* :170-178, the animation loop */
    for (int f = 0; f < frames; ++f) {
        CHECK(fsim_step(sim));
        CHECK(fsim_density(sim));
    }
    CHECK(fsim_render_rgba8(sim, canvas));
    uint64_t sum = 1469598103934665603ull; /* FNV-1a of the canvas bytes */
    for (int64_t k = 0; k < 4 * nr * nz; ++k) sum = (sum ^ canvas[k]) * 1099511628211ull;
    if (ppm) {
        FILE *fp = fopen(ppm, "wb");
        if (!fp) return 3;
        fprintf(fp, "P6\n%d %d\n255\n", (int)nr, (int)nz);
        for (int64_t k = 0; k < nr * nz; ++k) fwrite(canvas + 4 * k, 1, 3, fp);
        fclose(fp);
    }
    printf("{\"frames\": %d, \"particles\": %lld, \"launches\": %lld, \"canvas_fnv1a\": \"%016llx\"}\n", frames,
           (long long)fsim_particle_count(sim), (long long)fsim_launch_count(sim), (unsigned long long)sum);
    CHECK(fsim_destroy(sim));
    free(sink); free(source); free(pos); free(vel); free(canvas);
    return 0;
}
