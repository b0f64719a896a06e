/* Minimal plain-C host of libmpm_b200.so: what a maintainer's P/Invoke layer does, without the managed runtime.
 * Builds with any C99 compiler against include/mpm_b200.h:
 *     gcc -std=c99 -Iinclude mls-mpm-godot_b200/host/example.c -Lmls-mpm-godot_b200 -lmpm_b200 -Wl,-rpath,$PWD/mls-mpm-godot_b200 -o example
 * Without a GPU it prints the ABI version and the shipping scene's parameters and stops (mpm_create fails: there is no CPU
 * fallback); with one it runs the reference's shipping scene (64^3 grid, 54^3 particles, MLSMPM3DFluidMultithreadGPU.cs:54-84)
 * for a few frames the way _Process does (H:234-251): set_sphere, step(sim_iterations), positions. */
#include <stdio.h>
#include <stdlib.h>

#include "mpm_b200.h"

int main(void)
{
    MpmParams p;
    MpmSolver* s = NULL;
    int rc;
    printf("abi %d, devices %d\n", mpm_abi_version(), mpm_device_count());
    rc = mpm_default_params(MPM_VARIANT_3D_GPU, &p);
    if (rc != MPM_OK) { printf("default params: %d\n", rc); return 1; }
    printf("grid %d x %d x %d, dt %g, rest density %g, eos stiffness %g\n", p.grid_size[0], p.grid_size[1], p.grid_size[2], p.dt,
           p.rest_density, p.eos_stiffness);
    if (mpm_device_count() < 1) { printf("no CUDA device: nothing to run (the library has no CPU path)\n"); return 0; }
    p.math_mode = MPM_MATH_FAST; /* the benchmarked path (cell kernels); the default is the bit-exact strict path */
    rc = mpm_create(&p, 200000, 0, &s);
    if (rc != MPM_OK) { printf("mpm_create: %d %s\n", rc, mpm_last_error(NULL)); return 1; }
    {
        const float lo[3] = {4.0f, 4.0f, 4.0f}, hi[3] = {31.0f, 31.0f, 31.0f};
        int64_t n = 0;
        float* pos;
        int frame;
        rc = mpm_init_block(s, lo, hi, 0.5f); /* InitialiseSim (H:654-707): 54^3 = 157 464 particles */
        if (rc == MPM_OK) rc = mpm_num_particles(s, &n);
        if (rc != MPM_OK) { printf("init: %d %s\n", rc, mpm_last_error(s)); mpm_destroy(s); return 1; }
        pos = (float*)malloc(sizeof(float) * 4 * (size_t)n);
        for (frame = 0; frame < 3 && rc == MPM_OK; ++frame) {
            const float sphere[3] = {-21.648403f + 0.5f * (float)frame, 0.0f, 31.707275f};
            rc = mpm_set_sphere(s, sphere);
            if (rc == MPM_OK) rc = mpm_step(s, 2); /* sim_iterations = 2 */
            if (rc == MPM_OK) rc = mpm_get_positions(s, pos, n, NULL, NULL);
            if (rc == MPM_OK) printf("frame %d: %lld particles, first at (%g, %g, %g), |v| %g\n", frame, (long long)n, pos[0], pos[1], pos[2], pos[3]);
        }
        if (rc != MPM_OK) printf("step: %d %s\n", rc, mpm_last_error(s));
        free(pos);
    }
    mpm_destroy(s);
    return rc == MPM_OK ? 0 : 1;
}
