/* minimal.c — the C ABI of include/fishgym.h from plain C (no Python, no C++).
 *
 *   cc -std=c99 -Iinclude examples/minimal.c -o minimal -Lgym-fish_b200/csrc -lfishgym_cuda -Wl,-rpath,$PWD/gym-fish_b200/csrc
 *
 * Creates a small periodic box with a uniform body force, steps it, and checks that the momentum gained is
 * steps * force (Guo forcing adds exactly F per step).  The same source links against oracle/libfishgym_oracle.so,
 * which is how tests/test_abi.py runs it on a machine without a GPU. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fishgym.h"

int main(void) {
    FgConfig cfg;
    FgSim *sim = NULL;
    FgStats st;
    const int n = 16, steps = 10;
    size_t cells = (size_t)n * n * n, i;
    float *rho, *u;
    double jz = 0.0;
    int rc;

    if (fg_config_default(&cfg) != FG_OK) return 2;
    cfg.nx = cfg.ny = cfg.nz = n;
    cfg.collision = FG_MRT;
    cfg.tau = 0.7;
    cfg.body_force[2] = 1e-4;
    rc = fg_create(&cfg, &sim);
    if (rc != FG_OK) {
        fprintf(stderr, "fg_create failed (%d): %s\n", rc, fg_last_error(NULL));
        return 1;
    }
    printf("backend %s, ABI %d\n", fg_backend_name(), fg_abi_version());
    rc = fg_step(sim, steps);
    if (rc != FG_OK) { fprintf(stderr, "fg_step: %s\n", fg_last_error(sim)); return 1; }
    rho = (float *)malloc(cells * sizeof(float));
    u = (float *)malloc(3 * cells * sizeof(float));
    if (!rho || !u) return 2;
    rc = fg_get_fields(sim, rho, u);
    if (rc != FG_OK) { fprintf(stderr, "fg_get_fields: %s\n", fg_last_error(sim)); return 1; }
    for (i = 0; i < cells; ++i) jz += (double)rho[i] * u[2 * cells + i];
    jz /= (double)cells;
    fg_get_stats(sim, &st);
    printf("steps %lld, mean rho*u_z %.6e (expected %.6e)\n", (long long)st.steps, jz, steps * cfg.body_force[2]);
    free(rho);
    free(u);
    fg_destroy(sim);
    return fabs(jz - steps * cfg.body_force[2]) < 1e-8 ? 0 : 3;
}
