/* latency_c.c -- small-call latency of the C ABI as a ccall / cgo binding sees it (no Python): blocking host-buffer calls.
 *   C1: ONE zonotope (n = 6, p = 14) through an ACC-shaped network 5 -> 20^5 -> 1 with the affine pre-processing (6 -> 5)
 *   C2: 10^3 box cells through a TORA-shaped network 4 -> 100^3 -> 1 (+ shift post-processing)
 * Random weights of the reference's shapes (latency does not depend on the values).
 * Build: gcc -O2 -I include profiles/latency_c.c -L closedloopreachability.jl_b200 -lclr_b200 -lm -o /tmp/latency_c */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "clr_b200.h"

static double rnd(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) / 16777216.0) * 2.0 - 1.0; }
static double now_us(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }

static clr_net* make_net(clr_ctx* ctx, int L, const int32_t* dims, const int32_t* act, unsigned seed) {
    double* W[16]; double* b[16];
    for (int l = 0; l < L; l++) {
        W[l] = malloc(sizeof(double) * dims[l] * dims[l + 1]); b[l] = malloc(sizeof(double) * dims[l + 1]);
        for (int i = 0; i < dims[l] * dims[l + 1]; i++) W[l][i] = rnd(&seed) * sqrt(3.0 / dims[l]);
        for (int i = 0; i < dims[l + 1]; i++) b[l][i] = 0.1 * rnd(&seed);
    }
    clr_net* net = NULL;
    if (clr_net_upload(ctx, L, dims, act, (const double* const*)W, (const double* const*)b, &net) != 0) { fprintf(stderr, "%s\n", clr_last_error(ctx)); exit(1); }
    for (int l = 0; l < L; l++) { free(W[l]); free(b[l]); }
    return net;
}

int main(void) {
    clr_ctx* ctx = NULL;
    if (clr_create(0, &ctx) != 0) { printf("clr_create failed: %s\n", clr_last_error(NULL)); return 3; }
    unsigned seed = 11;
    /* ---- C1 ---- */
    const int32_t d1[7] = {5, 20, 20, 20, 20, 20, 1};
    const int32_t a1[6] = {CLR_ACT_RELU, CLR_ACT_RELU, CLR_ACT_RELU, CLR_ACT_RELU, CLR_ACT_RELU, CLR_ACT_ID};
    clr_net* acc = make_net(ctx, 6, d1, a1, 3);
    double preA[30] = {0}, preb[5] = {30.0, 1.4, 0, 0, 0};          /* 5 x 6 column-major */
    preA[4 * 5 + 2] = 1.0; preA[0 * 5 + 3] = 1.0; preA[3 * 5 + 3] = -1.0; preA[1 * 5 + 4] = 1.0; preA[4 * 5 + 4] = -1.0;
    clr_prepost pp1; memset(&pp1, 0, sizeof pp1); pp1.pre_rows = 5; pp1.preA = preA; pp1.preb = preb;
    double C[6] = {90.3, 32.1, 0.2, 10.4, 30.2, 0.1}, G[6 * 14]; int64_t off[2] = {0, 14};
    for (int i = 0; i < 6 * 14; i++) G[i] = 0.05 * rnd(&seed);
    double lo1, hi1;
    for (int tiny = 1; tiny >= 0; tiny--) {
        clr_set_tiny(ctx, tiny);
        for (int i = 0; i < 50; i++) clr_deepz_zonotopes(ctx, acc, 6, 1, C, off, G, &pp1, &lo1, &hi1, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS);
        double t0 = now_us();
        const int reps = 2000;
        for (int i = 0; i < reps; i++)
            if (clr_deepz_zonotopes(ctx, acc, 6, 1, C, off, G, &pp1, &lo1, &hi1, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS) != 0) { fprintf(stderr, "%s\n", clr_last_error(ctx)); return 1; }
        printf("C1 ACC-shaped, ONE zonotope (n = 6, p = 14), %s kernel: %.1f us per propagation  [u in %.6f, %.6f]\n",
               tiny ? "one-CTA-per-cell" : "tiled DMMA", (now_us() - t0) / reps, lo1, hi1);
    }
    clr_set_tiny(ctx, 1);
    /* ---- C2 ---- */
    const int32_t d2[5] = {4, 100, 100, 100, 1};
    const int32_t a2[4] = {CLR_ACT_RELU, CLR_ACT_RELU, CLR_ACT_RELU, CLR_ACT_RELU};
    clr_net* tora = make_net(ctx, 4, d2, a2, 5);
    const double blo[4] = {0.6, -0.7, -0.4, 0.5}, bhi[4] = {0.7, -0.6, -0.3, 0.6};
    const int32_t part[4] = {10, 10, 10, 1};
    double centers[31], radii[31]; int t = 0;
    for (int d = 0; d < 4; d++) { const double r = (bhi[d] - blo[d]) / 2 / part[d]; for (int k = 0; k < part[d]; k++) { centers[t] = blo[d] + r + 2 * r * k; radii[t] = r; t++; } }
    double shift = -10.0; clr_prepost pp2; memset(&pp2, 0, sizeof pp2); pp2.post_kind = CLR_POST_SHIFT; pp2.postb = &shift;
    static double lo2[1000], hi2[1000];
    for (int i = 0; i < 50; i++) clr_deepz_boxgrid(ctx, tora, 4, part, centers, radii, 0, 1000, &pp2, lo2, hi2, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS);
    double t0 = now_us();
    const int reps2 = 1000;
    for (int i = 0; i < reps2; i++) clr_deepz_boxgrid(ctx, tora, 4, part, centers, radii, 0, 1000, &pp2, lo2, hi2, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS);
    const double us = (now_us() - t0) / reps2;
    printf("C2 TORA-shaped, 10^3 box cells per control step: %.1f us per step = %.2f M propagations/s\n", us, 1000.0 / us);
    clr_net_free(acc); clr_net_free(tora); clr_destroy(ctx);
    return 0;
}
