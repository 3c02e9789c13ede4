/* abi_check.c -- the C ABI of libclr_b200.so driven from plain C (no Python, no torch): what a cgo / ccall / JNI binding sees.
 *
 * Without a CUDA device clr_create must fail with CLR_ERR_CUDA and a message (there is no CPU fallback): exit code 3.
 * With a device: TORA-shaped random ReLU network (4 -> 20 -> 20 -> 1), BoxSplitter([4, 4, 3, 5]) of the TORA X0
 * (models/TORA/TORA.jl:98, :188), every cell's output box, the hull, and three structural checks that need no oracle:
 *   (1) the hull equals the min / max of the cell boxes; (2) two half shards concatenate to the full result bit for bit;
 *   (3) a degenerate (zero-radius) cell returns lo == hi == the pointwise forward of its centre (clr_forward_points).
 * Exit code 0 = all good.  Build: gcc -I include tests/c_abi/abi_check.c -L <pkg> -lclr_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "clr_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int32_t rc_ = (call);                                                            \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, clr_last_error(ctx)); return 1; } \
    } while (0)

static double rnd(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) / 16777216.0) * 2.0 - 1.0; }

int main(void) {
    clr_ctx* ctx = NULL;
    int32_t rc = clr_create(0, &ctx);
    if (rc != 0) {
        const char* msg = clr_last_error(NULL);
        printf("clr_create failed (%d): %s\n", rc, msg ? msg : "(no message)");
        return (rc == CLR_ERR_CUDA && msg && msg[0]) ? 3 : 4;
    }
    const int32_t dims[4] = {4, 20, 20, 1};
    const int32_t act[3] = {CLR_ACT_RELU, CLR_ACT_RELU, CLR_ACT_ID};
    double *W[3], *b[3];
    unsigned seed = 7;
    for (int l = 0; l < 3; l++) {
        W[l] = malloc(sizeof(double) * dims[l] * dims[l + 1]);
        b[l] = malloc(sizeof(double) * dims[l + 1]);
        for (int i = 0; i < dims[l] * dims[l + 1]; i++) W[l][i] = rnd(&seed) / sqrt((double)dims[l]);
        for (int i = 0; i < dims[l + 1]; i++) b[l][i] = 0.1 * rnd(&seed);
    }
    clr_net* net = NULL;
    CHECK(clr_net_upload(ctx, 3, dims, act, (const double* const*)W, (const double* const*)b, &net));

    const double lo[4] = {0.6, -0.7, -0.4, 0.5}, hi[4] = {0.7, -0.6, -0.3, 0.6};
    const int32_t part[4] = {4, 4, 3, 5};
    double centers[16], radii[16];
    int t = 0;
    for (int d = 0; d < 4; d++) {
        const double r = (hi[d] - lo[d]) / 2 / part[d];
        for (int k = 0; k < part[d]; k++) { centers[t] = lo[d] + r + 2 * r * k; radii[t] = r; t++; }
    }
    const int64_t N = 240;
    double out_lo[240], out_hi[240], hull_lo, hull_hi, a_lo[240], a_hi[240];
    clr_prepost pp;
    memset(&pp, 0, sizeof pp);
    double shift = -10.0;
    pp.post_kind = CLR_POST_SHIFT; pp.postb = &shift;
    clr_stats st;
    CHECK(clr_deepz_boxgrid(ctx, net, 4, part, centers, radii, 0, N, &pp, out_lo, out_hi, &hull_lo, &hull_hi, &st, 0, 0));
    double mn = INFINITY, mx = -INFINITY;
    for (int i = 0; i < N; i++) {
        if (!(out_lo[i] <= out_hi[i])) { fprintf(stderr, "cell %d: lo > hi\n", i); return 1; }
        mn = fmin(mn, out_lo[i]); mx = fmax(mx, out_hi[i]);
    }
    if (mn != hull_lo || mx != hull_hi) { fprintf(stderr, "hull mismatch\n"); return 1; }
    if (st.n_cells != N || st.sum_p_in[0] != 4 * N) { fprintf(stderr, "stats mismatch\n"); return 1; }
    /* shards reproduce the full run bit for bit (like for like: a run without statistics of a network this narrow takes the
       one-CTA-per-cell kernel, whose row sums are ordered differently from the tiled kernel's -- equal to rounding only) */
    double f_lo[240], f_hi[240];
    CHECK(clr_deepz_boxgrid(ctx, net, 4, part, centers, radii, 0, N, &pp, f_lo, f_hi, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS));
    for (int i = 0; i < N; i++)
        if (fabs(f_lo[i] - out_lo[i]) > 1e-12 || fabs(f_hi[i] - out_hi[i]) > 1e-12) { fprintf(stderr, "cell %d: run without statistics differs\n", i); return 1; }
    CHECK(clr_deepz_boxgrid(ctx, net, 4, part, centers, radii, 0, 100, &pp, a_lo, a_hi, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS));
    CHECK(clr_deepz_boxgrid(ctx, net, 4, part, centers, radii, 100, 140, &pp, a_lo + 100, a_hi + 100, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS));
    if (memcmp(a_lo, f_lo, sizeof f_lo) || memcmp(a_hi, f_hi, sizeof f_hi)) { fprintf(stderr, "shards differ from the full run\n"); return 1; }
    /* degenerate cell == pointwise forward */
    const int32_t one[4] = {1, 1, 1, 1};
    const double c1[4] = {0.65, -0.65, -0.35, 0.55}, r0[4] = {0, 0, 0, 0};
    double p_lo, p_hi, u;
    CHECK(clr_deepz_boxgrid(ctx, net, 4, one, c1, r0, 0, 1, &pp, &p_lo, &p_hi, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS));
    CHECK(clr_forward_points(ctx, net, &pp, 4, 1, c1, &u, 0));
    if (p_lo != p_hi || fabs(p_lo - u) > 1e-12) { fprintf(stderr, "degenerate cell: [%g, %g] vs point %g\n", p_lo, p_hi, u); return 1; }
    /* error path: a partition entry < 1 is an argument error with a message */
    const int32_t bad[4] = {4, 0, 3, 5};
    rc = clr_deepz_boxgrid(ctx, net, 4, bad, centers, radii, 0, 1, &pp, out_lo, out_hi, NULL, NULL, NULL, 0, CLR_FLAG_NO_STATS);
    if (rc != CLR_ERR_ARG || !clr_last_error(ctx)[0]) { fprintf(stderr, "bad partition not rejected (%d)\n", rc); return 1; }
    /* (4) two shards behind one handle, without Python / torch / NCCL (clr_create_multi; device 0 twice on a one-GPU box): the
       boxes in cell order and the peer-gathered hull equal the single-context run bit for bit */
    {
        const int32_t devs[2] = {0, 0};
        clr_multi* mc = NULL; clr_multi_net* mn = NULL;
        if (clr_create_multi(devs, 2, &mc) != 0) { fprintf(stderr, "clr_create_multi: %s\n", clr_last_error(NULL)); return 1; }
        if (clr_multi_size(mc) != 2 || !clr_multi_ctx(mc, 1)) { fprintf(stderr, "multi handle malformed\n"); return 1; }
        if (clr_multi_net_upload(mc, 3, dims, act, (const double* const*)W, (const double* const*)b, &mn) != 0) {
            fprintf(stderr, "clr_multi_net_upload: %s\n", clr_multi_last_error(mc)); return 1;
        }
        double m_lo[240], m_hi[240], mh_lo, mh_hi, ms[2]; int64_t bounds[3];
        for (int mode = 0; mode < 3; mode += 2) {       /* contiguous and block-cyclic shards */
            if (clr_multi_deepz_boxgrid(mc, mn, 4, part, centers, radii, 0, N, &pp, m_lo, m_hi, &mh_lo, &mh_hi, &st, mode, bounds, ms) != 0) {
                fprintf(stderr, "clr_multi_deepz_boxgrid: %s\n", clr_multi_last_error(mc)); return 1;
            }
            if (memcmp(m_lo, out_lo, sizeof out_lo) || memcmp(m_hi, out_hi, sizeof out_hi) || mh_lo != hull_lo || mh_hi != hull_hi ||
                st.n_cells != N || bounds[0] != 0 || bounds[2] != N) { fprintf(stderr, "two shards differ from one context (mode %d)\n", mode); return 1; }
        }
        clr_multi_net_free(mn);
        clr_destroy_multi(mc);
    }
    printf("abi_check ok: 240 cells, hull [%.6f, %.6f], %lld launches\n", hull_lo, hull_hi, (long long)clr_launch_count(ctx));
    clr_net_free(net);
    clr_destroy(ctx);
    for (int l = 0; l < 3; l++) { free(W[l]); free(b[l]); }
    return 0;
}
