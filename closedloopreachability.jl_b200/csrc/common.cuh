// common.cuh -- shared device/host definitions of libclr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/clr_b200.h"

#define CLR_DEV_MAX_LAYERS (CLR_MAX_LAYERS + 4)   // network layers + folded pre/post-processing layers
#define CLR_MAX_ROWS 512                          // widest layer supported (mask words below)
#define CLR_MASK_WORDS (CLR_MAX_ROWS / 32)
#define CLR_TILE_CELLS_MAX 64                     // cells per tile (smem metadata is sized for this)
#define CLR_NCOL_MAX 2048                         // columns per tile
#define CLR_HULL_MAX 64                           // output dims with a per-CTA hull accumulator

// ReachabilityBase.Comparison tolerances for Float64 (rtol = sqrt(eps), ztol = 10 eps, atol = 0)
#define CLR_ZTOL (10.0 * 2.220446049250313e-16)
#define CLR_RTOL 1.4901161193847656e-08

// One affine layer as the kernels see it.  Pre- and post-processing are folded in as extra
// Id layers by the host (see clr_api.cu: build_dev_net).
struct DevLayer {
    int n;            // input rows (K)
    int m;            // output rows (M)
    int ksteps;       // ceil(n / 4): DMMA k-steps
    int rblocks;      // ceil(m / 8): DMMA row blocks
    int pitch_in;     // column pitch (doubles) of the stores: DevNet::max_pitch for every layer (one pitch per network)
    int pitch_out;
    int act;          // CLR_ACT_*
    int collapse;     // 1: LazySets' one-row linear_map (single generator sum_j |W g_j|)
    int net_layer;    // index into clr_stats arrays, -1 for folded pre/post layers
    int remove_zero;  // 1: the reference calls remove_zero_columns after this layer (non-Id activation, projection)
    const double* Afrag;  // [rblocks][ksteps][32] DMMA A fragments of W (zero padded)
    const double* bias;   // [m]
    const double* Wcm;    // m x n column-major copy (pointwise kernels)
};

struct DevNet {
    int L;            // number of DevLayers
    int n_in;         // rows of the input sets
    int m_out;        // rows of the output sets
    int max_rows;     // max over layers of m and n_in
    int max_pitch;    // max over all stores
    int pad_[3];
    DevLayer layer[CLR_DEV_MAX_LAYERS];
};

// Device-side mirror of clr_stats plus scheduler state; lives in global memory, zeroed per call.
struct DevCallState {
    unsigned long long sum_p_in[CLR_MAX_LAYERS];
    unsigned long long sum_unstable[CLR_MAX_LAYERS];
    unsigned long long near_threshold;
    unsigned long long tiles;          // tiles processed (diagnostics)
    unsigned long long dropped;        // cells deferred because a tile overflowed
    int max_p_out;
    int error;                         // first device-side error (CLR_ERR_*)
    unsigned long long next_cell;      // tile scheduler cursor of the main pass
    unsigned long long next_retry;     // cursor of the retry pass (1 cell per tile, smem)
    unsigned long long next_big;       // cursor of the global-store pass
    unsigned long long n_retry;        // cells in the retry list
    unsigned long long n_big;          // cells in the big-cell list
    unsigned long long timers[8];      // DZ_TIMING builds: cycles per phase (input, product, appended, relax, layout, output, fetch, other)
};

__host__ __device__ inline int clr_pad4(int x) { return (x + 3) & ~3; }
// smallest pitch >= pad4(rows) with pitch % 16 in {4, 12}: conflict-free DMMA B-fragment loads
__host__ __device__ inline int clr_pitch(int rows) {
    int p = clr_pad4(rows < 1 ? 1 : rows);
    while ((p & 15) != 4 && (p & 15) != 12) p += 4;
    return p;
}
