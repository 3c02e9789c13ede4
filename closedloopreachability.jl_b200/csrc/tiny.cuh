// tiny.cuh -- DeepZ propagation for NARROW networks (every layer at most 32 neurons wide): one small CTA per cell.
//
// The reference's own CPU-runnable configurations (BASELINE configs[0]: ACC, 5 -> 20^5 -> 1; VerticalCAS; InvertedPendulum)
// propagate ONE or a few zonotopes per control period through a network whose layers are 20-25 neurons wide
// (src/solve.jl:153-166, :187-195 -> controller_forward, src/solve.jl:210-218).  The tiled DMMA kernel (deepz.cuh) is built
// for 100-neuron layers and thousands of cells: on one ACC zonotope its phases (product, barriers, relaxation, layout) are a
// latency chain of ~40 us for ~10^5 FLOPs.  Here a CTA of TN_WARPS warps owns a cell whose [centre | generators] store sits
// in shared memory (column-major, 32-row pitch: conflict free).  Lane r of EVERY warp owns row r and keeps its row of W in
// registers; the columns are dealt to the warps in groups of four, so the layer product is one broadcast LDS + one FP64
// multiply-add per (column, k) with no hazard between warps (a column is read and written by one warp), no tensor pipe
// (the matrices are 20 x 20) and no layout phase: appended generators are real columns behind the existing ones.
// The radius of a neuron is the sum over the warps (ascending) of each warp's sum over its own columns (ascending) -- a
// fixed order, whatever the batch.
// Same semantics as deepz_kernel (ReachabilityBase comparisons in the ReLU rule, sigmoid / tanh rule, LazySets' one-row
// linear_map, bounds c -/+ rad); results agree with deepz_kernel to rounding (~1e-16), not bit for bit.  Used for box
// outputs without statistics / kept zonotopes.
#pragma once
#include "deepz.cuh"

#define TN_WARPS 8
#define TN_ROWS 32

// radius of every row over columns 1 .. w-1: per-warp partial sums over the warp's own column groups, combined in warp order
__device__ __forceinline__ double tn_radius(const double* X, double* part, int w, int lane, int warp) {
    double p = 0.0;
    for (int j0 = warp * 4; j0 < w; j0 += TN_WARPS * 4)
#pragma unroll
        for (int q = 0; q < 4; q++) { const int j = j0 + q; if (j >= 1 && j < w) p += fabs(X[j * TN_ROWS + lane]); }
    part[warp * TN_ROWS + lane] = p;
    __syncthreads();
    double rad = 0.0;
#pragma unroll
    for (int q = 0; q < TN_WARPS; q++) rad += part[q * TN_ROWS + lane];
    return rad;
}

// cells: box grid (in.kind == 0) or explicit / padded zonotopes; capCols = columns the CTA's store holds
__global__ void __launch_bounds__(TN_WARPS * 32) deepz_tiny_kernel(const DevNet* __restrict__ net, DeepzInput in, DeepzOutput out,
                                                                   DevCallState* st, int capCols) {
    extern __shared__ __align__(16) double tn_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* X = tn_smem;                                    // [capCols][32]
    double* part = tn_smem + (size_t)capCols * TN_ROWS;     // [TN_WARPS][32]
    const int L = net->L, n = in.n, mOut = net->m_out;
    double hlo = INFINITY, hhi = -INFINITY;               // warp 0, lane r: row r's share of the hull
    for (long long cell = blockIdx.x; cell < in.N; cell += gridDim.x) {
        __syncthreads();                                    // the previous cell's store is free
        // ---- input zonotope -> store -------------------------------------------------------------------------------------
        int w;                                             // columns in use: centre + generators
        if (in.kind == 0) {
            const unsigned long long g = (unsigned long long)(in.cyc_blk > 0 ? in.cell_begin + (cell / in.cyc_blk) * in.cyc_stride + cell % in.cyc_blk
                                                                             : in.cell_begin + cell);
            double c = 0.0, r = 0.0;
            if (lane < n) {
                const unsigned long long sd = (unsigned long long)in.stride[lane];
                const int k = sd > g ? 0 : (int)((g / sd) % (unsigned long long)in.part[lane]);
                c = __ldg(in.centers + in.base[lane] + k);
                r = __ldg(in.radii + in.base[lane] + k);
            }
            const unsigned nz = __ballot_sync(0xffffffffu, lane < n && r != 0.0);      // flat dimensions have no generator
            const int p = __popc(nz);
            if (warp == 0) {
                X[lane] = c;
                for (int j = 0; j < p; j++) X[(1 + j) * TN_ROWS + lane] = 0.0;
                __syncwarp();
                if ((nz >> lane) & 1u) X[(1 + __popc(nz & ((1u << lane) - 1u))) * TN_ROWS + lane] = r;
            }
            w = 1 + p;
        } else {
            const int p = in.col_off ? (int)(in.col_off[cell + 1] - in.col_off[cell]) : in.ncols[cell];
            const long long go = in.col_off ? in.col_off[cell] : cell * (long long)in.pad_cap;
            if (warp == 0) X[lane] = lane < n ? in.C[cell * n + lane] : 0.0;
            for (int j = warp; j < p; j += TN_WARPS) X[(1 + j) * TN_ROWS + lane] = lane < n ? in.G[(go + j) * n + lane] : 0.0;
            w = 1 + p;
        }
        __syncthreads();
        bool overflow = false;
        // ---- layers ---------------------------------------------------------------------------------------------------------
        for (int l = 0; l < L; l++) {
            const DevLayer& Ly = net->layer[l];
            const int m = Ly.m, K = Ly.n, act = Ly.act;
            const bool collapse = Ly.collapse != 0;
            double wr[TN_ROWS];                            // this lane's row of W
#pragma unroll
            for (int k = 0; k < TN_ROWS; k++) wr[k] = (k < K && lane < m) ? __ldg(Ly.Wcm + (size_t)k * m + lane) : 0.0;
            const double bias = lane < m ? __ldg(Ly.bias + lane) : 0.0;
            // product in place: this warp's column groups; all lanes read the four columns (broadcast), then write their rows
            for (int j0 = warp * 4; j0 < w; j0 += TN_WARPS * 4) {
                double y[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < TN_ROWS; k++) {
                    if (k < K) {
#pragma unroll
                        for (int q = 0; q < 4; q++) y[q] += wr[k] * X[(min(j0 + q, w - 1)) * TN_ROWS + k];
                    }
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; q++) if (j0 + q < w) X[(j0 + q) * TN_ROWS + lane] = (j0 + q == 0) ? y[q] + bias : y[q];
            }
            __syncthreads();
            // relaxation: every warp computes the rule of every row (same inputs, same result) and scales its own columns
            if (!(act == CLR_ACT_ID && !collapse)) {
                const double rad = tn_radius(X, part, w, lane, warp);
                const double c = X[lane];
                double cn = c, la = 1.0, mm = 0.0;
                const bool unst = lane < m && dz_rule<true>(act, c, c - rad, c + rad, cn, la, mm);
                const unsigned um = __ballot_sync(0xffffffffu, unst);
                const int u = __popc(um);
                const int wNew = collapse ? 2 : w;
                __syncthreads();                            // every warp has read the centre column and the partial sums
                if (wNew + u > capCols) { overflow = true; break; }
                if (lane < m) {
                    if (collapse) { if (warp == 0) { X[lane] = cn; X[TN_ROWS + lane] = la * rad; } }
                    else {
                        if (warp == 0) X[lane] = cn;
                        if (la != 1.0)
                            for (int j0 = warp * 4; j0 < w; j0 += TN_WARPS * 4)
#pragma unroll
                                for (int q = 0; q < 4; q++) { const int j = j0 + q; if (j >= 1 && j < w) X[j * TN_ROWS + lane] = la == 0.0 ? 0.0 : X[j * TN_ROWS + lane] * la; }
                    }
                }
                // appended generators mu_i e_i, ascending i: column wNew + rank(i), dealt to the warps
                const int rank = __popc(um & ((1u << lane) - 1u));
                for (int t = warp; t < u; t += TN_WARPS) X[(wNew + t) * TN_ROWS + lane] = (unst && rank == t) ? mm : 0.0;
                w = wNew + u;
                __syncthreads();
            }
        }
        // (a plain store: every writer stores the same value, and on the zero-copy path of small calls the word lives in mapped host memory)
        if (overflow) { if (threadIdx.x == 0) { *(volatile int*)&st->error = (int)CLR_ERR_CAPACITY; __threadfence_system(); } continue; }
        // ---- output box --------------------------------------------------------------------------------------------------------
        const double rad = tn_radius(X, part, w, lane, warp);
        if (warp == 0 && lane < mOut) {
            const double c = X[lane];
            const double lo = c - rad, hi = c + rad;
            if (out.out_lo) { out.out_lo[cell * mOut + lane] = lo; out.out_hi[cell * mOut + lane] = hi; }
            hlo = fmin(hlo, lo); hhi = fmax(hhi, hi);
        }
    }
    if (out.hull_lo && warp == 0 && lane < mOut && hlo <= hhi) { atomicMinD(&out.hull_lo[lane], hlo); atomicMaxD(&out.hull_hi[lane], hhi); }
}
