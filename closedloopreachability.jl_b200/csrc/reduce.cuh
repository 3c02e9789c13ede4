// reduce.cuh -- batched order reduction of the kept output zonotopes (Girard's method, GIR05).
//
// Replaces, for all cells of a control step at once, the per-cell
//     Z0 = _reduce_order(_convert_or_overapproximate(U0, Zonotope), 2; force_reduction=true)
// of TaylorModelReconstructor(box=false) (reference src/setops.jl:84-86), i.e. LazySets'
// reduce_order(Z, r, GIR05()): the generators are ordered by decreasing  |g|_1 - |g|_inf  (stable: ties keep
// their original order), the first  kept = m (r - 1)  are kept as they are and the others are replaced by their
// interval hull, a diagonal block whose entry i is the sum of |G[i, j]| over the boxed generators IN SORTED ORDER
// (the loop of _interval_hull).  The result has exactly m r generators: [G[:, kept ones] | diag(d)], the layout
// src/setops.jl:96-101 asserts (size (m, 2m), diagonal tail).  r = 1 boxes everything without sorting.
//
// One warp per cell (weights, bitonic sort of the indices, gather, row sums).  HBM-bound streaming kernel: 8 m (1 + p) bytes read and 8 m (1 + m r) bytes written per cell;
// the weights and the sorted order live in shared memory, the generators are re-read through L1.
#pragma once
#include "common.cuh"

struct ReduceArgs {
    long long N;
    int m, cap, order;        // rows, generator capacity of keepG, target order r
    int cap2;                 // cap rounded up to a power of two (length of a warp's index array)
    const int* keepN;         // N generator counts
    const double* keepC;      // m x N
    const double* keepG;      // m x cap x N
    double* outC;             // m x N
    double* outG;             // m x (m r) x N
    int* error;               // set to 1 when a cell has fewer than m (r - 1) generators
};

#define RD_WARPS 4

__global__ void __launch_bounds__(RD_WARPS * 32) reduce_order_kernel(ReduceArgs A) {
    extern __shared__ __align__(16) unsigned char rd_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* wgt = reinterpret_cast<double*>(rd_smem) + (size_t)warp * A.cap;
    int* sorted = reinterpret_cast<int*>(rd_smem + sizeof(double) * (size_t)RD_WARPS * A.cap) + (size_t)warp * A.cap2;
    const int m = A.m, kept = m * (A.order - 1), gout = m * A.order;
    for (long long cell = (long long)blockIdx.x * RD_WARPS + warp; cell < A.N; cell += (long long)gridDim.x * RD_WARPS) {
        const int p = A.keepN[cell];
        const double* G = A.keepG + (size_t)cell * A.cap * m;
        double* oG = A.outG + (size_t)cell * gout * m;
        for (int i = lane; i < m; i += 32) A.outC[cell * m + i] = A.keepC[cell * m + i];
        if (p < kept) {                      // the reference indexes past the end here (BoundsError)
            if (lane == 0) atomicExch(A.error, 1);
            continue;
        }
        if (kept > 0) {
            // weights: |g|_1 (summed over the rows in ascending order) minus |g|_inf
            for (int j = lane; j < p; j += 32) {
                const double* g = G + (size_t)j * m;
                double s = 0.0, mx = 0.0;
                for (int i = 0; i < m; i++) { const double a = fabs(g[i]); s += a; mx = fmax(mx, a); }
                wgt[j] = s - mx;
            }
            __syncwarp();
            // stable descending order: a bitonic sort of the generator indices under the total order
            //   a before b  <=>  w_a > w_b, or w_a == w_b and a < b          (ties keep their original order)
            // over the next power of two (the padding slots sort behind every generator).  O(p log^2 p / 32) steps per lane
            // instead of the O(p^2 / 32) of ranking by counting (65 % of the kernel's samples in round 1).
            int n2 = 1;
            while (n2 < p) n2 <<= 1;
            for (int j = lane; j < n2; j += 32) sorted[j] = j < p ? j : 0x7fffffff;
            __syncwarp();
            for (int k = 2; k <= n2; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = lane; t < (n2 >> 1); t += 32) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                        const int a = sorted[i], b = sorted[l];
                        const double wa = a < p ? wgt[a] : -1.0, wb = b < p ? wgt[b] : -1.0;     // weights are >= 0
                        const bool a_first = wa > wb || (wa == wb && a < b);
                        const bool up = (i & k) == 0;                 // this run is sorted "first things first", the next one reversed
                        if (a_first != up) { sorted[i] = b; sorted[l] = a; }
                    }
                    __syncwarp();
                }
            }
        }
        // kept generators, in sorted order
        for (int t = lane; t < kept * m; t += 32) {
            const int k = t / m, i = t - k * m;
            oG[t] = G[(size_t)sorted[k] * m + i];
        }
        // interval hull of the others: lane i owns row i; sequential sums in sorted order, 8 loads in flight
        for (int i = lane; i < m; i += 32) {
            double d = 0.0;
            for (int k0 = kept; k0 < p; k0 += 8) {
                double t[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int k = k0 + q;
                    t[q] = k < p ? G[(size_t)(kept > 0 ? sorted[k] : k) * m + i] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 8; q++) d += fabs(t[q]);       // + 0.0 past the end: exact
            }
            for (int i2 = 0; i2 < m; i2++) oG[(size_t)(kept + i2) * m + i] = 0.0;
            oG[(size_t)(kept + i) * m + i] = d;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// The same reduction for zonotopes of at most 32 E generators (E = 1, 2, 4, 8: every controller of the reference's models),
// restructured around what the profile of the kernel above showed (round 2: 64 % of the samples in the interval-hull loop --
// m of 32 lanes chasing gathers that miss L1 -- and 28 % in the shared-memory sort):
//   * a cell's generators are read from HBM ONCE, coalesced, into the warp's shared-memory slice;
//   * the (weight, index) pairs are sorted in REGISTERS: element e = 32 r + lane sits in register r of its lane, the bitonic
//     network exchanges through warp shuffles (distance < 32) or between registers (distance >= 32); the key is the weight's
//     bit pattern (weights are >= +0, so the integer order is the floating-point order), ties by index = the stable order;
//   * the generators to be boxed are gathered into sorted order by all 32 lanes first, so that the m sequential row sums
//     (the reference's summation order) read consecutive shared-memory words instead of chasing indices.
// Same results bit for bit as reduce_order_kernel (and as its CPU restatement in the tests).
// ---------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(RD_WARPS * 32) reduce_order_reg_kernel(ReduceArgs A) {
    extern __shared__ __align__(16) unsigned char rd_smem[];
    // warp-uniform values go through a lane-0 broadcast: ptxas then knows that the branches on them do not diverge and emits the
    // sort's shuffles without a WARPSYNC / convergence-barrier pair around each of them
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int m = A.m, kept = m * (A.order - 1), gout = m * A.order;
    const size_t slice = (size_t)A.cap * m;
    double* gs = reinterpret_cast<double*>(rd_smem) + (size_t)warp * 2 * slice;      // the cell's generators as stored
    double* hs = gs + slice;                                                         // |.| of the boxed ones, in sorted order
    int* sorted = reinterpret_cast<int*>(rd_smem + sizeof(double) * (size_t)RD_WARPS * 2 * slice) + (size_t)warp * 32 * E;
    for (long long cell = (long long)blockIdx.x * RD_WARPS + warp; cell < A.N; cell += (long long)gridDim.x * RD_WARPS) {
        const int p = __shfl_sync(0xffffffffu, A.keepN[cell], 0);
        const double* G = A.keepG + (size_t)cell * slice;
        double* oG = A.outG + (size_t)cell * gout * m;
        for (int i = lane; i < m; i += 32) A.outC[cell * m + i] = A.keepC[cell * m + i];
        if (p < kept) {                      // the reference indexes past the end here (BoundsError)
            if (lane == 0) atomicExch(A.error, 1);
            continue;
        }
        __syncwarp();                        // the previous cell's shared-memory reads are done
        for (int t = lane; t < p * m; t += 32) gs[t] = G[t];
        __syncwarp();
        if (kept > 0) {
            long long key[E];
            int idx[E];
#pragma unroll
            for (int r = 0; r < E; r++) {
                const int j = r * 32 + lane;
                idx[r] = j;
                key[r] = -1;                 // padding sorts behind every generator
                if (j < p) {
                    double s = 0.0, mx = 0.0;
                    for (int i = 0; i < m; i++) { const double a = fabs(gs[j * m + i]); s += a; mx = fmax(mx, a); }
                    key[r] = __double_as_longlong(s - mx);
                }
            }
            constexpr int n2 = 32 * E;
#pragma unroll
            for (int k = 2; k <= n2; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    if (j >= 32) {
                        const int rj = j >> 5;
#pragma unroll
                        for (int r = 0; r < E; r++) {
                            if ((r & rj) == 0) {
                                const int r2 = r | rj;
                                const bool up = (((r * 32 + lane) & k) == 0);
                                const bool a_first = key[r] > key[r2] || (key[r] == key[r2] && idx[r] < idx[r2]);
                                if (a_first != up) {
                                    const long long tk = key[r]; key[r] = key[r2]; key[r2] = tk;
                                    const int ti = idx[r]; idx[r] = idx[r2]; idx[r2] = ti;
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < E; r++) {
                            const long long ok = __shfl_xor_sync(0xffffffffu, key[r], j);
                            const int oi = __shfl_xor_sync(0xffffffffu, idx[r], j);
                            const bool up = (((r * 32 + lane) & k) == 0), lower = (lane & j) == 0;
                            const bool mine_first = key[r] > ok || (key[r] == ok && idx[r] < oi);
                            if (mine_first != (lower == up)) { key[r] = ok; idx[r] = oi; }
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < E; r++) sorted[r * 32 + lane] = idx[r];
            __syncwarp();
        }
        // kept generators, in sorted order
        for (int t = lane; t < kept * m; t += 32) {
            const int k = t / m, i = t - k * m;
            oG[t] = gs[sorted[k] * m + i];
        }
        // the others, gathered into sorted order by all lanes ...
        const int nb = p - kept;
        for (int t = lane; t < nb * m; t += 32) {
            const int k = t / m, i = t - k * m;
            hs[t] = fabs(gs[(kept > 0 ? sorted[kept + k] : k) * m + i]);
        }
        __syncwarp();
        // ... and their interval hull: lane i owns row i, sequential sums in sorted order (8 loads in flight)
        for (int i = lane; i < m; i += 32) {
            double d = 0.0;
            for (int k0 = 0; k0 < nb; k0 += 8) {
                double t[8];
#pragma unroll
                for (int q = 0; q < 8; q++) t[q] = k0 + q < nb ? hs[(k0 + q) * m + i] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) d += t[q];             // + 0.0 past the end: exact
            }
            for (int i2 = 0; i2 < m; i2++) oG[(size_t)(kept + i2) * m + i] = 0.0;
            oG[(size_t)(kept + i) * m + i] = d;
        }
    }
}
