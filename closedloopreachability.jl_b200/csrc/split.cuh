// split.cuh -- the reference's other splitters on the device.
//
//   apply(ZonotopeSplitter(gens, nparts), Z)  = LazySets split(Z, gens, nparts)      (reference src/splitters.jl:30-43)
//   apply(SignSplitter(), U::Interval)                                                (reference src/splitters.jl:60-71)
//
// split(Z, gens, nparts) [UPSTREAM-RECALL, LazySets AbstractZonotope.jl]: generator gens[i] is halved nparts[i] times; every
// halving replaces each zonotope of the list by the pair (centre - g/2, centre + g/2) with that generator halved, in this
// order, so H = sum(nparts) halvings give 2^H children per input zonotope; the first halving is the most significant bit of
// a child's index, the last the least significant (0: minus, 1: plus).  A child's centre is the input centre with the H
// offsets applied one after the other in the order of the halvings (the same floating-point additions as the host loop in
// clr_b200/splitters.py: bit-identical); all children share the generator matrix.
// Children are written in the padded layout clr_deepz_zonotopes_padded consumes, so a control period's cells can be split
// and propagated without leaving the device.
#pragma once
#include "common.cuh"

struct SplitArgs {
    long long N;                 // input zonotopes
    int n, cap, H, ng;           // rows, generator capacity of the padded layout, halvings, listed generators
    const double* C;             // [N][n]
    const double* G;             // [N][cap][n]
    const int* ncols;            // [N]
    const int* gens;             // [ng] 0-based
    const int* nparts;           // [ng]
    double* outC;                // [N << H][n]
    double* outG;                // [N << H][cap][n]
    int* outN;                   // [N << H]
    int* error;
};

// one warp per child: lanes over the rows for the centre, over the (column, row) entries for the generators
__global__ void split_zonotopes_kernel(SplitArgs A) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long total = A.N << A.H;
    if (warp >= total) return;
    const long long cell = warp >> A.H;
    const unsigned long long bits = (unsigned long long)(warp - (cell << A.H));
    const int n = A.n, p = A.ncols[cell];
    const double* Gi = A.G + (size_t)cell * A.cap * n;
    double* Go = A.outG + (size_t)warp * A.cap * n;
    for (int i = 0; i < A.ng; i++)
        if (A.gens[i] < 0 || A.gens[i] >= p) { if (lane == 0) atomicCAS(A.error, 0, (int)CLR_ERR_ARG); return; }
    // centre: the offsets of the H halvings, first halving first
    for (int r = lane; r < n; r += 32) {
        double c = A.C[cell * n + r];
        int h = 0;
        for (int i = 0; i < A.ng; i++) {
            double half = Gi[(size_t)A.gens[i] * n + r];
            for (int t = 0; t < A.nparts[i]; t++, h++) {
                half = half / 2;
                c = ((bits >> (A.H - 1 - h)) & 1ull) ? c + half : c - half;
            }
        }
        A.outC[warp * n + r] = c;
    }
    // generators: copied, the listed ones halved nparts times (a generator listed twice is halved by both entries)
    for (int e = lane; e < A.cap * n; e += 32) {
        const int j = e / n;
        double v = j < p ? Gi[e] : 0.0;
        if (j < p)
            for (int i = 0; i < A.ng; i++)
                if (A.gens[i] == j) for (int t = 0; t < A.nparts[i]; t++) v = v / 2;
        Go[e] = v;
    }
    if (lane == 0) A.outN[warp] = p;
}

// ---- SignSplitter: an interval with 0 in its interior becomes [lo, 0], [0, hi]; order kept (prefix sum over the counts) ----
#define SS_BLOCK 256
__global__ void sign_count_kernel(long long N, const double* lo, const double* hi, long long* block_sums) {
    __shared__ int s[SS_BLOCK / 32];
    const long long i = blockIdx.x * (long long)SS_BLOCK + threadIdx.x;
    const int c = i < N ? ((lo[i] < 0.0 && hi[i] > 0.0) ? 2 : 1) : 0;
    int w = c;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < SS_BLOCK / 32; k++) t += s[k]; block_sums[blockIdx.x] = t; }
}
// exclusive scan of the block sums by one block (the number of blocks is N / 256: a few thousand for 10^6 intervals)
__global__ void sign_scan_kernel(long long nblocks, long long* block_sums, long long* total) {
    __shared__ long long carry;
    __shared__ long long s[SS_BLOCK];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nblocks; base += SS_BLOCK) {
        const long long i = base + threadIdx.x;
        const long long v = i < nblocks ? block_sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < SS_BLOCK; d <<= 1) {
            const long long t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) block_sums[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += s[SS_BLOCK - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void sign_scatter_kernel(long long N, const double* lo, const double* hi, const long long* block_offs,
                                    double* out_lo, double* out_hi, long long* out_parent) {
    __shared__ int s[SS_BLOCK];
    const long long i = blockIdx.x * (long long)SS_BLOCK + threadIdx.x;
    const double l = i < N ? lo[i] : 0.0, h = i < N ? hi[i] : 0.0;
    const int c = i < N ? ((l < 0.0 && h > 0.0) ? 2 : 1) : 0;
    s[threadIdx.x] = c;
    __syncthreads();
    for (int d = 1; d < SS_BLOCK; d <<= 1) {
        const int t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i >= N) return;
    const long long o = block_offs[blockIdx.x] + s[threadIdx.x] - c;
    if (c == 2) {
        out_lo[o] = l; out_hi[o] = 0.0; out_parent[o] = i;
        out_lo[o + 1] = 0.0; out_hi[o + 1] = h; out_parent[o + 1] = i;
    } else {
        out_lo[o] = l; out_hi[o] = h; out_parent[o] = i;
    }
}
