// deepz.cuh -- batched Float64 DeepZ propagation of zonotope cells through a dense network.
//
// Replaces the per-cell loop of the reference (src/solve.jl:153-166, :187-195) around
// controller_forward (src/solve.jl:210-218) with one persistent kernel:
//
//   * a CTA owns a TILE of cells; the tile's zonotopes ([centre | generators] per cell, all cells
//     side by side) live in shared memory as one column store (column-major, conflict-free pitch);
//   * per layer, W * [c | G] for all columns of the tile runs on the FP64 tensor pipe
//     (mma.sync.m8n8k4.f64 -> SASS DMMA; tcgen05 has no FP64 kind): a warp owns 8 output rows,
//     keeps their W fragments in registers and streams the columns as B fragments from smem;
//   * the interval bounds (sequential |.| row sums, same order as LazySets low/high), the
//     ReLU / sigmoid / tanh relaxation, the row scaling and the append of the new generators
//     are fused in the same kernel and never leave the SM;
//   * the output box (src/solve.jl:214-216, src/setops.jl:76-83), the hull over cells and the
//     optional output zonotopes are written straight from smem.
//
// Tiles are handed out by an atomic cursor; the tile size adapts to the generator growth seen so
// far; cells that overflow a tile are retried alone and, if a single cell does not fit in
// shared memory, in a pass whose column store lives in a global (L2-resident) workspace.
#pragma once
#include "deepz_types.cuh"


__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ bool dz_approxzero(double x) { return fabs(x) <= CLR_ZTOL; }
__device__ __forceinline__ bool dz_approx(double x, double y) {
    if (dz_approxzero(x) && dz_approxzero(y)) return true;
    return x == y || fabs(x - y) <= CLR_RTOL * fmax(fabs(x), fabs(y));
}
__device__ __forceinline__ bool dz_leq(double x, double y) { return x <= y || dz_approx(x, y); }

__device__ __forceinline__ void atomicMinD(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}
__device__ __forceinline__ void atomicMaxD(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v) {
        unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}

// ---------------------------------------------------------------------------------------------
// Warp -> (row block, column part) mapping of the layer product.  16 warps sit on 4 SM sub-partitions
// (warp w on w % 4); the DMMA rate is per sub-partition, so the row blocks are dealt such that the
// sub-partitions get (nearly) the same number of DMMAs: with 13 row blocks (100 neurons) warps 0..11
// own one row block each and warps 12..15 share the last one, a quarter of the columns each.
// ---------------------------------------------------------------------------------------------
struct DzMap { int rb0, rbStep, part, nParts; };
__device__ __forceinline__ DzMap dz_map(int RB, int warp, int nWarps) {
    DzMap m;
    if (RB >= nWarps) { m.rb0 = warp; m.rbStep = nWarps; m.part = 0; m.nParts = 1; }
    else if (nWarps >= 2 * RB) {
        m.nParts = nWarps / RB; m.rb0 = warp % RB; m.part = warp / RB; m.rbStep = RB;
        if (m.part >= m.nParts) m.rb0 = RB;      // idle warp
    } else if (warp < RB - 1) { m.rb0 = warp; m.rbStep = RB; m.part = 0; m.nParts = 1; }
    else { m.rb0 = RB - 1; m.rbStep = RB; m.part = warp - (RB - 1); m.nParts = nWarps - (RB - 1); }
    return m;
}

// A fragments of this warp's first row block of layer L (kept in registers across the relaxation
// of the previous layer, so the L2 latency of the weights is off the critical path)
template <int KSREG>
__device__ __forceinline__ void dz_load_a(const DevLayer& L, int warp, int lane, int nWarps, double (&a)[KSREG]) {
    const DzMap mp = dz_map(L.rblocks, warp, nWarps);
    if (L.ksteps > KSREG || mp.rb0 >= L.rblocks) return;
    const double* Af = L.Afrag + (size_t)mp.rb0 * L.ksteps * 32 + lane;
#pragma unroll
    for (int k = 0; k < KSREG; k++) a[k] = k < L.ksteps ? __ldg(Af + k * 32) : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Layer product: Y[:, j] = W * X[:, j] for all nCols columns (bias is added by the relaxation, or
// here on the centre columns for Id layers).  `a` holds the fragments of the warp's first row block.
// ---------------------------------------------------------------------------------------------
template <int KSREG>
__device__ __forceinline__ void dz_gemm(const DevLayer& L, const double* __restrict__ X, double* __restrict__ Y,
                                        int nCols, const unsigned* centreMask, bool addBias, int warp, int lane,
                                        int nWarps, double (&a)[KSREG]) {
    constexpr int NB = 4;                       // column blocks (8 columns each) per group
    const int RB = L.rblocks, KS = L.ksteps;
    const int CG = (nCols + 8 * NB - 1) / (8 * NB);
    const int pin = L.pitch_in, pout = L.pitch_out;
    const int rowsW = clr_pad4(L.m);            // rows written (pad rows of W are zero)
    const int qr = lane >> 2, qc = lane & 3;
    const DzMap mp = dz_map(RB, warp, nWarps);
    for (int rb = mp.rb0; rb < RB; rb += mp.rbStep) {
        const double* Af = L.Afrag + (size_t)rb * KS * 32 + lane;
        const int row = rb * 8 + qr;
        const double bias = (addBias && row < L.m) ? __ldg(L.bias + row) : 0.0;
        const bool resident = KS <= KSREG;
        if (resident && rb != mp.rb0) {
#pragma unroll
            for (int k = 0; k < KSREG; k++) a[k] = k < KS ? __ldg(Af + k * 32) : 0.0;
        }
        for (int cg = mp.part; cg < CG; cg += mp.nParts) {
            double d[NB][2];
#pragma unroll
            for (int j = 0; j < NB; j++) { d[j][0] = 0.0; d[j][1] = 0.0; }
            const double* xb = X + (cg * NB * 8 + qr) * pin + qc;
            const int nbAct = min(NB, (nCols - cg * NB * 8 + 7) >> 3);   // column blocks that hold real columns
            if (resident) {
#pragma unroll
                for (int k = 0; k < KSREG; k++) {
                    if (k < KS) {
#pragma unroll
                        for (int j = 0; j < NB; j++) {
                            if (j < nbAct) {
                                double b = xb[j * 8 * pin + k * 4];
                                dmma884(d[j][0], d[j][1], a[k], b);
                            }
                        }
                    }
                }
            } else {
                // wide-K layers: stream the A fragments
                for (int k = 0; k < KS; k++) {
                    double av = __ldg(Af + k * 32);
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        if (j < nbAct) {
                            double b = xb[j * 8 * pin + k * 4];
                            dmma884(d[j][0], d[j][1], av, b);
                        }
                    }
                }
            }
            if (row < rowsW) {
#pragma unroll
                for (int j = 0; j < NB; j++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        int col = cg * NB * 8 + j * 8 + qc * 2 + e;
                        if (col < nCols) {
                            double v = d[j][e];
                            if (addBias && ((centreMask[col >> 5] >> (col & 31)) & 1u)) v += bias;
                            Y[col * pout + row] = v;
                        }
                    }
                }
            }
        }
    }
}

// remove_zero_columns bookkeeping: the reference drops all-zero generators after every non-Id
// activation (and after a projection); columns are never physically removed here (a zero column
// stays zero and contributes nothing), they are only marked so that generator counts, statistics
// and fetched zonotopes agree with the reference.
__device__ __forceinline__ void dz_mark_dead(DeepzSmem& S, const double* X, int pitch, int rows, int nCols,
                                             int tid, int NT) {
    for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) S.deadMask[i] = 0;
    __syncthreads();
    for (int col = tid; col < nCols; col += NT) {
        if ((S.centreMask[col >> 5] >> (col & 31)) & 1u) continue;
        bool any = false;
        for (int r = 0; r < rows && !any; r++) any = X[(size_t)col * pitch + r] != 0.0;
        if (!any) atomicOr(&S.deadMask[col >> 5], 1u << (col & 31));
    }
    __syncthreads();
}

__device__ __forceinline__ void dz_build_centre_mask(DeepzSmem& S, const int* colOff, int nCells, int tid, int NT) {
    for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) S.centreMask[i] = 0;
    __syncthreads();
    if (tid < nCells) {
        int cb = colOff[tid];
        atomicOr(&S.centreMask[cb >> 5], 1u << (cb & 31));
    }
    __syncthreads();
}

__device__ __forceinline__ double dz_sigm(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double dz_sigm_p(double x) { double ex = exp(x); return ex / ((1.0 + ex) * (1.0 + ex)); }
__device__ __forceinline__ double dz_tanh_p(double x) { double t = tanh(x); return 1.0 - t * t; }

// Column layout of a tile from per-cell widths, computed redundantly by every warp (no serial
// section, no barrier): lane i holds the widths of cells i and i + 32 and gets their offsets back;
// `keep` = number of leading cells that fit into `cap` doubles at pitch `pm`, `total` their columns.
__device__ __forceinline__ void dz_layout(int w0, int w1, int nCells, int pm, int cap, int lane, int& off0, int& off1,
                                          int& keep, int& total) {
    int s0 = w0, s1 = w1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t0 = __shfl_up_sync(0xffffffffu, s0, d);
        int t1 = __shfl_up_sync(0xffffffffu, s1, d);
        if (lane >= d) { s0 += t0; s1 += t1; }
    }
    s1 += __shfl_sync(0xffffffffu, s0, 31);
    const bool f0 = lane < nCells && (long long)(s0 + 8) * pm <= cap && s0 <= CLR_NCOL_MAX;
    const bool f1 = lane + 32 < nCells && (long long)(s1 + 8) * pm <= cap && s1 <= CLR_NCOL_MAX;
    keep = __popc(__ballot_sync(0xffffffffu, f0)) + __popc(__ballot_sync(0xffffffffu, f1));
    off0 = s0 - w0; off1 = s1 - w1;
    const int e0 = __shfl_sync(0xffffffffu, s0, (keep - 1) & 31);
    const int e1 = __shfl_sync(0xffffffffu, s1, (keep - 33) & 31);
    total = keep == 0 ? 0 : (keep <= 32 ? e0 : e1);
}

// cells keep..nCells-1 of the tile did not fit: hand them to the next pass (one thread)
__device__ __forceinline__ void dz_defer(const DeepzSched& sp, DevCallState* st, const long long* cellId, int keep, int nCells) {
    const int nd = nCells - keep;
    if (sp.pass == 0 && keep > 0) {
        unsigned long long o = atomicAdd(&st->n_retry, (unsigned long long)nd);
        for (int i = 0; i < nd; i++) sp.retry_list[o + i] = cellId[keep + i];
    } else if (sp.pass == 0) {
        unsigned long long o = atomicAdd(&st->n_big, 1ull);
        sp.big_list[o] = cellId[0];
        if (nd > 1) {
            unsigned long long o2 = atomicAdd(&st->n_retry, (unsigned long long)(nd - 1));
            for (int i = 1; i < nd; i++) sp.retry_list[o2 + i - 1] = cellId[i];
        }
    } else if (sp.pass == 1) {
        unsigned long long o = atomicAdd(&st->n_big, (unsigned long long)nd);
        for (int i = 0; i < nd; i++) sp.big_list[o + i] = cellId[keep + i];
    } else {
        atomicCAS(&st->error, 0, (int)CLR_ERR_CAPACITY);
    }
    atomicAdd(&st->dropped, (unsigned long long)nd);
}

// ---------------------------------------------------------------------------------------------
// The persistent tile kernel.
// ---------------------------------------------------------------------------------------------
template <int KSREG, bool GLOBAL, bool STATS>
__global__ void __launch_bounds__(DZ_THREADS, 1)
deepz_kernel(const DevNet* __restrict__ net, DeepzInput in, DeepzOutput out, DevCallState* st, DeepzSched sp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DeepzSmem& S = *reinterpret_cast<DeepzSmem*>(smem_raw);
    double* buf0;
    double* buf1;
    if (GLOBAL) {
        buf0 = sp.workspace + (size_t)blockIdx.x * 2 * sp.cap;
        buf1 = buf0 + sp.cap;
    } else {
        buf0 = reinterpret_cast<double*>(smem_raw + dz_meta_bytes(STATS));
        buf1 = buf0 + sp.cap;
    }
    const int tid = threadIdx.x, NT = DZ_THREADS;
    const int lane = tid & 31, warp = tid >> 5, nWarps = NT >> 5;
    const int CAP = sp.cap;
    const int L = net->L;
    const int mOut = net->m_out;
    const bool useHull = out.hull_lo != nullptr && mOut <= CLR_HULL_MAX;
    const bool needDead = STATS || out.keep_cap > 0;

    // ---- one-time setup: layer descriptors to smem, accumulators, the first tile --------------------------
    for (int i = tid; i < L * (int)(sizeof(DevLayer) / sizeof(int)); i += NT) ((int*)S.layers)[i] = ((const int*)net->layer)[i];
    if (tid < CLR_MAX_LAYERS) { S.sPin[tid] = 0; S.sUnst[tid] = 0; }
    if (tid < CLR_HULL_MAX) { S.hlo[tid] = INFINITY; S.hhi[tid] = -INFINITY; }
    const unsigned long long total_units =
        sp.pass == 0 ? (unsigned long long)in.N : (sp.pass == 1 ? st->n_retry : st->n_big);
    unsigned long long* cursor = sp.pass == 0 ? &st->next_cell : (sp.pass == 1 ? &st->next_retry : &st->next_big);
    unsigned long long pfFirst = 0;   // thread 0: prefetched tile
    int pfT = 0;
    if (tid == 0) {
        int t0 = sp.tile_cells;
        if (sp.pass != 0) t0 = 1;
        else if (t0 <= 0) {
            int guessCols = in.kind == 0 ? in.n + 1 : 8;
            t0 = CAP / ((guessCols * 2 + 8) * net->max_pitch);
            unsigned long long fair = (total_units + 2ull * gridDim.x - 1) / (2ull * gridDim.x);
            if ((unsigned long long)t0 > fair) t0 = (int)(fair < 1 ? 1 : fair);
        }
        t0 = max(1, min(t0, min(CLR_TILE_CELLS_MAX, sp.items_cells)));
        S.tileT = t0;
        S.sNear = 0; S.sMaxP = 0;
        pfT = t0;
        pfFirst = atomicAdd(cursor, (unsigned long long)pfT);
    }
    __syncthreads();
    bool needCentre = needDead;
    for (int l = 0; l < L; l++) needCentre = needCentre || (S.layers[l].act == CLR_ACT_ID && !S.layers[l].collapse);
    double a[KSREG];
#pragma unroll
    for (int k = 0; k < KSREG; k++) a[k] = 0.0;
    dz_load_a<KSREG>(S.layers[0], warp, lane, nWarps, a);

    for (;;) {
        // ---- publish the prefetched tile, prefetch the next one -------------------------------------------------
        if (tid == 0) {
            int nc = 0;
            if (pfFirst < total_units) nc = (int)min((unsigned long long)pfT, total_units - pfFirst);
            S.firstIdx = (long long)pfFirst;
            S.nCells = nc;
            S.more = nc > 0;
            S.peakNeed = 0.f;
            if (nc > 0) {
                int T = S.tileT;
                if (sp.pass == 0 && sp.tile_cells <= 0) {
                    unsigned long long seen = pfFirst + (unsigned long long)nc;
                    unsigned long long rem = seen < total_units ? total_units - seen : 0;
                    unsigned long long fair = (rem + 2ull * gridDim.x - 1) / (2ull * gridDim.x);
                    if (fair < 1) fair = 1;
                    if ((unsigned long long)T > fair) T = (int)fair;
                }
                pfT = T;
                pfFirst = atomicAdd(cursor, (unsigned long long)T);
            }
        }
        __syncthreads();
        if (!S.more) break;
        int nCells = S.nCells;
        if (tid < nCells) {
            long long idx = S.firstIdx + tid;
            S.cellId[tid] = sp.pass == 0 ? idx : (sp.pass == 1 ? sp.retry_list[idx] : sp.big_list[idx]);
            if (STATS) S.cellNear[tid] = 0;
        }
        if (needDead) for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) S.deadMask[i] = 0;
        __syncthreads();

        // ---- input zonotopes -> column store -------------------------------------------------------------------
        const int n = in.n;
        const int pin0 = S.layers[0].pitch_in;
        int flip = 0;
        int nCols;
        double* cur = buf0;
        double* oth = buf1;
        double* tmpC = oth;                       // box grid: per-cell centre / radius scratch in the free buffer
        double* tmpR = oth + nCells * n;
        if (in.kind == 0) {
            for (int it = tid; it < nCells * n; it += NT) {
                const int ci = it / n, d = it - ci * n;
                const unsigned long long g = (unsigned long long)(in.cell_begin + S.cellId[ci]);
                int k;
                if ((g >> 32) == 0) k = (int)(((unsigned)g / (unsigned)in.stride[d]) % (unsigned)in.part[d]);
                else k = (int)((g / (unsigned long long)in.stride[d]) % (unsigned long long)in.part[d]);
                tmpC[it] = __ldg(in.centers + in.base[d] + k);
                tmpR[it] = __ldg(in.radii + in.base[d] + k);
            }
            __syncthreads();
        }
        {
            int w[2] = {0, 0};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int ci = lane + 32 * h;
                if (ci < nCells) {
                    int p = 0;
                    if (in.kind == 0) {
                        for (int d = 0; d < n; d++) p += tmpR[ci * n + d] != 0.0;
                    } else {
                        const long long cell = S.cellId[ci];
                        p = (int)(in.col_off[cell + 1] - in.col_off[cell]);
                    }
                    w[h] = 1 + p;
                }
            }
            int off0, off1, keep, total;
            dz_layout(w[0], w[1], nCells, max(pin0, S.layers[0].pitch_out), CAP, lane, off0, off1, keep, total);
            if (warp == 0) {
                if (lane < keep) S.colOff[0][lane] = off0;
                if (lane + 32 < keep) S.colOff[0][lane + 32] = off1;
                if (lane == 0) {
                    S.colOff[0][keep] = total;
                    if (keep < nCells) dz_defer(sp, st, S.cellId, keep, nCells);
                    if (keep > 0) S.peakNeed = (float)((total + 8) * max(pin0, S.layers[0].pitch_out)) / (float)keep;
                }
            }
            nCells = keep;
            nCols = total;
        }
        __syncthreads();
        if (nCells == 0) continue;
        {
            const int rows4 = clr_pad4(n);
            const int* colOff = S.colOff[0];
            if (in.kind == 0) {
                for (int it = tid; it < nCells * rows4; it += NT) {
                    const int ci = it / rows4, d = it - ci * rows4;
                    const int cb = colOff[ci], p = colOff[ci + 1] - cb - 1;
                    double c = 0.0, r = 0.0;
                    int rank = 0;
                    if (d < n) {
                        c = tmpC[ci * n + d]; r = tmpR[ci * n + d];
                        for (int dd = 0; dd < d; dd++) rank += tmpR[ci * n + dd] != 0.0;
                    }
                    cur[cb * pin0 + d] = c;
                    for (int j = 0; j < p; j++) cur[(cb + 1 + j) * pin0 + d] = (r != 0.0 && j == rank) ? r : 0.0;
                }
            } else {
                for (int it = tid; it < nCells * rows4; it += NT) {
                    const int ci = it / rows4, d = it - ci * rows4;
                    const int cb = colOff[ci], p = colOff[ci + 1] - cb - 1;
                    const long long cell = S.cellId[ci];
                    const long long go = in.col_off[cell];
                    cur[cb * pin0 + d] = d < n ? in.C[cell * n + d] : 0.0;
                    for (int j = 0; j < p; j++) cur[(cb + 1 + j) * pin0 + d] = d < n ? in.G[(go + j) * n + d] : 0.0;
                }
            }
            if (needCentre) dz_build_centre_mask(S, colOff, nCells, tid, NT);
            __syncthreads();   // input store complete
        }

        // ---- layers ------------------------------------------------------------------------------------------------
        bool tileDead = false;
        for (int l = 0; l < L; l++) {
            const DevLayer& Ly = S.layers[l];
            if (STATS && Ly.net_layer >= 0 && tid < nCells) {
                // generator columns entering this layer, as the reference counts them
                const int* colOff = S.colOff[flip];
                int cb = colOff[tid], p = colOff[tid + 1] - cb - 1, nz = 0;
                for (int j = 1; j <= p; j++) nz += !((S.deadMask[(cb + j) >> 5] >> ((cb + j) & 31)) & 1u);
                S.cellPin[tid][Ly.net_layer] = (unsigned short)nz;
                S.cellUnst[tid][Ly.net_layer] = 0;
            }
            const int act = Ly.act, collapse = Ly.collapse;
            const bool relax = !(act == CLR_ACT_ID && !collapse);
            dz_gemm<KSREG>(Ly, cur, oth, nCols, S.centreMask, !relax, warp, lane, nWarps, a);
            // the masks were last read by the previous relaxation, which ended with a barrier
            for (int i = tid; i < nCells * CLR_MASK_WORDS; i += NT) (&S.mask[0][0])[i] = 0;
            // weights of the next layer (or of the first layer for the next tile) while the relaxation runs
            dz_load_a<KSREG>(S.layers[l + 1 < L ? l + 1 : 0], warp, lane, nWarps, a);
            __syncthreads();
            { double* t = cur; cur = oth; oth = t; }
            if (!relax) {
                if (needDead && Ly.remove_zero) dz_mark_dead(S, cur, Ly.pitch_out, Ly.m, nCols, tid, NT);
                continue;
            }

            // ---- relaxation pass 1: bounds, lambda, mu; thread = (cell slot, row) ---------------------------------------
            const int P = Ly.pitch_out, m = Ly.m, rowsP = clr_pad4(m);
            const int H = (rowsP + NT - 1) / NT;      // row passes (2 only for layers wider than the CTA)
            const int G = H > 1 ? 1 : NT / rowsP;     // cell slots processed side by side
            const int slot = H > 1 ? 0 : tid / rowsP, r0 = tid - slot * rowsP;
            const bool slotActive = slot < G;
            const int* colOff = S.colOff[flip];
            double lam[DZ_IMAX], mu[DZ_IMAX];
#pragma unroll
            for (int q = 0; q < DZ_IMAX; q++) {
                const int ci = slot + (H > 1 ? q >> 1 : q) * G;
                const int r = H > 1 ? r0 + (q & 1) * NT : r0;
                lam[q] = 0.0; mu[q] = 0.0;
                if (slotActive && ci < nCells && r < m) {
                    const double bias = __ldg(Ly.bias + r);
                    int cb = colOff[ci], p = colOff[ci + 1] - cb - 1;
                    double* y = cur + cb * P + r;
                    double c = y[0] + bias, lo, hi, rad = 0.0;
                    if (collapse) {
                        for (int j = 1; j <= p; j++) rad += fabs(y[j * P]);
                        lo = c - rad; hi = c + rad;
                    } else {
                        lo = c; hi = c;
                        for (int j = 1; j <= p; j++) { double av = fabs(y[j * P]); lo -= av; hi += av; }
                    }
                    bool unst = false;
                    double cn = c, la = 1.0, mm = 0.0;
                    if (act == CLR_ACT_RELU) {
                        if (STATS && (fabs(lo) < 1e-12 || fabs(hi) < 1e-12)) atomicAdd(&S.cellNear[ci], 1u);
                        if (dz_leq(lo, 0.0)) {
                            if (dz_leq(hi, 0.0)) { cn = 0.0; la = 0.0; }
                            else {
                                la = hi / (hi - lo);
                                mm = -la * lo / 2;
                                cn = c * la + mm;
                                unst = true;
                            }
                        }
                    } else if (act == CLR_ACT_SIGMOID || act == CLR_ACT_TANH) {
                        double ly = act == CLR_ACT_SIGMOID ? dz_sigm(lo) : tanh(lo);
                        double uy = act == CLR_ACT_SIGMOID ? dz_sigm(hi) : tanh(hi);
                        if (dz_approx(lo, hi)) { cn = uy; la = 0.0; }
                        else {
                            double dl = act == CLR_ACT_SIGMOID ? dz_sigm_p(lo) : dz_tanh_p(lo);
                            double du = act == CLR_ACT_SIGMOID ? dz_sigm_p(hi) : dz_tanh_p(hi);
                            la = dl < du ? dl : du;
                            double mu1 = (uy + ly - la * (hi + lo)) / 2;
                            mm = (uy - ly - la * (hi - lo)) / 2;
                            cn = c * la + mu1;
                            unst = true;
                        }
                    }
                    lam[q] = collapse ? la * rad : la;   // collapse: the single generator, already scaled
                    mu[q] = mm;
                    y[0] = cn;                            // the new centre waits in place for pass 2
                    if (unst) atomicOr(&S.mask[ci][r >> 5], 1u << (r & 31));
                }
            }
            __syncthreads();
            // ---- new layout (every warp computes it; warp 0 publishes it) -------------------------------------------------
            int off0, off1, keep, total;
            {
                int w[2] = {0, 0};
                const int mw = (m + 31) >> 5;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int ci = lane + 32 * h;
                    if (ci < nCells) {
                        int u = 0;
                        for (int ww = 0; ww < mw; ww++) u += __popc(S.mask[ci][ww]);
                        const int p = colOff[ci + 1] - colOff[ci] - 1;
                        w[h] = 1 + (collapse ? 1 : p) + u;
                        if (STATS && warp == 0 && Ly.net_layer >= 0) S.cellUnst[ci][Ly.net_layer] = (unsigned short)u;
                    }
                }
                const int pm = l + 1 < L ? max(P, S.layers[l + 1].pitch_out) : P;
                dz_layout(w[0], w[1], nCells, pm, CAP, lane, off0, off1, keep, total);
                if (warp == 0) {
                    int* cn = S.colOff[flip ^ 1];
                    if (lane < keep) cn[lane] = off0;
                    if (lane + 32 < keep) cn[lane + 32] = off1;
                    if (lane == 0) {
                        cn[keep] = total;
                        if (keep < nCells) dz_defer(sp, st, S.cellId, keep, nCells);
                        if (keep > 0) S.peakNeed = fmaxf(S.peakNeed, (float)((total + 8) * pm) / (float)keep);
                    }
                }
            }
            // ---- relaxation pass 2: scale rows, append generators, move to the new layout --------------------------------
#pragma unroll
            for (int q = 0; q < DZ_IMAX; q++) {
                const int ci = slot + (H > 1 ? q >> 1 : q) * G;
                const int r = H > 1 ? r0 + (q & 1) * NT : r0;
                const int cs = min(ci, CLR_TILE_CELLS_MAX - 1);
                const int nbA = __shfl_sync(0xffffffffu, off0, cs & 31);
                const int nbB = __shfl_sync(0xffffffffu, off1, cs & 31);
                const int nb = cs < 32 ? nbA : nbB;
                if (slotActive && ci < keep && r < rowsP) {
                    const int cb = colOff[ci], p = colOff[ci + 1] - cb - 1;
                    const unsigned mwv = r < m ? S.mask[ci][r >> 5] : 0u;
                    int u = 0;
                    for (int ww = 0; ww < ((m + 31) >> 5); ww++) u += __popc(S.mask[ci][ww]);
                    const double* y = cur + cb * P + r;
                    double* z = oth + nb * P + r;
                    if (r >= m) {
                        const int wtot = 1 + (collapse ? 1 : p) + u;
                        for (int j = 0; j < wtot; j++) z[j * P] = 0.0;
                    } else {
                        z[0] = y[0];
                        int pk;
                        if (collapse) { z[P] = lam[q]; pk = 1; }
                        else {
                            const double la = lam[q];
                            for (int j = 1; j <= p; j++) z[j * P] = y[j * P] * la;
                            pk = p;
                        }
                        if (u > 0) {
                            const bool unst = (mwv >> (r & 31)) & 1u;
                            int rank = __popc(mwv & ((1u << (r & 31)) - 1u));
                            for (int ww = 0; ww < (r >> 5); ww++) rank += __popc(S.mask[ci][ww]);
                            for (int k = 0; k < u; k++) z[(1 + pk + k) * P] = (unst && k == rank) ? mu[q] : 0.0;
                        }
                    }
                }
            }
            __syncthreads();
            { double* t = cur; cur = oth; oth = t; }
            flip ^= 1;
            nCells = keep;
            nCols = total;
            if (nCells == 0) {
                tileDead = true;
                dz_load_a<KSREG>(S.layers[0], warp, lane, nWarps, a);   // the next tile starts at layer 0 again
                break;
            }
            if (needCentre) dz_build_centre_mask(S, S.colOff[flip], nCells, tid, NT);
            if (needDead) {
                if (Ly.remove_zero) dz_mark_dead(S, cur, P, m, nCols, tid, NT);
                else { for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) S.deadMask[i] = 0; }
            }
        }
        __syncthreads();
        if (tileDead) continue;

        // ---- outputs ---------------------------------------------------------------------------------------------------
        {
            const int P = S.layers[L - 1].pitch_out;
            const int* colOff = S.colOff[flip];
            // ranks of the surviving output generators (remove_zero_columns), in the free buffer
            unsigned short* colRank = reinterpret_cast<unsigned short*>(oth);
            if (needDead) {
                if (tid < nCells) {
                    int cb = colOff[tid], p = colOff[tid + 1] - cb - 1, k = 0;
                    colRank[cb] = 0xFFFF;
                    for (int j = 1; j <= p; j++) {
                        const bool dead = (S.deadMask[(cb + j) >> 5] >> ((cb + j) & 31)) & 1u;
                        colRank[cb + j] = dead ? (unsigned short)0xFFFF : (unsigned short)(k++);
                    }
                    if (STATS) atomicMax(&S.sMaxP, k);
                    if (out.keep_cap > 0) {
                        out.keepN[S.cellId[tid]] = k;
                        if (k > out.keep_cap) atomicCAS(&st->error, 0, (int)CLR_ERR_CAPACITY);
                    }
                }
                __syncthreads();
            }
            for (int it = tid; it < nCells * mOut; it += NT) {
                int ci = it / mOut, r = it - ci * mOut;
                int cb = colOff[ci], p = colOff[ci + 1] - cb - 1;
                const double* y = cur + (size_t)cb * P + r;
                double c = y[0], lo, hi;
                if (mOut == 1) {
                    lo = c; hi = c;
                    for (int j = 1; j <= p; j++) { double av = fabs(y[(size_t)j * P]); lo -= av; hi += av; }
                } else {
                    double rad = 0.0;
                    for (int j = 1; j <= p; j++) rad += fabs(y[(size_t)j * P]);
                    lo = c - rad; hi = c + rad;
                }
                long long cell = S.cellId[ci];
                if (out.out_lo) { out.out_lo[cell * mOut + r] = lo; out.out_hi[cell * mOut + r] = hi; }
                if (out.hull_lo) {
                    if (useHull) { atomicMinD(&S.hlo[r], lo); atomicMaxD(&S.hhi[r], hi); }
                    else { atomicMinD(&out.hull_lo[r], lo); atomicMaxD(&out.hull_hi[r], hi); }
                }
                if (out.keep_cap > 0) {
                    out.keepC[cell * mOut + r] = c;
                    double* g = out.keepG + (size_t)cell * out.keep_cap * mOut + r;
                    for (int j = 1; j <= p; j++) {
                        unsigned short k = colRank[cb + j];
                        if (k != 0xFFFF && k < out.keep_cap) g[(size_t)k * mOut] = y[(size_t)j * P];
                    }
                }
            }
            if (STATS) {
                __syncthreads();
                if (tid < nCells) {
                    for (int l2 = 0; l2 < L; l2++) {
                        int nl = S.layers[l2].net_layer;
                        if (nl >= 0) {
                            atomicAdd(&S.sPin[nl], (unsigned long long)S.cellPin[tid][nl]);
                            atomicAdd(&S.sUnst[nl], (unsigned long long)S.cellUnst[tid][nl]);
                        }
                    }
                    atomicAdd(&S.sNear, (unsigned long long)S.cellNear[tid]);
                }
            }
        }
        // ---- adapt the tile size to the growth just seen ---------------------------------------------------------------
        if (tid == 0) {
            if (sp.pass == 0 && sp.tile_cells <= 0) {
                float need = S.peakNeed * 1.0625f + 1.f;
                int T = (int)((float)CAP / need);
                S.tileT = max(1, min(T, min(CLR_TILE_CELLS_MAX, sp.items_cells)));
            }
            if (STATS) atomicAdd(&st->tiles, 1ull);
        }
        __syncthreads();
    }

    // ---- per-CTA accumulators -> global -----------------------------------------------------------------------------------
    __syncthreads();
    if (STATS) {
        if (tid < CLR_MAX_LAYERS) {
            if (S.sPin[tid]) atomicAdd(&st->sum_p_in[tid], S.sPin[tid]);
            if (S.sUnst[tid]) atomicAdd(&st->sum_unstable[tid], S.sUnst[tid]);
        }
        if (tid == 0) {
            if (S.sNear) atomicAdd(&st->near_threshold, S.sNear);
            atomicMax(&st->max_p_out, S.sMaxP);
        }
    }
    if (useHull && tid < mOut) {
        if (S.hlo[tid] <= S.hhi[tid]) {
            atomicMinD(&out.hull_lo[tid], S.hlo[tid]);
            atomicMaxD(&out.hull_hi[tid], S.hhi[tid]);
        }
    }
}
