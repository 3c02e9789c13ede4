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
#include <cstddef>
#include "common.cuh"

struct DeepzInput {
    int kind;                 // 0: box grid, 1: explicit zonotopes
    int n;                    // rows of the input sets
    long long N;              // cells of this call
    long long cell_begin;     // box grid: id of the first cell in product order
    int part[64];
    int base[64];
    const double* centers;    // device
    const double* radii;      // device
    const double* C;          // device, n x N
    const long long* col_off; // device, N + 1
    const double* G;          // device, n x col_off[N]
};

struct DeepzOutput {
    double* out_lo;           // m_out x N or nullptr
    double* out_hi;
    double* hull_lo;          // m_out or nullptr (initialised to +inf / -inf by the host)
    double* hull_hi;
    int keep_cap;
    int* keepN;               // N
    double* keepC;            // m_out x N
    double* keepG;            // m_out x keep_cap x N
};

struct DeepzSched {
    int pass;                 // 0: main (adaptive tiles), 1: retry list (1 cell / tile), 2: big list (global store)
    int tile_cells;           // fixed tile size (0 = adaptive)
    int cap;                  // doubles per column-store buffer
    int items_cells;          // max cells per tile from the relaxation item budget
    long long* retry_list;
    long long* big_list;
    double* workspace;        // pass 2: gridDim.x * 2 * cap doubles
};

struct DeepzSmem {
    DevLayer layer;
    int nCells, nCols, tileT, more;
    long long firstIdx;
    float peakNeed;
    int sMaxP;
    int colOff[CLR_TILE_CELLS_MAX + 1];
    int colOffNew[CLR_TILE_CELLS_MAX + 1];
    long long cellId[CLR_TILE_CELLS_MAX];
    unsigned mask[CLR_TILE_CELLS_MAX][CLR_MASK_WORDS];
    unsigned centreMask[CLR_NCOL_MAX / 32];
    unsigned deadMask[CLR_NCOL_MAX / 32];   // generators the reference's remove_zero_columns has dropped
    double hlo[CLR_HULL_MAX], hhi[CLR_HULL_MAX];
    unsigned long long sPin[CLR_MAX_LAYERS], sUnst[CLR_MAX_LAYERS], sNear;
    // ---- only allocated by the STATS instantiations (see dz_meta_bytes) ----
    unsigned int cellNear[CLR_TILE_CELLS_MAX];
    unsigned short cellPin[CLR_TILE_CELLS_MAX][CLR_MAX_LAYERS];
    unsigned short cellUnst[CLR_TILE_CELLS_MAX][CLR_MAX_LAYERS];
};
__host__ __device__ inline size_t dz_meta_bytes(bool stats) {
    size_t b = stats ? sizeof(DeepzSmem) : offsetof(DeepzSmem, cellNear);
    return (b + 15) & ~(size_t)15;
}

#define DZ_IMAX 8   // relaxation items (cell, row) per thread

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
// Layer product: Y[:, j] = W * X[:, j] (+ bias on centre columns) for all nCols columns.
// Warp w owns row block rb (8 rows); A fragments of W stay in registers when ksteps <= KSREG.
// ---------------------------------------------------------------------------------------------
template <int KSREG>
__device__ __forceinline__ void dz_gemm(const DevLayer& L, const double* __restrict__ X, double* __restrict__ Y,
                                        int nCols, const unsigned* centreMask, int warp, int lane, int nWarps) {
    constexpr int NB = 4;                       // column blocks (8 columns each) per group
    const int RB = L.rblocks, KS = L.ksteps;
    const int CG = (nCols + 8 * NB - 1) / (8 * NB);
    const int pin = L.pitch_in, pout = L.pitch_out;
    const int rowsW = clr_pad4(L.m);            // rows written (pad rows of W are zero)
    const int qr = lane >> 2, qc = lane & 3;
    int rb0, rbStep, part, nParts;
    if (RB >= nWarps) { rb0 = warp; rbStep = nWarps; part = 0; nParts = 1; }
    else { nParts = nWarps / RB; rb0 = warp % RB; rbStep = RB; part = warp / RB; if (part >= nParts) return; }
    for (int rb = rb0; rb < RB; rb += rbStep) {
        const double* Af = L.Afrag + (size_t)rb * KS * 32 + lane;
        const int row = rb * 8 + qr;
        const double bias = row < L.m ? __ldg(L.bias + row) : 0.0;
        if (KS <= KSREG) {
            double a[KSREG];
#pragma unroll
            for (int k = 0; k < KSREG; k++) a[k] = k < KS ? __ldg(Af + k * 32) : 0.0;
            for (int cg = part; cg < CG; cg += nParts) {
                double d[NB][2];
#pragma unroll
                for (int j = 0; j < NB; j++) { d[j][0] = 0.0; d[j][1] = 0.0; }
                const double* xb = X + (size_t)(cg * NB * 8 + qr) * pin + qc;
                const int nbAct = min(NB, (nCols - cg * NB * 8 + 7) >> 3);   // column blocks that hold real columns
#pragma unroll
                for (int k = 0; k < KSREG; k++) {
                    if (k < KS) {
#pragma unroll
                        for (int j = 0; j < NB; j++) {
                            if (j < nbAct) {
                                double b = xb[(size_t)j * 8 * pin + k * 4];
                                dmma884(d[j][0], d[j][1], a[k], b);
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
                                bool centre = (centreMask[col >> 5] >> (col & 31)) & 1u;
                                Y[(size_t)col * pout + row] = centre ? d[j][e] + bias : d[j][e];
                            }
                        }
                    }
                }
            }
        } else {
            // wide-K layers: stream A fragments per column group
            for (int cg = part; cg < CG; cg += nParts) {
                double d[NB][2];
#pragma unroll
                for (int j = 0; j < NB; j++) { d[j][0] = 0.0; d[j][1] = 0.0; }
                const double* xb = X + (size_t)(cg * NB * 8 + qr) * pin + qc;
                const int nbAct = min(NB, (nCols - cg * NB * 8 + 7) >> 3);
                for (int k = 0; k < KS; k++) {
                    double a = __ldg(Af + k * 32);
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        if (j < nbAct) {
                            double b = xb[(size_t)j * 8 * pin + k * 4];
                            dmma884(d[j][0], d[j][1], a, b);
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
                                bool centre = (centreMask[col >> 5] >> (col & 31)) & 1u;
                                Y[(size_t)col * pout + row] = centre ? d[j][e] + bias : d[j][e];
                            }
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

__device__ __forceinline__ double dz_sigm(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double dz_sigm_p(double x) { double ex = exp(x); return ex / ((1.0 + ex) * (1.0 + ex)); }
__device__ __forceinline__ double dz_tanh_p(double x) { double t = tanh(x); return 1.0 - t * t; }

// ---------------------------------------------------------------------------------------------
// The persistent tile kernel.
// ---------------------------------------------------------------------------------------------
template <int KSREG, bool GLOBAL, bool STATS>
__global__ void __launch_bounds__(512, 1)
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
    const int tid = threadIdx.x, NT = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nWarps = NT >> 5;
    const int CAP = sp.cap;
    const int L = net->L;
    const int mOut = net->m_out;
    const bool useHull = out.hull_lo != nullptr && mOut <= CLR_HULL_MAX;
    const bool needDead = STATS || out.keep_cap > 0;

    if (tid == 0) {
        int t0 = sp.tile_cells;
        if (sp.pass != 0) t0 = 1;
        else if (t0 <= 0) {
            int guessCols = in.kind == 0 ? in.n + 1 : 8;
            t0 = CAP / ((guessCols * 3 + 8) * net->max_pitch);
        }
        t0 = max(1, min(t0, min(CLR_TILE_CELLS_MAX, sp.items_cells)));
        S.tileT = t0;
        S.sNear = 0; S.sMaxP = 0;
    }
    if (tid < CLR_MAX_LAYERS) { S.sPin[tid] = 0; S.sUnst[tid] = 0; }
    if (tid < CLR_HULL_MAX) { S.hlo[tid] = INFINITY; S.hhi[tid] = -INFINITY; }
    __syncthreads();

    for (;;) {
        // ---- fetch a tile ----------------------------------------------------------------------
        if (tid == 0) {
            unsigned long long total = sp.pass == 0 ? (unsigned long long)in.N : (sp.pass == 1 ? st->n_retry : st->n_big);
            unsigned long long* cur = sp.pass == 0 ? &st->next_cell : (sp.pass == 1 ? &st->next_retry : &st->next_big);
            int T = S.tileT;
            if (sp.pass == 0 && sp.tile_cells <= 0) {
                unsigned long long seen = *((volatile unsigned long long*)cur);
                unsigned long long rem = seen < total ? total - seen : 0;
                unsigned long long fair = (rem + 2ull * gridDim.x - 1) / (2ull * gridDim.x);
                if (fair < 1) fair = 1;
                if ((unsigned long long)T > fair) T = (int)fair;
            }
            unsigned long long first = atomicAdd(cur, (unsigned long long)T);
            int nc = 0;
            if (first < total) nc = (int)min((unsigned long long)T, total - first);
            S.firstIdx = (long long)first;
            S.nCells = nc;
            S.more = nc > 0;
            S.peakNeed = 0.f;
        }
        __syncthreads();
        if (!S.more) break;
        int nCells = S.nCells;
        if (tid < nCells) {
            long long idx = S.firstIdx + tid;
            S.cellId[tid] = sp.pass == 0 ? idx : (sp.pass == 1 ? sp.retry_list[idx] : sp.big_list[idx]);
            if (STATS) S.cellNear[tid] = 0;
        }
        __syncthreads();

        // ---- input zonotopes -> column store ------------------------------------------------------
        const int n = in.n;
        const int pin0 = net->layer[0].pitch_in;
        const int pout0 = net->layer[0].pitch_out;
        if (tid < nCells) {
            long long cell = S.cellId[tid];
            int p = 0;
            if (in.kind == 0) {
                long long g = in.cell_begin + cell;
                for (int d = 0; d < n; d++) {
                    int k = (int)(g % in.part[d]); g /= in.part[d];
                    if (in.radii[in.base[d] + k] != 0.0) p++;
                }
            } else {
                p = (int)(in.col_off[cell + 1] - in.col_off[cell]);
            }
            S.colOffNew[tid] = 1 + p;
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, keep = nCells;
            int pm = max(pin0, pout0);
            for (int i = 0; i < nCells; i++) {
                int w = S.colOffNew[i];
                if ((long long)(acc + w + 8) * pm > CAP || acc + w > CLR_NCOL_MAX) { keep = i; break; }
                S.colOff[i] = acc; acc += w;
            }
            S.colOff[keep] = acc;
            S.nCols = acc;
            if (keep < nCells) {
                // cells keep..nCells-1 are deferred
                int nd = nCells - keep;
                if (sp.pass == 0 && keep > 0) {
                    unsigned long long o = atomicAdd(&st->n_retry, (unsigned long long)nd);
                    for (int i = 0; i < nd; i++) sp.retry_list[o + i] = S.cellId[keep + i];
                } else if (sp.pass == 0) {
                    unsigned long long o = atomicAdd(&st->n_big, 1ull);
                    sp.big_list[o] = S.cellId[0];
                    if (nd > 1) {
                        unsigned long long o2 = atomicAdd(&st->n_retry, (unsigned long long)(nd - 1));
                        for (int i = 1; i < nd; i++) sp.retry_list[o2 + i - 1] = S.cellId[i];
                    }
                } else if (sp.pass == 1) {
                    unsigned long long o = atomicAdd(&st->n_big, (unsigned long long)nd);
                    for (int i = 0; i < nd; i++) sp.big_list[o + i] = S.cellId[keep + i];
                } else {
                    atomicCAS(&st->error, 0, (int)CLR_ERR_CAPACITY);
                }
                atomicAdd(&st->dropped, (unsigned long long)nd);
                S.nCells = keep;
            }
            if (keep > 0) S.peakNeed = (float)((acc + 8) * pm) / (float)keep;
        }
        for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) { S.centreMask[i] = 0; S.deadMask[i] = 0; }
        __syncthreads();
        nCells = S.nCells;
        if (nCells == 0) { __syncthreads(); continue; }
        int nCols = S.nCols;
        double* cur = buf0;
        double* oth = buf1;
        {
            const int rows4 = clr_pad4(n);
            if (in.kind == 0) {
                for (int it = tid; it < nCells * rows4; it += NT) {
                    int ci = it / rows4, d = it - ci * rows4;
                    int cb = S.colOff[ci], p = S.colOff[ci + 1] - cb - 1;
                    double c = 0.0, r = 0.0;
                    int rank = 0;
                    if (d < n) {
                        long long g = in.cell_begin + S.cellId[ci];
                        for (int dd = 0; dd <= d; dd++) {
                            int k = (int)(g % in.part[dd]); g /= in.part[dd];
                            double rr = in.radii[in.base[dd] + k];
                            if (dd < d) rank += rr != 0.0;
                            else { r = rr; c = in.centers[in.base[dd] + k]; }
                        }
                    }
                    cur[(size_t)cb * pin0 + d] = c;
                    for (int j = 0; j < p; j++) cur[(size_t)(cb + 1 + j) * pin0 + d] = (r != 0.0 && j == rank) ? r : 0.0;
                }
            } else {
                for (int it = tid; it < nCells * rows4; it += NT) {
                    int ci = it / rows4, d = it - ci * rows4;
                    int cb = S.colOff[ci], p = S.colOff[ci + 1] - cb - 1;
                    long long cell = S.cellId[ci];
                    long long go = in.col_off[cell];
                    cur[(size_t)cb * pin0 + d] = d < n ? in.C[cell * n + d] : 0.0;
                    for (int j = 0; j < p; j++) cur[(size_t)(cb + 1 + j) * pin0 + d] = d < n ? in.G[(go + j) * n + d] : 0.0;
                }
            }
            if (tid < nCells) {
                int cb = S.colOff[tid];
                atomicOr(&S.centreMask[cb >> 5], 1u << (cb & 31));
            }
        }

        // ---- layers ---------------------------------------------------------------------------------
        bool tileDead = false;
        for (int l = 0; l < L; l++) {
            __syncthreads();
            if (tid < (int)(sizeof(DevLayer) / sizeof(int))) ((int*)&S.layer)[tid] = ((const int*)&net->layer[l])[tid];
            for (int i = tid; i < nCells * CLR_MASK_WORDS; i += NT) (&S.mask[0][0])[i] = 0;
            __syncthreads();
            const DevLayer& Ly = S.layer;
            if (STATS && Ly.net_layer >= 0 && tid < nCells) {
                // generator columns entering this layer, as the reference counts them
                int cb = S.colOff[tid], p = S.colOff[tid + 1] - cb - 1, nz = 0;
                for (int j = 1; j <= p; j++) nz += !((S.deadMask[(cb + j) >> 5] >> ((cb + j) & 31)) & 1u);
                S.cellPin[tid][Ly.net_layer] = (unsigned short)nz;
                S.cellUnst[tid][Ly.net_layer] = 0;
            }
            dz_gemm<KSREG>(Ly, cur, oth, nCols, S.centreMask, warp, lane, nWarps);
            __syncthreads();
            { double* t = cur; cur = oth; oth = t; }
            const int act = Ly.act, collapse = Ly.collapse;
            if (act == CLR_ACT_ID && !collapse) {
                if (needDead && Ly.remove_zero) dz_mark_dead(S, cur, Ly.pitch_out, Ly.m, nCols, tid, NT);
                continue;
            }

            // ---- relaxation pass 1: bounds, lambda, mu ----------------------------------------------
            const int P = Ly.pitch_out, m = Ly.m, rows4 = clr_pad4(m);
            const int nItems = nCells * rows4;
            double cnew[DZ_IMAX], lam[DZ_IMAX], mu[DZ_IMAX];
#pragma unroll
            for (int q = 0; q < DZ_IMAX; q++) {
                int it = tid + q * NT;
                cnew[q] = 0.0; lam[q] = 0.0; mu[q] = 0.0;
                if (it < nItems) {
                    int ci = it / rows4, r = it - ci * rows4;
                    if (r < m) {
                        int cb = S.colOff[ci], p = S.colOff[ci + 1] - cb - 1;
                        const double* y = cur + (size_t)cb * P + r;
                        double c = y[0], lo, hi, rad = 0.0;
                        if (collapse) {
                            for (int j = 1; j <= p; j++) rad += fabs(y[(size_t)j * P]);
                            lo = c - rad; hi = c + rad;
                        } else {
                            lo = c; hi = c;
                            for (int j = 1; j <= p; j++) { double a = fabs(y[(size_t)j * P]); lo -= a; hi += a; }
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
                        if (collapse) lam[q] = la * rad;   // the single generator after collapse, already scaled
                        else lam[q] = la;
                        cnew[q] = cn; mu[q] = mm;
                        if (unst) atomicOr(&S.mask[ci][r >> 5], 1u << (r & 31));
                    }
                }
            }
            __syncthreads();
            // ---- new layout ----------------------------------------------------------------------------
            if (tid < nCells) {
                int u = 0;
                for (int w = 0; w < (m + 31) / 32; w++) u += __popc(S.mask[tid][w]);
                int p = S.colOff[tid + 1] - S.colOff[tid] - 1;
                S.colOffNew[tid] = 1 + (collapse ? 1 : p) + u;
                if (STATS && Ly.net_layer >= 0) S.cellUnst[tid][Ly.net_layer] = (unsigned short)u;
            }
            __syncthreads();
            if (tid == 0) {
                int pm = P;
                if (l + 1 < L) pm = max(pm, net->layer[l + 1].pitch_out);
                int acc = 0, keep = nCells;
                for (int i = 0; i < nCells; i++) {
                    int w = S.colOffNew[i];
                    if ((long long)(acc + w + 8) * pm > CAP || acc + w > CLR_NCOL_MAX) { keep = i; break; }
                    S.colOffNew[i] = acc; acc += w;
                }
                S.colOffNew[keep] = acc;
                if (keep < nCells) {
                    int nd = nCells - keep;
                    if (sp.pass == 0 && keep > 0) {
                        unsigned long long o = atomicAdd(&st->n_retry, (unsigned long long)nd);
                        for (int i = 0; i < nd; i++) sp.retry_list[o + i] = S.cellId[keep + i];
                    } else if (sp.pass == 0) {
                        unsigned long long o = atomicAdd(&st->n_big, 1ull);
                        sp.big_list[o] = S.cellId[0];
                        if (nd > 1) {
                            unsigned long long o2 = atomicAdd(&st->n_retry, (unsigned long long)(nd - 1));
                            for (int i = 1; i < nd; i++) sp.retry_list[o2 + i - 1] = S.cellId[i];
                        }
                    } else if (sp.pass == 1) {
                        unsigned long long o = atomicAdd(&st->n_big, (unsigned long long)nd);
                        for (int i = 0; i < nd; i++) sp.big_list[o + i] = S.cellId[keep + i];
                    } else {
                        atomicCAS(&st->error, 0, (int)CLR_ERR_CAPACITY);
                    }
                    atomicAdd(&st->dropped, (unsigned long long)nd);
                }
                S.nCells = keep;
                S.nCols = acc;
                if (keep > 0) S.peakNeed = fmaxf(S.peakNeed, (float)((acc + 8) * pm) / (float)keep);
            }
            __syncthreads();
            const int nKeep = S.nCells;
            // ---- relaxation pass 2: scale rows, append generators, move to the new layout ----------------
#pragma unroll
            for (int q = 0; q < DZ_IMAX; q++) {
                int it = tid + q * NT;
                if (it < nItems) {
                    int ci = it / rows4, r = it - ci * rows4;
                    if (ci < nKeep) {
                        int cb = S.colOff[ci], p = S.colOff[ci + 1] - cb - 1;
                        int nb = S.colOffNew[ci], u = S.colOffNew[ci + 1] - nb - 1 - (collapse ? 1 : p);
                        const double* y = cur + (size_t)cb * P + r;
                        double* z = oth + (size_t)nb * P + r;
                        if (r >= m) {
                            int w = S.colOffNew[ci + 1] - nb;
                            for (int j = 0; j < w; j++) z[(size_t)j * P] = 0.0;
                        } else {
                            z[0] = cnew[q];
                            int pk;
                            if (collapse) { z[P] = lam[q]; pk = 1; }
                            else {
                                const double la = lam[q];
                                for (int j = 1; j <= p; j++) z[(size_t)j * P] = y[(size_t)j * P] * la;
                                pk = p;
                            }
                            if (u > 0) {
                                const unsigned w = S.mask[ci][r >> 5];
                                const bool unst = (w >> (r & 31)) & 1u;
                                int rank = __popc(w & ((1u << (r & 31)) - 1u));
                                for (int ww = 0; ww < (r >> 5); ww++) rank += __popc(S.mask[ci][ww]);
                                for (int k = 0; k < u; k++) z[(size_t)(1 + pk + k) * P] = (unst && k == rank) ? mu[q] : 0.0;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            { double* t = cur; cur = oth; oth = t; }
            nCells = nKeep;
            if (nCells == 0) { tileDead = true; break; }
            if (tid <= nCells) S.colOff[tid] = S.colOffNew[tid];
            for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) S.centreMask[i] = 0;
            __syncthreads();
            if (tid < nCells) {
                int cb = S.colOff[tid];
                atomicOr(&S.centreMask[cb >> 5], 1u << (cb & 31));
            }
            nCols = S.nCols;
            if (needDead) {
                __syncthreads();
                if (Ly.remove_zero) dz_mark_dead(S, cur, P, m, nCols, tid, NT);
                else { for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) S.deadMask[i] = 0; }
            }
        }
        __syncthreads();
        if (tileDead) continue;

        // ---- outputs ------------------------------------------------------------------------------------
        {
            const int P = net->layer[L - 1].pitch_out;
            // non-zero flags / ranks of the output generators (remove_zero_columns), in the free buffer
            unsigned short* colRank = reinterpret_cast<unsigned short*>(oth);
            const bool needRank = needDead;
            if (needRank) {
                if (tid < nCells) {
                    int cb = S.colOff[tid], p = S.colOff[tid + 1] - cb - 1, k = 0;
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
                int cb = S.colOff[ci], p = S.colOff[ci + 1] - cb - 1;
                const double* y = cur + (size_t)cb * P + r;
                double c = y[0], lo, hi;
                if (mOut == 1) {
                    lo = c; hi = c;
                    for (int j = 1; j <= p; j++) { double a = fabs(y[(size_t)j * P]); lo -= a; hi += a; }
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
                        int nl = net->layer[l2].net_layer;
                        if (nl >= 0) {
                            atomicAdd(&S.sPin[nl], (unsigned long long)S.cellPin[tid][nl]);
                            atomicAdd(&S.sUnst[nl], (unsigned long long)S.cellUnst[tid][nl]);
                        }
                    }
                    atomicAdd(&S.sNear, (unsigned long long)S.cellNear[tid]);
                }
            }
        }
        // ---- adapt the tile size to the growth just seen ----------------------------------------------------
        if (tid == 0) {
            if (sp.pass == 0 && sp.tile_cells <= 0) {
                float need = S.peakNeed * 1.125f + 1.f;
                int T = (int)((float)CAP / need);
                S.tileT = max(1, min(T, min(CLR_TILE_CELLS_MAX, sp.items_cells)));
            }
            if (STATS) atomicAdd(&st->tiles, 1ull);
        }
        __syncthreads();
    }

    // ---- per-CTA accumulators -> global ---------------------------------------------------------------------
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
