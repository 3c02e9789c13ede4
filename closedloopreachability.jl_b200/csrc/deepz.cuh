// deepz.cuh -- batched Float64 DeepZ propagation of zonotope cells through a dense network.
//
// Replaces the per-cell loop of the reference (src/solve.jl:153-166, :187-195) around
// controller_forward (src/solve.jl:210-218) with one persistent kernel:
//
//   * a CTA owns a TILE of cells; the tile's zonotopes ([centre | generators] per cell, all cells
//     side by side) live in shared memory as ONE column store (column-major, one conflict-free pitch
//     for the whole network) that every layer transforms IN PLACE;
//   * per layer, W * [c | G] for all columns of the tile runs on the FP64 tensor pipe
//     (mma.sync.m8n8k4.f64 -> SASS DMMA; tcgen05 has no FP64 kind): a warp owns 8 output rows,
//     keeps their W fragments in registers and streams the columns as B fragments from smem.  The
//     sweep goes over the column groups back to front; results are written to the columns' positions
//     in the NEXT layout (cells only ever move right, making room for appended generators), and a
//     named barrier per column group keeps writers behind all readers;
//   * generators appended by the previous relaxation are mu * e_i: their image under W is mu * W[:, i],
//     so they never go through the tensor pipe -- they are materialised as scaled weight columns;
//   * the interval bounds (sequential |.| row sums, same order as LazySets low/high), the
//     ReLU / sigmoid / tanh relaxation and the row scaling happen in place in the same kernel and never
//     leave the SM; the appended generators are kept as (row mask, mu) until the next layer;
//   * the output box (src/solve.jl:214-216, src/setops.jl:76-83), the hull over cells and the
//     optional output zonotopes are written straight from smem.
//
// Tiles are handed out by an atomic cursor (prefetched one tile ahead); the tile size adapts to the
// generator growth seen so far; cells that overflow a tile are retried in smaller tiles and, if a
// single cell does not fit in shared memory, in a pass whose column store lives in a global
// (L2-resident) workspace.
#pragma once
#include "deepz_types.cuh"

#ifdef DZ_TIMING
#define DZ_T(i) do { if (tid == 0) { long long tn_ = clock64(); tacc[i] += tn_ - tlast; tlast = tn_; } } while (0)
#else
#define DZ_T(i) do { } while (0)
#endif

//@phase: product (dz_gemm)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

//@phase: relaxation
__device__ __forceinline__ bool dz_approxzero(double x) { return fabs(x) <= CLR_ZTOL; }
__device__ __forceinline__ bool dz_approx(double x, double y) {
    if (dz_approxzero(x) && dz_approxzero(y)) return true;
    return x == y || fabs(x - y) <= CLR_RTOL * fmax(fabs(x), fabs(y));
}
__device__ __forceinline__ bool dz_leq(double x, double y) { return x <= y || dz_approx(x, y); }

//@phase: output
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
// TMA (bulk async copy) staging of W fragments.  Layers whose K exceeds what a warp can keep in registers (K > 4 KSREG,
// e.g. Unicycle's 500 -> 2 layer) read their A fragments once per k-step and column group; instead of fetching them from
// L2 for every group, the layer's whole fragment image (rblocks x ksteps x 256 B) is copied to shared memory ONCE per
// (tile, layer) by one thread with cp.async.bulk (SASS UBLKCP) and completed on an mbarrier; the k-loop then reads it with
// conflict-free LDS.64.  Only when the image fits the staging area the host reserved (DeepzSched::astage_bytes).
// ---------------------------------------------------------------------------------------------
//@phase: product (dz_gemm)
__device__ __forceinline__ unsigned dz_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dz_mbar_init(unsigned long long* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(dz_smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
// one thread: copy `bytes` (multiple of 16) global -> shared, completion counted on `bar`
__device__ __forceinline__ void dz_tma_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // earlier generic reads of dst are ordered before the async write
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(dz_smem_u32(bar)), "r"(bytes) : "memory");
    for (unsigned o = 0; o < bytes; o += 32768u) {
        const unsigned n = min(32768u, bytes - o);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(dz_smem_u32((const char*)dst + o)), "l"((const char*)src + o), "r"(n), "r"(dz_smem_u32(bar)) : "memory");
    }
}
__device__ __forceinline__ void dz_mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(dz_smem_u32(bar)), "r"(phase) : "memory");
    } while (!ok);
}

// ---------------------------------------------------------------------------------------------
// Warp -> (row block, column part) mapping of the layer product.  16 warps sit on 4 SM sub-partitions
// (warp w on w % 4); the DMMA rate is per sub-partition, so the row blocks are dealt such that the
// sub-partitions get the same number of DMMAs: with 13 row blocks (100 neurons) warps 0..11 own one
// row block each and warps 12..15 share the last one, a quarter of the column groups each.
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
// Layer product.  Column j of X (current layout) goes to column colPos[j] of Y (next layout);
// Y == X when `inplace` (then the groups are swept back to front, one named barrier per group).
// Bias is added by the relaxation, or here on the centre columns for layers without one.
// ---------------------------------------------------------------------------------------------
template <int KSREG, bool STREAM>
__device__ __forceinline__ void dz_gemm(const DevLayer& L, const double* X, double* Y, int P, int nCols, int kAct,
                                        const unsigned short* colPos, bool addBias, bool inplace,
                                        int warp, int lane, int nWarps,
                                        double (&a)[KSREG], const double* Astage = nullptr) {
    constexpr int NB = DZ_NB;                   // column blocks (8 columns each) per group
    const int RB = L.rblocks, KS = L.ksteps;
    const int ksEff = min(KS, (kAct + 3) >> 2);    // rows >= kAct of every column are zero: their k-steps are skipped
    const int CG = (nCols + 8 * NB - 1) / (8 * NB);
    const int rowsW = clr_pad4(L.m);            // rows written (pad rows of W are zero)
    const int qr = lane >> 2, qc = lane & 3;
    const DzMap mp = dz_map(RB, warp, nWarps);
    const bool resident = KS <= KSREG;
    // in place every warp walks all groups exactly once (it owns at most one row block); otherwise row blocks outermost
    for (int rb = mp.rb0; rb < RB || (inplace && rb == mp.rb0); rb += mp.rbStep) {
        const bool active = rb < RB;
        const double* Af = L.Afrag + (size_t)(active ? rb : 0) * KS * 32 + lane;
        const int row = rb * 8 + qr;
        const double bias = (addBias && active && row < L.m) ? __ldg(L.bias + row) : 0.0;
        if (active && resident && rb != mp.rb0) {
#pragma unroll
            for (int k = 0; k < KSREG; k++) a[k] = k < KS ? __ldg(Af + k * 32) : 0.0;
        }
        // a row block shared by several warps is split by column block inside every group first (so that each group
        // carries the same work on every SM sub-partition), then by group
        const int bp = min(mp.nParts, NB), gp = max(1, mp.nParts / bp);
        const int myB = mp.part % bp, myG = mp.part / bp;
        const bool partOk = mp.part < bp * gp;
        unsigned blkMask = 0;     // column blocks j of a group with j % bp == myB (bp divides NB or is NB itself in all common maps)
        if (bp == 1) blkMask = (1u << NB) - 1u;
        else if (bp == NB) blkMask = 1u << myB;
        else if (bp == 2) { for (int j = 0; j < NB; j += 2) blkMask |= 1u << (j + myB); }
        else { for (int j = 0; j < NB; j++) if ((j % bp) == myB) blkMask |= 1u << j; }
        int cgm = gp == 1 ? 0 : (CG - 1) % gp;      // cg % gp, kept incrementally
        for (int cg = CG - 1; cg >= 0; cg--) {
            const bool mine = active && partOk && (gp == 1 || cgm == myG);
            if (gp != 1) cgm = cgm == 0 ? gp - 1 : cgm - 1;
            double d[NB][2];
#pragma unroll
            for (int j = 0; j < NB; j++) { d[j][0] = 0.0; d[j][1] = 0.0; }
            const int nbAct = min(NB, (nCols - cg * NB * 8 + 7) >> 3);   // column blocks that hold real columns
            if (mine) {
                const double* xb = X + (cg * NB * 8 + qr) * P + qc;
                if (resident && nbAct == NB && blkMask == ((1u << NB) - 1u)) {
                    // full group, whole row block: no predicates, so the loads run ahead of the tensor pipe
#pragma unroll
                    for (int k = 0; k < KSREG; k++) {
                        if (k >= ksEff) break;
#pragma unroll
                        for (int j = 0; j < NB; j++) {
                            const double b = xb[j * 8 * P + k * 4];
                            dmma884(d[j][0], d[j][1], a[k], b);
                        }
                    }
                } else if (resident) {
                    // shared row block or partial group: this warp's column blocks, one unpredicated chain each
                    for (int jj = myB; jj < nbAct; jj += bp) {
                        double e0 = 0.0, e1 = 0.0;     // k ascending: the same sums as every other path
                        const double* xs = xb + jj * 8 * P;
#pragma unroll
                        for (int k = 0; k < KSREG; k++) {
                            if (k >= ksEff) break;
                            dmma884(e0, e1, a[k], xs[k * 4]);
                        }
#pragma unroll
                        for (int j = 0; j < NB; j++) if (j == jj) { d[j][0] = e0; d[j][1] = e1; }
                    }
                } else if (STREAM) {
                    // wide-K layers: the A fragments come from the TMA-staged image in shared memory (or, when it does not fit, from L2)
                    const double* As = Astage ? Astage + (size_t)(active ? rb : 0) * KS * 32 + lane : nullptr;
                    for (int k = 0; k < ksEff; k++) {
                        double av = As ? As[k * 32] : __ldg(Af + k * 32);
#pragma unroll
                        for (int j = 0; j < NB; j++) {
                            if (j < nbAct && ((blkMask >> j) & 1u)) {
                                double b = xb[j * 8 * P + k * 4];
                                dmma884(d[j][0], d[j][1], av, b);
                            }
                        }
                    }
                }
            }
            // target columns of this thread's results: looked up before the barrier so that the wait hides the latency
            unsigned pos[NB][2];
#pragma unroll
            for (int j = 0; j < NB; j++) {
                // this thread's two columns are neighbours: one 32-bit load for both positions
                const int col0 = cg * NB * 8 + j * 8 + qc * 2;
                unsigned pp = 0xFFFFFFFFu;
                if (mine && col0 < nCols && ((blkMask >> j) & 1u)) pp = *reinterpret_cast<const unsigned*>(colPos + col0);
                pos[j][0] = pp & 0xFFFFu;
                pos[j][1] = col0 + 1 < nCols ? pp >> 16 : 0xFFFFu;
            }
            if (inplace) {
                // everybody has read group cg (and all groups behind it): only then may its results, which land
                // at or behind column 32 cg, be written
                // (a hardware named barrier: ~10x lower latency than an mbarrier phase with try_wait here)
                asm volatile("bar.sync 1, %0;\n" ::"r"(nWarps * 32) : "memory");
            }
            if (mine && row < rowsW) {
#pragma unroll
                for (int j = 0; j < NB; j++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        if (pos[j][e] != 0xFFFFu) {
                            double v = d[j][e];
                            if (addBias && (pos[j][e] & 0x8000u)) v += bias;
                            Y[(pos[j][e] & 0x7FFFu) * P + row] = v;
                        }
                    }
                }
            }
        }
    }
}

//@phase: relaxation
__device__ __forceinline__ double dz_sigm(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double dz_sigm_p(double x) { double ex = exp(x); return ex / ((1.0 + ex) * (1.0 + ex)); }
__device__ __forceinline__ double dz_tanh_p(double x) { double t = tanh(x); return 1.0 - t * t; }

// Column layout of a tile from per-cell widths, computed redundantly by every warp (no serial
// section): lane i holds the widths of cells i and i + 32 and gets their offsets back;
// `keep` = number of leading cells that fit into capCols columns, `total` their columns.
//@phase: layout
__device__ __forceinline__ void dz_layout(int w0, int w1, int nCells, int capCols, int lane, int& off0, int& off1,
                                          int& keep, int& total) {
    int s0 = w0, s1 = w1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t0 = __shfl_up_sync(0xffffffffu, s0, d);
        int t1 = __shfl_up_sync(0xffffffffu, s1, d);
        if (lane >= d) { s0 += t0; s1 += t1; }
    }
    s1 += __shfl_sync(0xffffffffu, s0, 31);
    const bool f0 = lane < nCells && s0 + 8 <= capCols;
    const bool f1 = lane + 32 < nCells && s1 + 8 <= capCols;
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

// remove_zero_columns bookkeeping: the reference drops all-zero generators after every non-Id
// activation (and after a projection); columns are never physically removed here (a zero column
// stays zero and contributes nothing), they are only marked so that generator counts, statistics
// and fetched zonotopes agree with the reference.  Marks the used generator columns of the
// current layout that are identically zero.
//@phase: layer loop head / stats
__device__ __forceinline__ void dz_mark_dead(unsigned* dead, const int* colOff, const unsigned short* wUsed, const double* X,
                                             int P, int rows, int nCells, int tid, int NT) {
    for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) dead[i] = 0;
    __syncthreads();
    for (int ci = tid >> 5; ci < nCells; ci += NT >> 5) {
        const int cb = colOff[ci], w = wUsed[ci];
        for (int j = 1 + (tid & 31); j < w; j += 32) {
            bool any = false;
            for (int r = 0; r < rows && !any; r++) any = X[(cb + j) * P + r] != 0.0;
            if (!any) atomicOr(&dead[(cb + j) >> 5], 1u << ((cb + j) & 31));
        }
    }
    __syncthreads();
}

//@phase: relaxation
// The activation rule of one neuron: bounds -> (new centre, slope, new generator); true when a generator is appended.
template <bool FULL>
__device__ __forceinline__ bool dz_rule(int act, double c, double lo, double hi, double& cn, double& la, double& mm) {
    cn = c; la = 1.0; mm = 0.0;
    if (act == CLR_ACT_RELU) {
        if (dz_leq(lo, 0.0)) {
            if (dz_leq(hi, 0.0)) { cn = 0.0; la = 0.0; }
            else {
                la = hi / (hi - lo);
                mm = -la * lo / 2;
                cn = c * la + mm;
                return true;
            }
        }
    } else if (FULL && (act == CLR_ACT_SIGMOID || act == CLR_ACT_TANH)) {
        // the transcendental calls side by side, so that their (independent) instruction chains interleave;
        // the same formulas and values as dz_sigm / dz_sigm_p / dz_tanh_p
        double ly, uy, dl, du;
        if (act == CLR_ACT_SIGMOID) {
            const double eml = exp(-lo), emh = exp(-hi), el = exp(lo), eh = exp(hi);
            ly = 1.0 / (1.0 + eml); uy = 1.0 / (1.0 + emh);
            dl = el / ((1.0 + el) * (1.0 + el)); du = eh / ((1.0 + eh) * (1.0 + eh));
        } else {
            ly = tanh(lo); uy = tanh(hi);
            dl = 1.0 - ly * ly; du = 1.0 - uy * uy;
        }
        if (dz_approx(lo, hi)) { cn = uy; la = 0.0; }
        else {
            la = dl < du ? dl : du;
            double mu1 = (uy + ly - la * (hi + lo)) / 2;
            mm = (uy - ly - la * (hi - lo)) / 2;
            cn = c * la + mu1;
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// The persistent tile kernel.
// ---------------------------------------------------------------------------------------------
// FULL = false: the lean instantiation for ReLU / Id networks whose layers all keep their W fragments in registers -- no
// sigmoid / tanh code and no streamed (K > 4 KSREG) products.  The kernel is sensitive to its own size (registers live
// across the relaxation, instruction fetch): the lean variant is 5 % faster on the same work.
//@phase: setup / tile fetch
// WIDE = true adds the several-threads-per-neuron relaxation of tiles that hold one or two wide cells (see the relaxation).
template <int KSREG, bool GLOBAL, bool STATS, bool FULL, bool WIDE>
__global__ void __launch_bounds__(DZ_THREADS, DZ_CTAS)
deepz_kernel(const DevNet* __restrict__ net, DeepzInput in, DeepzOutput out, DevCallState* st, DeepzSched sp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DeepzSmem& S = *reinterpret_cast<DeepzSmem*>(smem_raw);
    double* store = GLOBAL ? sp.workspace + (size_t)blockIdx.x * sp.cap
                           : reinterpret_cast<double*>(smem_raw + sp.store_off);
    // rows whose relaxation appended a generator, [pending | being written] x cell x word, and the position table
    // (column of the current layout -> column of the next one; 0x8000: centre, 0xFFFF: unused), sized by the host
    unsigned* const maskBase = reinterpret_cast<unsigned*>(smem_raw + dz_meta_bytes(STATS) + (FULL ? sp.astage_bytes : 0));
    const int mS = sp.mask_stride, mWords = sp.mask_words, maskHalf = sp.tile_max * sp.mask_stride;
    unsigned short* const colPosT = reinterpret_cast<unsigned short*>(maskBase + 2 * maskHalf);
    const int tid = threadIdx.x, NT = DZ_THREADS;
    const int lane = tid & 31, warp = tid >> 5, nWarps = NT >> 5;
    const int L = net->L;
    const int P = net->max_pitch;                 // one column pitch for the whole network
    const int mOut = net->m_out;
    const int n = in.n;
    const bool inplace = FULL ? sp.inplace != 0 : true;      // lean variant: every warp owns at most one row block
    const bool useHull = out.hull_lo != nullptr && mOut <= CLR_HULL_MAX;
    const bool needDead = STATS;                  // kept zonotopes are only written by the STATS instantiations
    const int stageW = max(clr_pad4(net->max_rows), 2 * n);   // doubles of staging per cell
    // A final Id layer with at most 4 rows (the scalar control of most models, or the folded post-processing) does not get a
    // product / relaxation / layout cycle of its own: it is a dot product per column, applied in the output stage.
    const bool fuse = sp.fuse_tail != 0;
    const int Lrun = fuse ? L - 1 : L;

    // ---- one-time setup: layer descriptors to smem, accumulators, the first tile ------------------
    for (int i = tid; i < L * (int)(sizeof(DevLayer) / sizeof(int)); i += NT) ((int*)S.layers)[i] = ((const int*)net->layer)[i];
    double* const astage = (FULL && sp.astage_bytes > 0) ? reinterpret_cast<double*>(smem_raw + sp.astage_off) : nullptr;
    unsigned aphase = 0;                          // parity of the staging barrier (every thread keeps its own copy)
    if (FULL && astage && tid == 0) dz_mbar_init(&S.abar);
    if (tid < CLR_MAX_LAYERS) { S.sPin[tid] = 0; S.sUnst[tid] = 0; }
    if (tid < CLR_HULL_MAX) { S.hlo[tid] = INFINITY; S.hhi[tid] = -INFINITY; }
    const unsigned long long total_units =
        sp.pass == 0 ? (unsigned long long)in.N : (sp.pass == 1 ? st->n_retry : st->n_big);
    unsigned long long* cursor = sp.pass == 0 ? &st->next_cell : (sp.pass == 1 ? &st->next_retry : &st->next_big);
    const int tileMax = min(min(CLR_TILE_CELLS_MAX, sp.tile_max), max(1, (sp.cap / 8) / stageW));   // staging takes 1/8 of the store at most
    unsigned long long pfFirst = 0;   // thread 0: prefetched tile
    int pfT = 0;
    if (tid == 0) {
        int t0 = sp.tile_cells;
        if (t0 <= 0) {
            int guessCols = in.kind == 0 ? in.n + 1 : 8;
            t0 = (sp.cap / (inplace ? 1 : 2)) / ((guessCols * 2 + 8 + sp.grow_cols) * P);
            if (sp.pass != 0) t0 = max(1, t0 / 4);
            // small calls (the whole call fits one tile per CTA): exactly one tile each -- a tile costs its latency chain
            // whatever its size; otherwise two tiles per CTA and round, so that the tail balances
            const unsigned long long one = (total_units + gridDim.x - 1) / gridDim.x;
            unsigned long long fair = one <= (unsigned long long)t0 ? one : (total_units + 2ull * gridDim.x - 1) / (2ull * gridDim.x);
            if ((unsigned long long)t0 > fair) t0 = (int)(fair < 1 ? 1 : fair);
        }
        if (sp.pass == 2) t0 = 1;
        t0 = max(1, min(t0, tileMax));
        S.tileT = t0;
        S.sNear = 0; S.sMaxP = 0;
        pfT = t0;
        pfFirst = atomicAdd(cursor, (unsigned long long)pfT);
    }
    __syncthreads();
#ifdef DZ_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif
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
            S.peakCols = 0.f;
            if (nc > 0) {
                int T = S.tileT;
                if (sp.tile_cells <= 0 && sp.pass != 2) {
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
        DZ_T(6);
        if (!S.more) break;
        int nCells = S.nCells;
        // store geometry of this tile: [columns ... | staging (nCells x stageW)]
        const int stageOff = sp.cap - nCells * stageW;
        double* stage = store + stageOff;
        const int region = inplace ? stageOff : stageOff / 2;
        const int capCols = min(region / P, sp.ncol_max);
        double* cur = store;
        double* oth = inplace ? store : store + region;
        if (tid < nCells) {
            long long idx = S.firstIdx + tid;
            S.cellId[tid] = sp.pass == 0 ? idx : (sp.pass == 1 ? sp.retry_list[idx] : sp.big_list[idx]);
            if (STATS) S.cellNear[tid] = 0;
        }
        for (int i = tid; i < nCells * mS; i += NT) maskBase[i] = 0;
        if (needDead) for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) { S.deadMask[0][i] = 0; S.deadMask[1][i] = 0; }
        __syncthreads();

//@phase: input + first layer
        // ---- input zonotopes -> column store -------------------------------------------------------------------
        int flip = 0, flipD = 0;
        int nCols;
        double* tmpC = stage;                     // box grid: per-cell centre / radius scratch
        double* tmpR = stage + nCells * n;
        if (in.kind == 0) {
            for (int it = tid; it < nCells * n; it += NT) {
                const int ci = it / n, d = it - ci * n;
                const long long li = S.cellId[ci];
                const unsigned long long g = (unsigned long long)(in.cyc_blk > 0 ? in.cell_begin + (li / in.cyc_blk) * in.cyc_stride + li % in.cyc_blk
                                                                                   : in.cell_begin + li);
                const unsigned long long sd = (unsigned long long)in.stride[d];   // can exceed 2^32 even when g does not
                int k;
                if (sd > g) k = 0;
                else if ((g >> 32) == 0) k = (int)(((unsigned)g / (unsigned)sd) % (unsigned)in.part[d]);
                else k = (int)((g / sd) % (unsigned long long)in.part[d]);
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
                        p = in.col_off ? (int)(in.col_off[cell + 1] - in.col_off[cell]) : in.ncols[cell];
                    }
                    w[h] = 1 + p;
                }
            }
            int off0, off1, keep, total;
            const int wu0 = w[0], wu1 = w[1];
            if (S.layers[0].collapse) { if (lane < nCells) w[0] = max(w[0], 2); if (lane + 32 < nCells) w[1] = max(w[1], 2); }
            dz_layout(w[0], w[1], nCells, capCols, lane, off0, off1, keep, total);
            if (warp == 0) {
                // with no pending generators the next layout equals the current one
                if (lane < keep) { S.colOff[0][lane] = off0; S.colOff[1][lane] = off0; S.wUsed[0][lane] = (unsigned short)wu0; S.uPend[0][lane] = 0; }
                if (lane + 32 < keep) { S.colOff[0][lane + 32] = off1; S.colOff[1][lane + 32] = off1; S.wUsed[0][lane + 32] = (unsigned short)wu1; S.uPend[0][lane + 32] = 0; }
                if (lane == 0) {
                    S.colOff[0][keep] = total; S.colOff[1][keep] = total; S.nItems = 0;
                    S.kAct = in.kind == 0 ? S.layers[0].m : n;
                    if (keep < nCells) dz_defer(sp, st, S.cellId, keep, nCells);
                    if (keep > 0) { S.peakNeed = (float)(total + 8) / (float)keep; S.peakCols = (float)total / (float)keep; }
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
                // Box cells: the generators are r_d * e_d, so the first layer needs no tensor product: its output is
                // W c (summed over d in ascending order, like the reference's affine map) and the columns r_d * W[:, d].
                const DevLayer& L0 = S.layers[0];
                const int m0 = L0.m, rowsP0 = clr_pad4(m0);
                const bool bias0 = L0.act == CLR_ACT_ID && !L0.collapse;     // no relaxation follows: bias goes in here
                double* dst = inplace ? cur : oth;
                const int G0 = NT / rowsP0, slot = tid / rowsP0, r = tid - slot * rowsP0;
                if (slot < G0) {
                    // this thread's row of W (the same for every cell): fetched once, ahead of the cell loop
                    double wr[8];
#pragma unroll
                    for (int d = 0; d < 8; d++) wr[d] = (d < n && r < m0) ? __ldg(L0.Wcm + (size_t)d * m0 + r) : 0.0;
                    const double b0 = (bias0 && r < m0) ? __ldg(L0.bias + r) : 0.0;
                    for (int ci = slot; ci < nCells; ci += G0) {
                        const int cb = colOff[ci];
                        double* y = dst + cb * P + r;
                        if (r < m0 && n <= 8) {
                            const double* cc = tmpC + ci * n;
                            const double* rr = tmpR + ci * n;
                            double cv[8], rv[8];
#pragma unroll
                            for (int d = 0; d < 8; d++) { cv[d] = d < n ? cc[d] : 0.0; rv[d] = d < n ? rr[d] : 0.0; }
                            double sacc = 0.0;
                            int j = 1;
#pragma unroll
                            for (int d = 0; d < 8; d++) {
                                if (d < n) {
                                    sacc += wr[d] * cv[d];
                                    if (rv[d] != 0.0) { y[j * P] = wr[d] * rv[d]; j++; }
                                }
                            }
                            y[0] = bias0 ? sacc + b0 : sacc;
                        } else if (r < m0) {
                            const double* cc = tmpC + ci * n;
                            const double* rr = tmpR + ci * n;
                            double sacc = 0.0;
                            int j = 1;
                            for (int d = 0; d < n; d++) {
                                const double wv = __ldg(L0.Wcm + (size_t)d * m0 + r);
                                sacc += wv * cc[d];
                                const double rd = rr[d];
                                if (rd != 0.0) { y[j * P] = wv * rd; j++; }
                            }
                            y[0] = bias0 ? sacc + __ldg(L0.bias + r) : sacc;
                        } else {
                            const int w = S.wUsed[0][ci];
                            for (int j = 0; j < w; j++) y[j * P] = 0.0;
                        }
                    }
                }
            } else {
                for (int it = tid; it < nCells * rows4; it += NT) {
                    const int ci = it / rows4, d = it - ci * rows4;
                    const int cb = colOff[ci], p = S.wUsed[0][ci] - 1;
                    const long long cell = S.cellId[ci];
                    const long long go = in.col_off ? in.col_off[cell] : cell * (long long)in.pad_cap;
                    cur[cb * P + d] = d < n ? in.C[cell * n + d] : 0.0;
                    for (int j = 0; j < p; j++) cur[(cb + 1 + j) * P + d] = d < n ? in.G[(go + j) * n + d] : 0.0;
                }
            }
            for (int ci = warp; ci < nCells; ci += nWarps) {
                const int cb = colOff[ci], w = S.wUsed[0][ci], wa = colOff[ci + 1] - cb;
                for (int j = lane; j < wa; j += 32)
                    colPosT[cb + j] = j < w ? (unsigned short)((cb + j) | (j == 0 ? 0x8000 : 0)) : (unsigned short)0xFFFF;
            }
            __syncthreads();   // input store and position table complete
        }
        DZ_T(0);

//@phase: layer loop head / stats
        // ---- layers ------------------------------------------------------------------------------------------------
        // invariant at the top (parity `flip`): data in layout colOff[flip] with wUsed[flip] columns per cell, generators
        // appended by the previous relaxation pending in mask[flip] / stage, next layout in colOff[flip ^ 1], colPos maps
        // current -> next columns
        bool tileDead = false;
        for (int l = 0; l < Lrun; l++) {
            const DevLayer& Ly = S.layers[l];
            const int* offNxt = S.colOff[flip ^ 1];
            const unsigned short* wCur = S.wUsed[flip];
            const unsigned* pend = maskBase + flip * maskHalf;
            unsigned* fresh = maskBase + (flip ^ 1) * maskHalf;
            if (STATS && Ly.net_layer >= 0 && tid < nCells) {
                // generators entering this layer, as the reference counts them
                const int cb = S.colOff[flip][tid], w = wCur[tid];
                int nz = 0;
                for (int j = 1; j < w; j++) nz += !((S.deadMask[flipD][(cb + j) >> 5] >> ((cb + j) & 31)) & 1u);
                for (int ww = 0; ww < mWords; ww++) nz += __popc(pend[tid * mS + ww]);
                S.cellPin[tid][Ly.net_layer] = (unsigned short)nz;
                S.cellUnst[tid][Ly.net_layer] = 0;
            }
            const int act = Ly.act, collapse = Ly.collapse;
            const bool relax = !(act == CLR_ACT_ID && !collapse);
            const int m = Ly.m, rowsP = clr_pad4(m), mw = (m + 31) >> 5;
            if (!(l == 0 && in.kind == 0)) {  // box cells enter with the first layer's product already applied
                const double* As = nullptr;
                if (FULL && astage && Ly.ksteps > KSREG && (unsigned)(Ly.rblocks * Ly.ksteps * 256) <= (unsigned)sp.astage_bytes) {
                    // every warp is past the barrier that ended the previous product: the staging area is free
                    if (tid == 0) dz_tma_load(astage, Ly.Afrag, (unsigned)(Ly.rblocks * Ly.ksteps * 256), &S.abar);
                    dz_mbar_wait(&S.abar, aphase);
                    aphase ^= 1u;
                    As = astage;
                }
                dz_gemm<KSREG, FULL>(Ly, cur, oth, P, nCols, S.kAct, colPosT, !relax, inplace, warp, lane, nWarps, a, As);
            }
            DZ_T(1);
//@phase: appended columns
            // Pending generators mu * e_i of the previous relaxation: their image is mu * W[:, i].  In place this is safe
            // without another barrier: after the last group's barrier of the sweep every warp has read every column.
            if (l > 0) {
                const int nIt = S.nItems;
                if (nIt <= DZ_ITEMS_MAX) {
                    if (rowsP <= 128) {
                        // all loads of a column (and of the warp's next item) in flight before the stores
                        for (int t = warp; t < nIt; t += 2 * nWarps) {
                            const int t2 = t + nWarps;
                            const bool two = t2 < nIt;
                            const int i0 = S.itemRow[t], c0 = S.itemCol[t];
                            const int i1 = two ? S.itemRow[t2] : i0, c1 = two ? S.itemCol[t2] : c0;
                            const double mu0 = stage[S.itemCell[t] * stageW + i0];
                            const double mu1 = two ? stage[S.itemCell[t2] * stageW + i1] : 0.0;
                            const double* w0 = Ly.Wcm + (size_t)i0 * m;
                            const double* w1 = Ly.Wcm + (size_t)i1 * m;
                            double v0[4], v1[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const int r = lane + 32 * q;
                                v0[q] = r < m ? __ldg(w0 + r) : 0.0;
                                v1[q] = (two && r < m) ? __ldg(w1 + r) : 0.0;
                            }
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const int r = lane + 32 * q;
                                if (r < rowsP) {
                                    oth[c0 * P + r] = mu0 * v0[q];
                                    if (two) oth[c1 * P + r] = mu1 * v1[q];
                                }
                            }
                        }
                    } else {
                        for (int t = warp; t < nIt; t += nWarps) {
                            const int i = S.itemRow[t], col = S.itemCol[t];
                            const double mu = stage[S.itemCell[t] * stageW + i];
                            const double* wc = Ly.Wcm + (size_t)i * m;
                            for (int r = lane; r < rowsP; r += 32) oth[col * P + r] = r < m ? mu * __ldg(wc + r) : 0.0;
                        }
                    }
                } else {
                    const int mwPrev = (S.layers[l - 1].m + 31) >> 5;
                    int t = 0;
                    for (int ci = 0; ci < nCells; ci++) {
                        int col = offNxt[ci] + wCur[ci];
                        for (int ww = 0; ww < mwPrev; ww++) {
                            unsigned bits = pend[ci * mS + ww];
                            while (bits) {
                                const int i = ww * 32 + __ffs(bits) - 1;
                                bits &= bits - 1;
                                if ((t & (nWarps - 1)) == warp) {
                                    const double mu = stage[ci * stageW + i];
                                    const double* wc = Ly.Wcm + (size_t)i * m;
                                    for (int r = lane; r < rowsP; r += 32) oth[col * P + r] = r < m ? mu * __ldg(wc + r) : 0.0;
                                }
                                t++; col++;
                            }
                        }
                    }
                }
            }
            for (int i = tid; i < nCells * mS; i += NT) fresh[i] = 0;
            if (tid == 0) S.kActNew = 0;
            // weights of the next layer (or of the first layer for the next tile) while the relaxation runs
            dz_load_a<KSREG>(S.layers[l + 1 < Lrun ? l + 1 : 0], warp, lane, nWarps, a);
            if (needDead) {
                // dead marks follow their columns to the new positions
                unsigned* dOld = S.deadMask[flipD];
                unsigned* dNew = S.deadMask[flipD ^ 1];
                for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) dNew[i] = 0;
                __syncthreads();
                for (int col = tid; col < nCols; col += NT) {
                    const unsigned pos = colPosT[col];
                    if (pos != 0xFFFFu && ((dOld[col >> 5] >> (col & 31)) & 1u)) atomicOr(&dNew[(pos & 0x7FFFu) >> 5], 1u << (pos & 31));
                }
                flipD ^= 1;
            }
            { double* tp = cur; cur = oth; oth = tp; }
            __syncthreads();   // [1] all columns moved to layout colOff[flip ^ 1]; pending generators materialised
            DZ_T(2);
            const int* colOff = S.colOff[flip ^ 1];

//@phase: relaxation
            // ---- relaxation: bounds, lambda, mu, rows scaled in place; thread = (cell slot, row) ------------------------
            // The radius of a neuron is DEFINED chunk-wise: rad = ((S_0 + S_1) + S_2) + ..., S_c the ascending sum of |g_j| over
            // the generators 64 c + 1 .. 64 c + 64 -- for cells of up to 64 generators LazySets' plain ascending sum, for wider
            // cells the same terms grouped by 64 (a difference of rounding only, ~1e-16).  The definition does not depend on how
            // a tile is relaxed, so a cell gets the same bits in every tile, pass and instantiation.
            if (relax) {
                unsigned rowLive = 0;                     // 1 + the highest row this thread left non-zero
                const int G = NT / rowsP;                 // cell slots processed side by side
                const int slot = tid / rowsP, r = tid - slot * rowsP;
                // Tiles of one or two wide cells (the dense regime, sigmoid / tanh controllers) would leave most cell slots idle
                // behind a chain of hundreds of dependent additions: there T = G / nCells threads share a neuron, one chunk of 64
                // generators at a time; the partial sums go through the 8 slack columns behind the layout.
                int T = 1;
                if (WIDE && nCells * 2 <= G) {
                    int wmax = 0;
                    for (int ci = 0; ci < nCells; ci++) wmax = max(wmax, (int)wCur[ci] + (int)S.uPend[flip][ci]);
                    if (wmax > 65 && ((wmax + 62) >> 6) * nCells <= 8) T = G / nCells;
                }
                if (WIDE && T > 1) {
                    const int chCap = 8 / nCells;             // chunk slots per cell in the slack columns
                    const int ci = slot / T, sub = slot - ci * T;
                    const bool on = ci < nCells && r < m;
                    double* part = cur + (size_t)colOff[nCells] * P;
                    int w = 0, nch = 0;
                    double* y = cur;
                    double c = 0.0;
                    if (on) {
                        w = wCur[ci] + S.uPend[flip][ci];
                        nch = (w + 62) >> 6;
                        y = cur + colOff[ci] * P + r;
                        c = y[0] + __ldg(Ly.bias + r);
                        for (int ch = sub; ch < nch; ch += T) {
                            const int j1 = min(w, 65 + 64 * ch);
                            double sacc = 0.0;
                            for (int j0 = 1 + 64 * ch; j0 < j1; j0 += 8) {
                                double t[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) t[q] = j0 + q < j1 ? y[(j0 + q) * P] : 0.0;
#pragma unroll
                                for (int q = 0; q < 8; q++) sacc += fabs(t[q]);
                            }
                            part[(ci * chCap + ch) * P + r] = sacc;
                        }
                    }
                    __syncthreads();   // [2a] partial sums complete; every centre has been read
                    // ReLU: all T threads of a neuron evaluate the (cheap) rule themselves; sigmoid / tanh (four transcendental
                    // calls): thread 0 of the neuron does and hands the slope over through the neuron's first partial-sum slot
                    const bool shareRule = FULL && (act == CLR_ACT_SIGMOID || act == CLR_ACT_TANH);
                    double la = 1.0;
                    if (on && (!shareRule || sub == 0)) {
                        double rad = nch > 0 ? part[(ci * chCap) * P + r] : 0.0;
                        for (int ch = 1; ch < nch; ch++) rad += part[(ci * chCap + ch) * P + r];
                        const double lo = c - rad, hi = c + rad;
                        double cn, mm;
                        const bool unst = dz_rule<FULL>(act, c, lo, hi, cn, la, mm);     // the same values in all T threads of a neuron
                        if (sub == 0) {
                            if (STATS && act == CLR_ACT_RELU && (fabs(lo) < 1e-12 || fabs(hi) < 1e-12)) atomicAdd(&S.cellNear[ci], 1u);
                            if (cn != 0.0 || la != 0.0) {
                                rowLive = r + 1;
                                if (out.liveCount) atomicAdd(&out.liveCount[l * CLR_MAX_ROWS + r], 1);
                            }
                            y[0] = cn;
                            if (collapse) y[P] = la * rad;
                            if (unst) {
                                stage[ci * stageW + r] = mm;
                                atomicOr(&fresh[ci * mS + (r >> 5)], 1u << (r & 31));
                            }
                            if (shareRule) part[(ci * chCap) * P + r] = la;
                        }
                    }
                    if (shareRule) {
                        __syncthreads();   // [2b]
                        if (on && sub != 0) la = part[(ci * chCap) * P + r];
                    }
                    if (on && !collapse && la != 1.0) {
                        for (int ch = sub; ch < nch; ch += T) {
                            const int j1 = min(w, 65 + 64 * ch);
                            if (la == 0.0) { for (int j = 1 + 64 * ch; j < j1; j++) y[j * P] = 0.0; }
                            else {
                                for (int j0 = 1 + 64 * ch; j0 < j1; j0 += 8) {
                                    double t[8];
#pragma unroll
                                    for (int q = 0; q < 8; q++) t[q] = j0 + q < j1 ? y[(j0 + q) * P] : 0.0;
#pragma unroll
                                    for (int q = 0; q < 8; q++) if (j0 + q < j1) y[(j0 + q) * P] = t[q] * la;
                                }
                            }
                        }
                    }
                } else if (slot < G && r < m) {
                    const double bias = __ldg(Ly.bias + r);
                    for (int ci = slot; ci < nCells; ci += G) {
                        const int w = wCur[ci] + S.uPend[flip][ci];
                        const int cb = colOff[ci];
                        double* y = cur + cb * P + r;
                        // the row is read in groups of 8 generators, all loads of a group in flight before the sequential
                        // (ascending, like LazySets' loops over the generators) sum; the first group stays in registers for the
                        // scaling below.  Bounds: c -/+ rad with ONE radius sum, for one-row layers and all others alike.
                        double v[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) v[q] = q + 1 < w ? y[(q + 1) * P] : 0.0;
                        const double c = y[0] + bias;
                        double rad = 0.0;
#pragma unroll
                        for (int q = 0; q < 8; q++) rad += fabs(v[q]);       // + 0.0 beyond w: exact
                        const int wA = min(w, 65);
                        for (int j0 = 9; j0 < wA; j0 += 8) {
                            double t[8];
#pragma unroll
                            for (int q = 0; q < 8; q++) t[q] = j0 + q < w ? y[(j0 + q) * P] : 0.0;
#pragma unroll
                            for (int q = 0; q < 8; q++) rad += fabs(t[q]);
                        }
                        for (int jc = 65; jc < w; jc += 64) {            // further chunks of 64 generators (wide cells only)
                            const int j1 = min(w, jc + 64);
                            double sacc = 0.0;
                            for (int j0 = jc; j0 < j1; j0 += 8) {
                                double t[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) t[q] = j0 + q < j1 ? y[(j0 + q) * P] : 0.0;
#pragma unroll
                                for (int q = 0; q < 8; q++) sacc += fabs(t[q]);
                            }
                            rad += sacc;
                        }
                        const double lo = c - rad, hi = c + rad;
                        if (STATS && act == CLR_ACT_RELU && (fabs(lo) < 1e-12 || fabs(hi) < 1e-12)) atomicAdd(&S.cellNear[ci], 1u);
                        double cn, la, mm;
                        const bool unst = dz_rule<FULL>(act, c, lo, hi, cn, la, mm);
                        if (cn != 0.0 || la != 0.0) {
                            rowLive = r + 1;
                            if (out.liveCount) atomicAdd(&out.liveCount[l * CLR_MAX_ROWS + r], 1);
                        }
                        y[0] = cn;
                        if (collapse) y[P] = la * rad;       // LazySets' one-row linear_map: a single generator
                        else if (la == 0.0) {                                                   // inactive row: exactly zero, no reads
#pragma unroll
                            for (int q = 0; q < 8; q++) if (q + 1 < w) y[(q + 1) * P] = 0.0;
                            for (int j = 9; j < w; j++) y[j * P] = 0.0;
                        }
                        else if (la != 1.0) {
#pragma unroll
                            for (int q = 0; q < 8; q++) if (q + 1 < w) y[(q + 1) * P] = v[q] * la;
                            for (int j0 = 9; j0 < w; j0 += 8) {
                                double t[8];
#pragma unroll
                                for (int q = 0; q < 8; q++) t[q] = j0 + q < w ? y[(j0 + q) * P] : 0.0;
#pragma unroll
                                for (int q = 0; q < 8; q++) if (j0 + q < w) y[(j0 + q) * P] = t[q] * la;
                            }
                        }
                        if (unst) {
                            stage[ci * stageW + r] = mm;
                            atomicOr(&fresh[ci * mS + (r >> 5)], 1u << (r & 31));
                        }
                    }
                }
                rowLive = __reduce_max_sync(0xffffffffu, rowLive);
                if (lane == 0 && rowLive) atomicMax(&S.kActNew, (int)rowLive);
                __syncthreads();   // [2]
            } else if (tid == 0) {
                S.kActNew = m;                            // no relaxation: every row may be non-zero
            }
            DZ_T(3);
//@phase: layout
            // ---- next layout (every warp computes it; warp 0 publishes it) and the position table -----------------------
            if (l + 1 == Lrun) {
                // no product follows the last layer: only the widths the output stage reads -- and, when the last layer of the
                // network is applied in the output stage (fuse), the list of the generators just appended (cell, row, rank)
                int uu[2] = {0, 0};
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int ci = lane + 32 * h;
                    if (ci < nCells) for (int ww = 0; ww < mw; ww++) uu[h] += __popc(fresh[ci * mS + ww]);
                }
                int s0 = uu[0], s1 = uu[1];                  // inclusive scan over the cells: first item of every cell
                if (fuse) {
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const int t0 = __shfl_up_sync(0xffffffffu, s0, dd);
                        const int t1 = __shfl_up_sync(0xffffffffu, s1, dd);
                        if (lane >= dd) { s0 += t0; s1 += t1; }
                    }
                    s1 += __shfl_sync(0xffffffffu, s0, 31);
                }
                const int ub0 = s0 - uu[0], ub1 = s1 - uu[1];
                if (warp == 0) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int ci = lane + 32 * h;
                        if (ci < nCells) {
                            S.wUsed[flip ^ 1][ci] = (unsigned short)(collapse ? 2 : wCur[ci] + S.uPend[flip][ci]);
                            S.uPend[flip ^ 1][ci] = (unsigned short)uu[h];
                            if (fuse) S.uPend[flip][ci] = (unsigned short)(h ? ub1 : ub0);     // the array just read is free: item base of the cell
                            if (STATS && Ly.net_layer >= 0) S.cellUnst[ci][Ly.net_layer] = (unsigned short)uu[h];
                        }
                    }
                }
                if (tid == 0) S.kAct = S.kActNew;
                if (fuse) {
                    const int nIt = __shfl_sync(0xffffffffu, nCells <= 32 ? s0 : s1, (nCells - 1) & 31);
                    if (tid == 0) S.nItems = nIt;
                    if (nIt <= DZ_ITEMS_MAX) {
                        for (int ci = warp; ci < nCells; ci += nWarps) {
                            const int base = ci < 32 ? __shfl_sync(0xffffffffu, ub0, ci & 31) : __shfl_sync(0xffffffffu, ub1, ci & 31);
                            int before = 0;
                            for (int ww = 0; ww < mw; ww++) {
                                const unsigned bits = fresh[ci * mS + ww];
                                if ((bits >> lane) & 1u) {
                                    const int k = before + __popc(bits & ((1u << lane) - 1u));
                                    S.itemCell[base + k] = (unsigned short)ci;
                                    S.itemRow[base + k] = (unsigned short)(ww * 32 + lane);
                                }
                                before += __popc(bits);
                            }
                        }
                    }
                }
            } else {
                int off0, off1, keep, total;
                int w[2] = {0, 0}, wu[2] = {0, 0}, uu[2] = {0, 0};
                const int minW = (l + 1 < Lrun && S.layers[l + 1].collapse) ? 2 : 1;   // a one-row map writes [c, r]
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int ci = lane + 32 * h;
                    if (ci < nCells) {
                        int u = 0;
                        for (int ww = 0; ww < mw; ww++) u += __popc(fresh[ci * mS + ww]);
                        uu[h] = u;
                        wu[h] = collapse ? 2 : wCur[ci] + S.uPend[flip][ci];
                        w[h] = max(max(colOff[ci + 1] - colOff[ci], wu[h] + u), minW);   // cells never shrink: they only move right
                        if (STATS && warp == 0 && Ly.net_layer >= 0) S.cellUnst[ci][Ly.net_layer] = (unsigned short)u;
                    }
                }
                // one scan for both the column offsets (low 16 bits) and the item offsets (high 16 bits)
                int ub0, ub1, nIt;
                {
                    int s0 = w[0] | (uu[0] << 16), s1 = w[1] | (uu[1] << 16);
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const int t0 = __shfl_up_sync(0xffffffffu, s0, dd);
                        const int t1 = __shfl_up_sync(0xffffffffu, s1, dd);
                        if (lane >= dd) { s0 += t0; s1 += t1; }
                    }
                    s1 += __shfl_sync(0xffffffffu, s0, 31);
                    const int e0 = s0 & 0xFFFF, e1 = s1 & 0xFFFF;                     // inclusive column ends
                    const bool f0 = lane < nCells && e0 + 8 <= capCols;
                    const bool f1 = lane + 32 < nCells && e1 + 8 <= capCols;
                    keep = __popc(__ballot_sync(0xffffffffu, f0)) + __popc(__ballot_sync(0xffffffffu, f1));
                    off0 = e0 - w[0]; off1 = e1 - w[1];
                    ub0 = (s0 >> 16) - uu[0]; ub1 = (s1 >> 16) - uu[1];               // first item of cells lane, lane + 32
                    const int l0 = __shfl_sync(0xffffffffu, s0, (keep - 1) & 31);
                    const int l1 = __shfl_sync(0xffffffffu, s1, (keep - 33) & 31);
                    const int last = keep == 0 ? 0 : (keep <= 32 ? l0 : l1);
                    total = last & 0xFFFF;
                    nIt = last >> 16;                                                 // generators appended in the kept cells
                }
                int* cn = S.colOff[flip];                // the layout the data just left is free: it becomes the next one
                if (warp == 0) {
                    if (lane < keep) cn[lane] = off0;
                    if (lane + 32 < keep) cn[lane + 32] = off1;
                    if (lane == 0) {
                        cn[keep] = total;
                        S.nItems = nIt;
                        S.kAct = S.kActNew;
                        if (keep < nCells) dz_defer(sp, st, S.cellId, keep, nCells);
                        if (keep > 0) {
                            S.peakNeed = fmaxf(S.peakNeed, (float)(total + 8) / (float)keep);
                            S.peakCols = fmaxf(S.peakCols, (float)total / (float)keep);
                        }
                    }
                } else if (warp == 1) {
                    if (lane < keep) { S.wUsed[flip ^ 1][lane] = (unsigned short)wu[0]; S.uPend[flip ^ 1][lane] = (unsigned short)uu[0]; }
                    if (lane + 32 < keep) { S.wUsed[flip ^ 1][lane + 32] = (unsigned short)wu[1]; S.uPend[flip ^ 1][lane + 32] = (unsigned short)uu[1]; }
                }
                // per cell (dealt to the warps, lanes over columns / rows): the position table and the list of the
                // generators just appended (cell, row, column in the layout after the next product)
                DZ_T(7);
                const bool listItems = nIt <= DZ_ITEMS_MAX;
                for (int ci = warp; ci < keep; ci += nWarps) {
                    const int src = ci & 31;
                    const int nb = ci < 32 ? __shfl_sync(0xffffffffu, off0, src) : __shfl_sync(0xffffffffu, off1, src);
                    const int wuse = ci < 32 ? __shfl_sync(0xffffffffu, wu[0], src) : __shfl_sync(0xffffffffu, wu[1], src);
                    const int base = ci < 32 ? __shfl_sync(0xffffffffu, ub0, src) : __shfl_sync(0xffffffffu, ub1, src);
                    const int cb = colOff[ci], wa = colOff[ci + 1] - cb;
                    for (int j = lane; j < wa; j += 32)
                        colPosT[cb + j] = j < wuse ? (unsigned short)((nb + j) | (j == 0 ? 0x8000 : 0)) : (unsigned short)0xFFFF;
                    if (listItems) {
                        int before = 0;
                        for (int ww = 0; ww < mw; ww++) {
                            const unsigned bits = fresh[ci * mS + ww];
                            if ((bits >> lane) & 1u) {
                                const int k = before + __popc(bits & ((1u << lane) - 1u));
                                S.itemCol[base + k] = (unsigned short)(nb + wuse + k);
                                S.itemCell[base + k] = (unsigned short)ci;
                                S.itemRow[base + k] = (unsigned short)(ww * 32 + lane);
                            }
                            before += __popc(bits);
                        }
                    }
                }
                nCells = keep;
            }
            flip ^= 1;
            if (nCells == 0) {
                tileDead = true;
                dz_load_a<KSREG>(S.layers[0], warp, lane, nWarps, a);   // the next tile starts at layer 0 again
                __syncthreads();
                break;
            }
            nCols = colOff[nCells];          // columns of the current layout that the next product reads
            __syncthreads();   // [3] next layout, position table and widths published
            DZ_T(4);
            if (needDead) {
                if (Ly.remove_zero) dz_mark_dead(S.deadMask[flipD], colOff, S.wUsed[flip], cur, P, m, nCells, tid, NT);
                else if (relax) { for (int i = tid; i < CLR_NCOL_MAX / 32; i += NT) S.deadMask[flipD][i] = 0; __syncthreads(); }
            }
        }
        if (tileDead) continue;

//@phase: output
        // ---- outputs ---------------------------------------------------------------------------------------------------
        {
            const int* colOff = S.colOff[flip];
            const unsigned* dead = S.deadMask[flipD];
            const DevLayer& Lt = S.layers[L - 1];
            const bool tailCollapse = fuse && Lt.collapse;
            if (fuse) {
                const int nt = Lt.n, mt = Lt.m;
                if (STATS && Lt.net_layer >= 0 && tid < nCells) {
                    // generators entering the fused layer, as the reference counts them
                    const int cb = colOff[tid], w = S.wUsed[flip][tid];
                    int nz = 0;
                    for (int j = 1; j < w; j++) nz += !((dead[(cb + j) >> 5] >> ((cb + j) & 31)) & 1u);
                    for (int ww = 0; ww < mWords; ww++) nz += __popc(maskBase[flip * maskHalf + tid * mS + ww]);
                    S.cellPin[tid][Lt.net_layer] = (unsigned short)nz;
                    S.cellUnst[tid][Lt.net_layer] = 0;
                }
                // W x for every used column, in place: rows 0 .. mt-1 of the column become its image (a thread per column, k
                // ascending; rows >= kAct of the store are zero).  The fused layer has at most 4 rows: the scalar control of most
                // models, or the folded post-processing.
                if (nt <= 4) {
                    for (int col = tid; col < nCols; col += NT) {
                        int lo_ = 0, hi_ = nCells;                 // cell of this column: the last ci with colOff[ci] <= col
                        while (hi_ - lo_ > 1) { const int mid = (lo_ + hi_) >> 1; if (colOff[mid] <= col) lo_ = mid; else hi_ = mid; }
                        const int j = col - colOff[lo_];
                        if (j >= S.wUsed[flip][lo_]) continue;
                        double* x = cur + col * P;
                        double xv[4], yv[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) xv[k] = k < nt ? x[k] : 0.0;
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            double sacc = 0.0;
#pragma unroll
                            for (int k = 0; k < 4; k++) if (k < nt && r < mt) sacc += __ldg(Lt.Wcm + (size_t)k * mt + r) * xv[k];
                            yv[r] = sacc;
                        }
#pragma unroll
                        for (int r = 0; r < 4; r++) if (r < mt) x[r] = j == 0 ? yv[r] + __ldg(Lt.bias + r) : yv[r];
                    }
                } else {
                    // A real layer (e.g. 100 -> 1, 64 -> 3): on the tensor pipe, like every other product, but without the
                    // product / relaxation / layout cycle -- a warp owns whole column blocks (8 columns), multiplies them by the
                    // layer's single row block (fragments from L2, B fragments conflict-free from the store: a thread per column
                    // walking its column hits an 8-way bank conflict at P = 100 or 68) and overwrites rows 0 .. mt-1 of ITS columns:
                    // no barrier.  Two blocks per warp and round for two independent DMMA chains.  The bias of the centre
                    // columns is added behind the barrier below.
                    const int ksE = min(Lt.ksteps, (S.kAct + 3) >> 2);
                    const double* Af = Lt.Afrag + lane;
                    const int qr = lane >> 2, qc = lane & 3;
                    const int nblk = (nCols + 7) >> 3;
                    for (int b0 = warp * 2; b0 < nblk; b0 += nWarps * 2) {
                        const bool two = b0 + 1 < nblk;
                        const double* x0 = cur + (b0 * 8 + qr) * P + qc;
                        const double* x1 = two ? x0 + 8 * P : x0;
                        double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;
                        for (int k = 0; k < ksE; k++) {
                            const double av = __ldg(Af + k * 32);
                            dmma884(d00, d01, av, x0[k * 4]);
                            dmma884(d10, d11, av, x1[k * 4]);
                        }
                        if (qr < mt) {
                            const int c0 = b0 * 8 + qc * 2;
                            if (c0 < nCols) cur[c0 * P + qr] = d00;
                            if (c0 + 1 < nCols) cur[(c0 + 1) * P + qr] = d01;
                            if (two && c0 + 8 < nCols) cur[(c0 + 8) * P + qr] = d10;
                            if (two && c0 + 9 < nCols) cur[(c0 + 9) * P + qr] = d11;
                        }
                    }
                }
                // images mu_i W[:, i] of the generators appended by the last relaxation, item-major in the 8 slack columns
                // behind the layout (when the tile's list fits; otherwise the output loops walk the row masks)
                if (S.nItems <= DZ_ITEMS_MAX && S.nItems * mt <= 8 * P) {
                    double* img = cur + (size_t)nCols * P;
                    for (int t = tid; t < S.nItems * mt; t += NT) {
                        const int item = t / mt, r = t - item * mt, i = S.itemRow[item];
                        img[t] = stage[S.itemCell[item] * stageW + i] * __ldg(Lt.Wcm + (size_t)i * mt + r);
                    }
                }
                __syncthreads();
                if (nt > 4) {
                    for (int t = tid; t < nCells * mt; t += NT) {
                        const int ci = t / mt, r = t - ci * mt;
                        cur[colOff[ci] * P + r] += __ldg(Lt.bias + r);
                    }
                    __syncthreads();
                }
            }
            // ranks of the surviving output generators (remove_zero_columns); the position table is free now
            unsigned short* colRank = colPosT;
            int* rank0 = S.colOff[flip ^ 1];              // rank of a cell's first pending generator
            if (needDead) {
                if (tid < nCells) {
                    const int cb = colOff[tid], w = S.wUsed[flip][tid];
                    int k = 0;
                    colRank[cb] = 0xFFFF;
                    for (int j = 1; j < w; j++) {
                        const bool dd = (dead[(cb + j) >> 5] >> ((cb + j) & 31)) & 1u;
                        colRank[cb + j] = dd ? (unsigned short)0xFFFF : (unsigned short)(k++);
                    }
                    int u = 0;
                    for (int ww = 0; ww < mWords; ww++) u += __popc(maskBase[flip * maskHalf + tid * mS + ww]);
                    rank0[tid] = k;
                    if (STATS) atomicMax(&S.sMaxP, tailCollapse ? 1 : k + u);      // a one-row map returns a single generator
                    if (STATS && out.keep_cap > 0) {
                        out.keepN[S.cellId[tid]] = k + u;
                        if (k + u > out.keep_cap) atomicCAS(&st->error, 0, (int)CLR_ERR_CAPACITY);
                    }
                }
                __syncthreads();
            }
            for (int it = tid; it < nCells * mOut; it += NT) {
                const int ci = it / mOut, r = it - ci * mOut;
                const int cb = colOff[ci], w = S.wUsed[flip][ci];
                const double* y = cur + cb * P + r;
                const unsigned mwv = (r >> 5) < mWords ? maskBase[flip * maskHalf + ci * mS + (r >> 5)] : 0u;
                const bool pend = !fuse && ((mwv >> (r & 31)) & 1u);     // this row's own appended generator mu * e_r
                const double mu = pend ? stage[ci * stageW + r] : 0.0;
                double c = y[0], lo, hi;
                if (fuse && (w - 1) + (int)S.uPend[flip][ci] > 64) continue;     // wide cell: summed by a whole warp below
                if (fuse) {
                    // the columns hold W x already; the generators appended by the last relaxation are mu_i e_i, their
                    // images mu_i W[r, i] come after them, in ascending i (the order of the materialised columns)
                    const int mwN = (Lt.n + 31) >> 5;
                    if (mOut == 1 && !tailCollapse) {
                        lo = c; hi = c;
                        for (int j = 1; j < w; j++) { double av = fabs(y[j * P]); lo -= av; hi += av; }
                        for (int ww = 0; ww < mwN; ww++) {
                            unsigned bits = maskBase[flip * maskHalf + ci * mS + ww];
                            while (bits) {
                                const int i = ww * 32 + __ffs(bits) - 1;
                                bits &= bits - 1;
                                const double av = fabs(stage[ci * stageW + i] * __ldg(Lt.Wcm + (size_t)i * mOut + r));
                                lo -= av; hi += av;
                            }
                        }
                    } else {
                        double rad = 0.0;
                        for (int j = 1; j < w; j++) rad += fabs(y[j * P]);
                        for (int ww = 0; ww < mwN; ww++) {
                            unsigned bits = maskBase[flip * maskHalf + ci * mS + ww];
                            while (bits) {
                                const int i = ww * 32 + __ffs(bits) - 1;
                                bits &= bits - 1;
                                rad += fabs(stage[ci * stageW + i] * __ldg(Lt.Wcm + (size_t)i * mOut + r));
                            }
                        }
                        lo = c - rad; hi = c + rad;
                    }
                } else if (mOut == 1) {
                    lo = c; hi = c;
                    for (int j = 1; j < w; j++) { double av = fabs(y[j * P]); lo -= av; hi += av; }
                    if (pend) { lo -= fabs(mu); hi += fabs(mu); }
                } else {
                    double rad = 0.0;
                    for (int j = 1; j < w; j++) rad += fabs(y[j * P]);
                    if (pend) rad += fabs(mu);
                    lo = c - rad; hi = c + rad;
                }
                const long long cell = S.cellId[ci];
                if (out.out_lo) { out.out_lo[cell * mOut + r] = lo; out.out_hi[cell * mOut + r] = hi; }
                if (out.hull_lo) {
                    if (useHull) { atomicMinD(&S.hlo[r], lo); atomicMaxD(&S.hhi[r], hi); }
                    else { atomicMinD(&out.hull_lo[r], lo); atomicMaxD(&out.hull_hi[r], hi); }
                }
                if (STATS && out.keep_cap > 0) {
                    out.keepC[cell * mOut + r] = c;
                    double* g = out.keepG + (size_t)cell * out.keep_cap * mOut + r;
                    for (int j = 1; j < w; j++) {
                        const unsigned short k = colRank[cb + j];
                        if (k != 0xFFFF && k < out.keep_cap) g[(size_t)k * mOut] = y[j * P];
                    }
                    int k0 = rank0[ci], u = 0;
                    for (int ww = 0; ww < mWords; ww++) u += __popc(maskBase[flip * maskHalf + ci * mS + ww]);
                    int rank = __popc(mwv & ((1u << (r & 31)) - 1u));
                    for (int ww = 0; ww < (r >> 5); ww++) rank += __popc(maskBase[flip * maskHalf + ci * mS + ww]);
                    for (int k = 0; k < u; k++)
                        if (k0 + k < out.keep_cap) g[(size_t)(k0 + k) * mOut] = (pend && k == rank) ? mu : 0.0;
                }
            }
            if (fuse) {
                // Cells of more than 64 generators: a warp per (cell, output row).  The radius is DEFINED as the sum over the lanes
                // (fixed butterfly) of each lane's ascending sum over the generators g = lane, lane + 32, ... of the list
                // [materialised columns | appended generators in ascending row order] -- the same for every tile and pass.
                const int mt = Lt.m, mwN = (Lt.n + 31) >> 5;
                const bool listed = S.nItems <= DZ_ITEMS_MAX && S.nItems * mt <= 8 * P;
                const double* img = cur + (size_t)nCols * P;
                const unsigned short* ibase = S.uPend[flip ^ 1];
                for (int task = warp; task < nCells * mOut; task += nWarps) {
                    const int ci = task / mOut, r = task - ci * mOut;
                    const int w = S.wUsed[flip][ci], p = (w - 1) + (int)S.uPend[flip][ci];
                    if (p <= 64) continue;
                    const double* y = cur + colOff[ci] * P + r;
                    double acc = 0.0;
                    if (listed) {
                        const double* im = img + (size_t)ibase[ci] * mt + r;
                        for (int g = lane; g < p; g += 32) acc += fabs(g < w - 1 ? y[(g + 1) * P] : im[(g - (w - 1)) * mt]);
                    } else {
                        for (int g = lane; g < w - 1; g += 32) acc += fabs(y[(g + 1) * P]);
                        int g = w - 1;
                        for (int ww = 0; ww < mwN; ww++) {
                            unsigned bits = maskBase[flip * maskHalf + ci * mS + ww];
                            while (bits) {
                                const int i = ww * 32 + __ffs(bits) - 1;
                                bits &= bits - 1;
                                if ((g & 31) == lane) acc += fabs(stage[ci * stageW + i] * __ldg(Lt.Wcm + (size_t)i * mt + r));
                                g++;
                            }
                        }
                    }
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                    if (lane == 0) {
                        const double c = y[0], lo = c - acc, hi = c + acc;
                        const long long cell = S.cellId[ci];
                        if (out.out_lo) { out.out_lo[cell * mOut + r] = lo; out.out_hi[cell * mOut + r] = hi; }
                        if (out.hull_lo) {
                            if (useHull) { atomicMinD(&S.hlo[r], lo); atomicMaxD(&S.hhi[r], hi); }
                            else { atomicMinD(&out.hull_lo[r], lo); atomicMaxD(&out.hull_hi[r], hi); }
                        }
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
        DZ_T(5);
//@phase: tile size / tail
        // ---- adapt the tile size to the growth just seen ---------------------------------------------------------------
        if (tid == 0) {
            if (sp.tile_cells <= 0 && sp.pass != 2) {
                const float margin = sp.pass == 0 ? DZ_MARGIN : 1.5f;
                // columns and staging both scale with the number of cells
                const float perCell = S.peakNeed * margin * (float)P * (inplace ? 1.f : 2.f) + (float)stageW;
                int T = (int)((float)(sp.cap - 16 * P) / perCell);
                if (sp.grow_cols > 0 && sp.pass == 0) {
                    // sigmoid / tanh layers: the growth of a cell is known, not estimated -- the largest tile whose columns,
                    // 8 columns of slack and staging fit exactly (wide cells: one more cell per tile is a large step)
                    const float per = S.peakCols * (float)P * (inplace ? 1.f : 2.f) + (float)stageW;
                    T = max(T, (int)((float)(sp.cap - 8 * P * (inplace ? 1 : 2)) / per));
                }
                S.tileT = max(1, min(T, tileMax));
            }
            if (STATS) atomicAdd(&st->tiles, 1ull);
        }
        __syncthreads();
    }

#ifdef DZ_TIMING
    if (tid == 0) for (int i = 0; i < 8; i++) atomicAdd(&st->timers[i], (unsigned long long)tacc[i]);
#endif
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
