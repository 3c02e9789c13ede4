// project.cuh -- batched _project_oa of Taylor-model reach-sets.
//
// The step immediately BEFORE the controller propagation at control periods k > 1 (reference src/solve.jl:181-184):
//     X0 = set(_project_oa(R, st_vars, t))            with        (src/setops.jl:9-12)
//     _project_oa(R::AbstractTaylorModelReachSet, vars, t) = project(set(overapproximate(R, Zonotope, t)), vars)
// for all live chains at once, so that the cells entering clr_deepz_zonotopes stay array-backed.
//
// Upstream semantics restated here [UPSTREAM-RECALL] (ReachabilityAnalysis overapproximate(::TaylorModelReachSet,
// Zonotope, t) -> TaylorModels evaluate / fp_rpa -> LazySets overapproximate(::Vector{<:TaylorModelN}, Zonotope)):
//   1. every component v is a TaylorModel1 in time whose coefficients are TaylorN polynomials in the n normalised
//      space variables xi in [-1, 1]^n, plus an interval remainder D_v; it is evaluated at the time offset dt from
//      its expansion point by Horner's rule on every monomial coefficient (plain Float64 arithmetic, as TaylorSeries
//      does), and D_v is added to the constant term:                  I0 = c_0 (+) D_v   (outward rounded);
//   2. fp_rpa: the constant term becomes mid(I0), the remainder  I0 - mid(I0);
//   3. LazySets: the constant and linear terms give the centre and one generator column per space variable; the
//      nonlinear monomials are bounded over [-1, 1]^n (a monomial with only even exponents ranges over [0, 1], any
//      other over [-1, 1]) and summed with the remainder in outward-rounded interval arithmetic; the midpoint of that
//      interval moves the centre, its radius becomes a generator r_v e_v;
//   4. project: the rows `vars` of the centre and of [G_lin | diag(r)]; zero columns dropped when requested.
//   3b. (merge != 0) LazySets' remove_redundant_generators, the default of that overapproximate: on the FULL n-dimensional
//      zonotope zero columns are dropped and every column that is a multiple of an earlier (accumulated) column is added to
//      it -- ReachabilityBase ismultiple with its approximate comparisons.  All n components are then evaluated, not only
//      the kept ones: whether two columns are parallel is decided in n dimensions, before the projection.
//
// One warp per (reach-set, kept variable): lanes over the monomials, coalesced reads of the coefficient block.
// HBM-bound streaming kernel: 8 M (order_t + 1) bytes read per (reach-set, kept variable), a few hundred written.
#pragma once
#include "common.cuh"

struct ProjArgs {
    long long N;
    int n, KT, M, nkeep;        // space variables, time coefficients (order_t + 1), monomials, kept variables
    int remove_zero;
    int merge;                  // 1: remove_redundant_generators before the projection (evaluates all n components: nkeep == n in phase 1)
    int nproj; const int* pvars; // merge: the projection applied by the pack kernel (nproj kept variables)
    const int* cls;             // per monomial: -1 constant, j >= 0 linear in variable j, -2 nonlinear all-even, -3 nonlinear
    const double* coeffs;       // [N][n][KT][M]   (monomial index fastest)
    const double* rem;          // [N][n][2]       lo, hi of the remainder
    const double* dt;           // [N]
    const int* vars;            // [nkeep] 0-based
    double* outC;               // [N][nkeep]
    double* outG;               // [N][2 n][nkeep]   generator j of reach-set i at outG + (i * 2 n + j) * nkeep
    int* outN;                  // [N] generator count
    double* scratch;            // [N][nkeep][n + 2]: linear coefficients, centre, radius of every kept variable
};

#define PJ_WARPS 4

// phase 1: one warp per (reach-set, kept variable): time evaluation, bounds of the nonlinear part
__global__ void __launch_bounds__(PJ_WARPS * 32, 6) project_eval_kernel(ProjArgs A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long units = A.N * A.nkeep;
    for (long long u = (long long)blockIdx.x * PJ_WARPS + warp; u < units; u += (long long)gridDim.x * PJ_WARPS) {
        const long long cell = u / A.nkeep;
        const int kv = (int)(u - cell * A.nkeep), v = A.vars[kv];
        const double* cf = A.coeffs + ((size_t)cell * A.n + v) * A.KT * A.M;
        const double dt = A.dt[cell];
        double* sc = A.scratch + ((size_t)cell * A.nkeep + kv) * (A.n + 2);
        double c0 = 0.0, nl_lo = 0.0, nl_hi = 0.0;
        auto account = [&](int m, double s) {
            const int c = A.cls[m];
            if (c == -1) c0 = s;
            else if (c >= 0) sc[c] = s;                         // linear coefficient of variable c: generator entry
            else if (c == -2) {                                 // s * [0, 1]
                nl_lo = __dadd_rd(nl_lo, fmin(s, 0.0)); nl_hi = __dadd_ru(nl_hi, fmax(s, 0.0));
            } else {                                            // s * [-1, 1]
                nl_lo = __dadd_rd(nl_lo, -fabs(s)); nl_hi = __dadd_ru(nl_hi, fabs(s));
            }
        };
        // Horner in time, highest coefficient first (TaylorSeries evaluate: suma = suma * dt + a[k]).
        if (A.KT <= 12 && A.M <= 96) {
            // the common shapes (order_t <= 11, up to 96 monomials): all loads of the unit in flight before the first
            // chain -- the kernel is bound by the bytes it keeps in flight (0 * dt + a[KT-1] = a[KT-1] exactly)
            double t[3][12];
#pragma unroll
            for (int q = 0; q < 3; q++)
#pragma unroll
                for (int k = 0; k < 12; k++)      // clamped addresses: unconditional loads, which the compiler can batch
                    t[q][k] = __ldg(cf + (size_t)min(k, A.KT - 1) * A.M + min(lane + 32 * q, A.M - 1));
#pragma unroll
            for (int q = 0; q < 3; q++) {
                double s = 0.0;
#pragma unroll
                for (int k = 11; k >= 0; k--) if (k < A.KT) s = __dadd_rn(__dmul_rn(s, dt), t[q][k]);
                if (lane + 32 * q < A.M) account(lane + 32 * q, s);
            }
        } else {
            for (int m = lane; m < A.M; m += 32) {
                double s = cf[(size_t)(A.KT - 1) * A.M + m];
                for (int k = A.KT - 2; k >= 0; k--) s = __dadd_rn(__dmul_rn(s, dt), cf[(size_t)k * A.M + m]);
                account(m, s);
            }
        }
        // interval sum over the lanes (outward rounded; the order differs from a sequential sum by rounding only)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            nl_lo = __dadd_rd(nl_lo, __shfl_xor_sync(0xffffffffu, nl_lo, d));
            nl_hi = __dadd_ru(nl_hi, __shfl_xor_sync(0xffffffffu, nl_hi, d));
            c0 += __shfl_xor_sync(0xffffffffu, c0, d);          // exactly one lane holds the constant term
        }
        if (lane == 0) {
            const double dlo = A.rem[((size_t)cell * A.n + v) * 2], dhi = A.rem[((size_t)cell * A.n + v) * 2 + 1];
            const double i0lo = __dadd_rd(c0, dlo), i0hi = __dadd_ru(c0, dhi);        // constant term + remainder
            const double c0m = 0.5 * (i0lo + i0hi);                                   // fp_rpa: midpoint ...
            const double rlo = __dadd_rd(i0lo, -c0m), rhi = __dadd_ru(i0hi, -c0m);    // ... and what is left of the interval
            const double lo = __dadd_rd(nl_lo, rlo), hi = __dadd_ru(nl_hi, rhi);      // nonlinear range + remainder
            sc[A.n] = c0m + 0.5 * (lo + hi);                                          // centre
            sc[A.n + 1] = 0.5 * __dadd_ru(hi, -lo);                                   // radius of the remainder generator
        }
    }
}

// phase 2: one thread per reach-set assembles project(Z, vars): columns [linear_0 .. linear_{n-1} | r_v e_v ...]
__global__ void project_pack_kernel(ProjArgs A) {
    const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cell >= A.N) return;
    const int n = A.n, nk = A.nkeep, cap = 2 * n;
    const double* sc = A.scratch + (size_t)cell * nk * (n + 2);
    double* G = A.outG + (size_t)cell * cap * nk;
    int p = 0;
    for (int j = 0; j < n; j++) {                      // linear generators
        bool any = false;
        for (int kv = 0; kv < nk; kv++) any = any || sc[(size_t)kv * (n + 2) + j] != 0.0;
        if (any || !A.remove_zero) {
            for (int kv = 0; kv < nk; kv++) G[(size_t)p * nk + kv] = sc[(size_t)kv * (n + 2) + j];
            p++;
        }
    }
    for (int v = 0; v < n; v++) {                      // remainder generators r_v e_v: zero columns unless v is kept
        int kv = -1;
        for (int k2 = 0; k2 < nk; k2++) if (A.vars[k2] == v) kv = k2;
        const double r = kv >= 0 ? sc[(size_t)kv * (n + 2) + n + 1] : 0.0;
        if (r != 0.0 || !A.remove_zero) {
            for (int k2 = 0; k2 < nk; k2++) G[(size_t)p * nk + k2] = k2 == kv ? r : 0.0;
            p++;
        }
    }
    for (int kv = 0; kv < nk; kv++) A.outC[cell * nk + kv] = sc[(size_t)kv * (n + 2) + n];
    A.outN[cell] = p;
    for (int j = p; j < cap; j++)
        for (int kv = 0; kv < nk; kv++) G[(size_t)j * nk + kv] = 0.0;
}

// ---- merge != 0: phase 2 with LazySets' remove_redundant_generators on the full zonotope, then the projection ----------
__device__ __forceinline__ bool pj_zero(double x) { return fabs(x) <= CLR_ZTOL; }
__device__ __forceinline__ bool pj_approx(double x, double y) {
    if (pj_zero(x) && pj_zero(y)) return true;
    return x == y || fabs(x - y) <= CLR_RTOL * fmax(fabs(x), fabs(y));
}
#define PJ_NMAX 64
// one thread per reach-set; scratch holds all n components: sc[v][0..n-1] linear, sc[v][n] centre, sc[v][n+1] radius
__global__ void project_pack_merge_kernel(ProjArgs A) {
    const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (cell >= A.N) return;
    const int n = A.n, nk = A.nproj, cap = 2 * n;
    const double* sc = A.scratch + (size_t)cell * n * (n + 2);
    auto entry = [&](int col, int v) -> double {          // column `col` of [G_lin | diag(r)], row v
        return col < n ? sc[(size_t)v * (n + 2) + col] : (col - n == v ? sc[(size_t)v * (n + 2) + n + 1] : 0.0);
    };
    double* G = A.outG + (size_t)cell * cap * nk;
    unsigned long long done[2] = {0ull, 0ull};            // dropped columns (zero, or merged into an earlier one)
    int p_all = 0;
    for (int j = 0; j < cap; j++) {                       // remove_zero_columns
        bool any = false;
        for (int v = 0; v < n; v++) any = any || entry(j, v) != 0.0;
        if (!any) done[j >> 6] |= 1ull << (j & 63); else p_all++;
    }
    double g1[PJ_NMAX];
    int p = 0;
    for (int j1 = 0; j1 < cap; j1++) {
        if ((done[j1 >> 6] >> (j1 & 63)) & 1ull) continue;
        for (int v = 0; v < n; v++) g1[v] = entry(j1, v);
        if (p_all > 1) {
            for (int j2 = j1 + 1; j2 < cap; j2++) {
                if ((done[j2 >> 6] >> (j2 & 63)) & 1ull) continue;
                // ReachabilityBase.Arrays.ismultiple(g1, column j2)
                bool ok = true, nofac = true; double fac = 0.0;
                for (int v = 0; v < n && ok; v++) {
                    const double u = g1[v], w = entry(j2, v);
                    if (pj_zero(u)) { if (!pj_zero(w)) ok = false; continue; }
                    if (pj_zero(w)) { ok = false; continue; }
                    if (nofac) { nofac = false; fac = u / w; }
                    else if (!pj_approx(fac, u / w)) ok = false;
                }
                if (ok) {
                    if (!nofac && fac > 0.0) { for (int v = 0; v < n; v++) g1[v] += entry(j2, v); }
                    else { for (int v = 0; v < n; v++) g1[v] -= entry(j2, v); }
                    done[j2 >> 6] |= 1ull << (j2 & 63);
                }
            }
        }
        // project(Z, vars) + remove_zero_generators of the projection
        bool any = false;
        for (int k = 0; k < nk; k++) any = any || g1[A.pvars[k]] != 0.0;
        if (any || !A.remove_zero) {
            for (int k = 0; k < nk; k++) G[(size_t)p * nk + k] = g1[A.pvars[k]];
            p++;
        }
    }
    for (int k = 0; k < nk; k++) A.outC[cell * nk + k] = sc[(size_t)A.pvars[k] * (n + 2) + n];
    A.outN[cell] = p;
    for (int j = p; j < cap; j++)
        for (int k = 0; k < nk; k++) G[(size_t)j * nk + k] = 0.0;
}
