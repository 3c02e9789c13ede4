// rollout.cuh -- batched closed-loop rollouts of simulate() and the pointwise controller.
//
// Replaces the period loop of the reference's simulate (src/simulate.jl:98-140): per trajectory and
// control period the controller forward(x, network) (src/simulate.jl:101-106), the extended state
// [x; w; u] (src/simulate.jl:117-124) and the ODE solve of _solve_ensemble
// (ext/ClosedLoopReachabilityOrdinaryDiffEqExt.jl:28-45: Tsit5() with OrdinaryDiffEq's defaults).
// Three forms of the same arithmetic per trajectory:
//   * narrow controllers (every materialised vector has at most 8 entries, e.g. Unicycle's 4 -> 500 -> 2) whose packed image fits
//     the constant bank: per control period ONE controller launch (controller_cw_kernel: weights fetched into uniform registers,
//     4 points per thread) and ONE plant launch (rollout_ode_kernel: one trajectory per thread); the state travels through the
//     X_end / U arrays the call returns anyway -- see the end of this file;
//   * other narrow controllers: rollout_blocked_kernel, a thread owns 4 trajectories for the whole horizon, weights in shared memory;
//   * general controllers: rollout_kernel, one trajectory per thread for the whole horizon.
// The only HBM traffic is the state / disturbance read and the X_end / U write per period.
#pragma once
#ifndef __CUDACC_RTC__
#include "common.cuh"
#endif
// (under NVRTC -- user-supplied plants, clr_plant_compile -- the host prepends the few constants this file needs)

#define RO_MAX_LAYERS (CLR_DEV_MAX_LAYERS)

// Pointwise network: row-major copy of each W (dot products with the input) and the column-major one
// (streaming a hidden neuron into the next layer's accumulators).
struct PointNet {
    int L;
    int n_in, m_out;
    int total_doubles;            // doubles needed to stage all weights in smem
    int dims[RO_MAX_LAYERS + 1];
    int act[RO_MAX_LAYERS];
    const double* Wrm[RO_MAX_LAYERS];
    const double* Wcm[RO_MAX_LAYERS];
    const double* b[RO_MAX_LAYERS];
    int off_rm[RO_MAX_LAYERS], off_cm[RO_MAX_LAYERS], off_b[RO_MAX_LAYERS];   // offsets in the smem image
    const double* image;          // device: all weights in smem-image order
    // packed image of narrow networks (every materialised vector has at most 8 entries), see ro_mlp_small
    int small_ni, small_no, stotal;
    int soff[RO_MAX_LAYERS];
    const double* simage;
    long long sserial;            // host side: identifies the packed image (which one is resident in the constant bank)
};

__device__ __forceinline__ double ro_act(int act, double s) {
    switch (act) {
        case CLR_ACT_RELU: return s > 0.0 ? s : 0.0;
        case CLR_ACT_SIGMOID: return 1.0 / (1.0 + exp(-s));
        case CLR_ACT_TANH: return tanh(s);
        default: return s;
    }
}

// max(s, 0) without FP64 instructions: clears every bit when the sign bit is set (-0.0 -> +0.0, like s > 0 ? s : 0)
__device__ __forceinline__ double ro_relu_bits(double s) {
    const int hi = __double2hiint(s), lo = __double2loint(s);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, lo & keep);
}

// Evaluate the network on one point.  Layers are processed in pairs: neuron i of layer l is computed
// from the stored input vector and immediately streamed into the accumulators of layer l+1, so
// only every other activation vector is materialised.  Summation order = ControllerFormats'
// W*x + b (k ascending, bias last); products and sums are fused (fma).
template <int MAXW>
__device__ __forceinline__ void ro_mlp(const PointNet& net, const double* __restrict__ w, const double* xin,
                                       double* uout) {
    double va[MAXW], vb[MAXW];
    if (MAXW <= 8) {
#pragma unroll
        for (int k = 0; k < MAXW; k++) va[k] = k < net.n_in ? xin[k] : 0.0;
    } else {
        for (int k = 0; k < net.n_in; k++) va[k] = xin[k];
    }
    int l = 0;
    while (l < net.L) {
        const int n = net.dims[l], m = net.dims[l + 1];
        if (l + 1 < net.L) {
            const int m2 = net.dims[l + 2];
            const double* W1 = w + net.off_rm[l];
            const double* b1 = w + net.off_b[l];
            const double* W2 = w + net.off_cm[l + 1];
            const double* b2 = w + net.off_b[l + 1];
            const int a1 = net.act[l], a2 = net.act[l + 1];
            if (MAXW <= 8) {
#pragma unroll
                for (int o = 0; o < MAXW; o++) vb[o] = 0.0;
            } else {
                for (int o = 0; o < m2; o++) vb[o] = 0.0;
            }
#pragma unroll 4
            for (int i = 0; i < m; i++) {
                double s = 0.0;
                if (MAXW <= 8) {
#pragma unroll
                    for (int k = 0; k < MAXW; k++) if (k < n) s = fma(W1[i * n + k], va[k], s);
                } else {
                    for (int k = 0; k < n; k++) s = fma(W1[i * n + k], va[k], s);
                }
                s += b1[i];
                const double h = ro_act(a1, s);
                if (MAXW <= 8) {
#pragma unroll
                    for (int o = 0; o < MAXW; o++) if (o < m2) vb[o] = fma(W2[i * m2 + o], h, vb[o]);
                } else {
                    for (int o = 0; o < m2; o++) vb[o] = fma(W2[i * m2 + o], h, vb[o]);
                }
            }
            if (MAXW <= 8) {
#pragma unroll
                for (int o = 0; o < MAXW; o++) va[o] = o < m2 ? ro_act(a2, vb[o] + b2[o]) : 0.0;
            } else {
                for (int o = 0; o < m2; o++) va[o] = ro_act(a2, vb[o] + b2[o]);
            }
            l += 2;
        } else {
            const double* W1 = w + net.off_rm[l];
            const double* b1 = w + net.off_b[l];
            const int a1 = net.act[l];
            if (MAXW <= 8) {
#pragma unroll
                for (int i = 0; i < MAXW; i++) {
                    if (i < m) {
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < MAXW; k++) if (k < n) s = fma(W1[i * n + k], va[k], s);
                        vb[i] = ro_act(a1, s + b1[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < MAXW; i++) va[i] = i < m ? vb[i] : 0.0;
            } else {
                for (int i = 0; i < m; i++) {
                    double s = 0.0;
                    for (int k = 0; k < n; k++) s = fma(W1[i * n + k], va[k], s);
                    vb[i] = ro_act(a1, s + b1[i]);
                }
                for (int i = 0; i < m; i++) va[i] = vb[i];
            }
            l += 1;
        }
    }
    if (MAXW <= 8) {
#pragma unroll
        for (int k = 0; k < MAXW; k++) if (k < net.m_out) uout[k] = va[k];
    } else {
        for (int k = 0; k < net.m_out; k++) uout[k] = va[k];
    }
}

// Narrow networks (e.g. Unicycle 4 -> 500 -> 2): every materialised vector fits NI / NO registers.  The host packs, per
// hidden neuron i of a layer pair, one zero-padded record [W1[i, 0..NI) | W2[0..NO, i] | b1[i] | 0], so the inner loop is
// NI/2 + NO/2 + 1 broadcast shared loads (128-bit) and NI + NO + 2 FP64 instructions, no predicates.
template <int NI, int NO>
__device__ __forceinline__ void ro_mlp_small(const PointNet& net, const double* __restrict__ w, const double* xin,
                                             double* uout) {
    constexpr int NV = NI > NO ? NI : NO;
    double va[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) va[k] = k < net.n_in ? xin[k] : 0.0;
    int l = 0;
    while (l < net.L) {
        const int m = net.dims[l + 1];
        if (l + 1 < net.L) {
            constexpr int ST = NI + NO + 2;
            const double* rec = w + net.soff[l];
            const double* b2 = w + net.soff[l + 1];
            const int a1 = net.act[l], a2 = net.act[l + 1];
            double vb[NO];
#pragma unroll
            for (int o = 0; o < NO; o++) vb[o] = 0.0;
            if (a1 == CLR_ACT_RELU) {
#pragma unroll 4
                for (int i = 0; i < m; i++) {
                    const double2* r2 = reinterpret_cast<const double2*>(rec + i * ST);
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NI / 2; k++) { const double2 wv = r2[k]; s = fma(wv.x, va[2 * k], s); s = fma(wv.y, va[2 * k + 1], s); }
                    s += rec[i * ST + NI + NO];
                    const double h = fmax(s, 0.0);
#pragma unroll
                    for (int o = 0; o < NO / 2; o++) { const double2 vv = r2[NI / 2 + o]; vb[2 * o] = fma(vv.x, h, vb[2 * o]); vb[2 * o + 1] = fma(vv.y, h, vb[2 * o + 1]); }
                }
            } else {
                for (int i = 0; i < m; i++) {
                    const double2* r2 = reinterpret_cast<const double2*>(rec + i * ST);
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NI / 2; k++) { const double2 wv = r2[k]; s = fma(wv.x, va[2 * k], s); s = fma(wv.y, va[2 * k + 1], s); }
                    s += rec[i * ST + NI + NO];
                    const double h = ro_act(a1, s);
#pragma unroll
                    for (int o = 0; o < NO / 2; o++) { const double2 vv = r2[NI / 2 + o]; vb[2 * o] = fma(vv.x, h, vb[2 * o]); vb[2 * o + 1] = fma(vv.y, h, vb[2 * o + 1]); }
                }
            }
            const int m2 = net.dims[l + 2];
#pragma unroll
            for (int o = 0; o < NV; o++) va[o] = (o < NO && o < m2) ? ro_act(a2, vb[o < NO ? o : 0] + b2[o < NO ? o : 0]) : 0.0;
            l += 2;
        } else {
            constexpr int ST1 = NI + 2;
            const double* rec = w + net.soff[l];
            const int a1 = net.act[l];
            double vb[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) {
                vb[i] = 0.0;
                if (i < m) {
                    const double2* r2 = reinterpret_cast<const double2*>(rec + i * ST1);
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NI / 2; k++) { const double2 wv = r2[k]; s = fma(wv.x, va[2 * k], s); s = fma(wv.y, va[2 * k + 1], s); }
                    vb[i] = ro_act(a1, s + rec[i * ST1 + NI]);
                }
            }
#pragma unroll
            for (int i = 0; i < NV; i++) va[i] = vb[i];
            l += 1;
        }
    }
#pragma unroll
    for (int k = 0; k < NV; k++) if (k < net.m_out) uout[k] = va[k];
}

// CODE selects the evaluator: NI * 16 + NO for narrow networks (66: 4/2, 72: 4/8, 136: 8/8), 1 / 2: generic with 128 / 512
// materialised entries.  xin / uout hold at least 8 entries for the narrow codes.
template <int CODE>
__device__ __forceinline__ void ro_eval(const PointNet& net, const double* __restrict__ w, const double* xin, double* uout) {
    if (CODE == 1) ro_mlp<128>(net, w, xin, uout);
    else if (CODE == 2) ro_mlp<512>(net, w, xin, uout);
    else ro_mlp_small<(CODE >> 4), (CODE & 15)>(net, w, xin, uout);
}

// widths that must be materialised by ro_mlp (host uses the same rule to pick MAXW)
__host__ __device__ inline int ro_stored_width(const int* dims, int L) {
    int w = dims[0], l = 0;
    while (l < L) {
        if (l + 1 < L) { if (dims[l + 2] > w) w = dims[l + 2]; l += 2; }
        else { if (dims[l + 1] > w) w = dims[l + 1]; l += 1; }
    }
    return w;
}

__device__ __forceinline__ const double* ro_stage_weights(const PointNet& net, double* sm, bool in_smem, bool narrow) {
    const double* img = narrow ? net.simage : net.image;
    const int cnt = narrow ? net.stotal : net.total_doubles;
    if (!in_smem) return img;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) sm[i] = __ldg(img + i);
    __syncthreads();
    return sm;
}

template <int CODE>
__global__ void __launch_bounds__(128)
forward_points_kernel(PointNet net, int nx, long long N, const double* __restrict__ X, double* __restrict__ U,
                      int in_smem) {
    extern __shared__ __align__(16) double ro_sm[];
    const double* w = ro_stage_weights(net, ro_sm, in_smem != 0, CODE > 2);
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < N; j += (long long)gridDim.x * blockDim.x) {
        if (CODE > 2) {
            double x[8], u[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = k < nx ? X[j * nx + k] : 0.0;
            ro_eval<CODE>(net, w, x, u);
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < net.m_out) U[j * net.m_out + k] = u[k];
        } else {
            ro_eval<CODE>(net, w, X + j * nx, U + j * net.m_out);
        }
    }
}

// ---- plants on the extended state [x; w; u]; only the n_state leading entries have dynamics ---------
template <int PLANT> struct PlantDims;
template <> struct PlantDims<CLR_PLANT_UNICYCLE> { static constexpr int NS = 4, ND = 1, NC = 2; };
template <> struct PlantDims<CLR_PLANT_TORA> { static constexpr int NS = 4, ND = 0, NC = 1; };
template <> struct PlantDims<CLR_PLANT_ACC> { static constexpr int NS = 6, ND = 0, NC = 1; };
template <> struct PlantDims<CLR_PLANT_INVPEND> { static constexpr int NS = 2, ND = 0, NC = 1; };

#ifdef CLR_USER_NS
// run-time compiled plant (clr_plant_compile): dimensions and right-hand side body come from the caller
#define CLR_PLANT_USER 1000
template <> struct PlantDims<CLR_PLANT_USER> { static constexpr int NS = CLR_USER_NS, ND = CLR_USER_ND, NC = CLR_USER_NC; };
__device__ void clr_user_rhs(const double* x, const double* q, double* dx);   // defined by the host after this source
#endif

// x: NS dynamic states, q: the constant tail [w; u]
template <int PLANT>
__device__ __forceinline__ void plant_rhs(const double* x, const double* q, double* dx) {
#ifdef CLR_USER_NS
    if (PLANT == CLR_PLANT_USER) { clr_user_rhs(x, q, dx); return; }
#endif
    if (PLANT == CLR_PLANT_UNICYCLE) {            // models/Unicycle/Unicycle.jl:36-47
        double sn, cs;
        sincos(x[2], &sn, &cs);
        dx[0] = x[3] * cs;
        dx[1] = x[3] * sn;
        dx[2] = q[2];
        dx[3] = q[1] + q[0];
    } else if (PLANT == CLR_PLANT_TORA) {         // models/TORA/TORA.jl:46-55
        dx[0] = x[1];
        dx[1] = -x[0] + (0.1 * sin(x[2]));
        dx[2] = x[3];
        dx[3] = q[0];
    } else if (PLANT == CLR_PLANT_ACC) {          // models/ACC/ACC.jl:41-60
        dx[0] = x[1];
        dx[1] = x[2];
        dx[2] = 2 * (-2.0 - x[2]) - 0.0001 * (x[1] * x[1]);
        dx[3] = x[4];
        dx[4] = x[5];
        dx[5] = 2 * (q[0] - x[5]) - 0.0001 * (x[4] * x[4]);
    } else {                                      // models/InvertedPendulum/InvertedPendulum.jl:44-51
        dx[0] = x[1];
        dx[1] = 2.0 * sin(x[0]) + 8.0 * (q[0] - 0.0 * x[1]);
    }
}

// DiffEqBase fastpow (Float32 round trip), bit-identical to the CPU restatement
__device__ __forceinline__ float ro_fastlog2f(float x) {
    const float a = 0.338953f, b = 2.198599f, c = 1.523692f;
    unsigned ux1i = __float_as_uint(x);
    unsigned ex = (ux1i & 0x7F800000u) >> 23;
    unsigned greater = ux1i & 0x00400000u;
    float signif, fexp;
    if (greater) { signif = __uint_as_float((ux1i & 0x007FFFFFu) | 0x3f000000u); fexp = __fsub_rn((float)ex, 126.0f); }
    else { signif = __uint_as_float((ux1i & 0x007FFFFFu) | 0x3f800000u); fexp = __fsub_rn((float)ex, 127.0f); }
    signif = __fsub_rn(signif, 1.0f);
    float num = __fmul_rn(signif, __fadd_rn(__fmul_rn(a, signif), b));
    return __fadd_rn(fexp, __fdiv_rn(num, __fadd_rn(signif, c)));
}
__device__ __forceinline__ float ro_fastpow2f(float x) {
    float offset = x < 0 ? 1.0f : 0.0f;
    float clipp = x < -126.0f ? -126.0f : x;
    int w = (int)clipp;
    float z = __fadd_rn(__fsub_rn(clipp, (float)w), offset);
    float t = __fadd_rn(clipp, 121.2740575f);
    t = __fadd_rn(t, __fdiv_rn(27.7280233f, __fsub_rn(4.84252568f, z)));
    t = __fsub_rn(t, __fmul_rn(1.49012907f, z));
    float v = __fmul_rn((float)(1 << 23), t);
    return __uint_as_float((unsigned)v);
}
__device__ __forceinline__ double ro_pow(double x, double y, int mode) {
    if (mode == 0) return pow(x, y);
    if (x == 0.0) return 0.0;
    return (double)ro_fastpow2f(__fmul_rn((float)y, ro_fastlog2f((float)x)));
}

struct RolloutArgs {
    long long N;
    int iters;
    const double* x0;      // NS x N
    const double* Wd;      // ND x N x iters or nullptr
    double t0, tau, T;
    int alg;
    double abstol, reltol;
    int substeps, pow_mode;
    double* X_end;         // NS x N x iters
    double* U;             // NC x N x iters
    int* naccept;          // N x iters or nullptr
    int* nreject;
    int log_cap;
    int* log_n;
    double* log_t;
    double* log_u;
    int* error;            // device flag: CLR_ERR_CAPACITY when a step log overflows, CLR_ERR_UNSTABLE when an integration gave up
};

// Tsit5 tableau in the constant bank: a coefficient is then a direct c[bank][offset] operand of the FP64 instruction instead
// of a 64-bit literal that takes two extra instructions (UMOV / IMAD.MOV) to materialise at every use
__constant__ double ro_ts[28] = {
    0.161,
    -0.008480655492356989,
    0.335480655492357,
    2.8971530571054935,
    -6.359448489975075,
    4.3622954328695815,
    5.325864828439257,
    -11.748883564062828,
    7.4955393428898365,
    -0.09249506636175525,
    5.86145544294642,
    -12.92096931784711,
    8.159367898576159,
    -0.071584973281401,
    -0.028269050394068383,
    0.09646076681806523,
    0.01,
    0.4798896504144996,
    1.379008574103742,
    -3.290069515436081,
    2.324710524099774,
    -0.00178001105222577714,
    -0.0008164344596567469,
    0.007880878010261995,
    -0.1447110071732629,
    0.5823571654525552,
    -0.45808210592918697,
    0.015151515151515152};
#define TS_A21 ro_ts[0]
#define TS_A31 ro_ts[1]
#define TS_A32 ro_ts[2]
#define TS_A41 ro_ts[3]
#define TS_A42 ro_ts[4]
#define TS_A43 ro_ts[5]
#define TS_A51 ro_ts[6]
#define TS_A52 ro_ts[7]
#define TS_A53 ro_ts[8]
#define TS_A54 ro_ts[9]
#define TS_A61 ro_ts[10]
#define TS_A62 ro_ts[11]
#define TS_A63 ro_ts[12]
#define TS_A64 ro_ts[13]
#define TS_A65 ro_ts[14]
#define TS_A71 ro_ts[15]
#define TS_A72 ro_ts[16]
#define TS_A73 ro_ts[17]
#define TS_A74 ro_ts[18]
#define TS_A75 ro_ts[19]
#define TS_A76 ro_ts[20]
#define TS_BT1 ro_ts[21]
#define TS_BT2 ro_ts[22]
#define TS_BT3 ro_ts[23]
#define TS_BT4 ro_ts[24]
#define TS_BT5 ro_ts[25]
#define TS_BT6 ro_ts[26]
#define TS_BT7 ro_ts[27]

__device__ __forceinline__ double ro_eps_of(double x) {
    x = fabs(x);
    return __longlong_as_double(__double_as_longlong(x) + 1) - x;
}

// One ODE.solve(prob, Tsit5()) over [t0, t1] on the NS dynamic states (tail q constant).
// Mirrors OrdinaryDiffEq v6: initdt (order 5), PI controller (beta1 7/50, beta2 2/25, gamma 9/10,
// qmin 1/5, qmax 10), error norm over the whole extended state (the constant tail contributes 0).
template <int PLANT>
__device__ __forceinline__ void ro_tsit5(double* x, const double* q, double t0, double t1, const RolloutArgs& A,
                                         long long cell, int& nacc, int& nrej) {
    constexpr int NS = PlantDims<PLANT>::NS, NQ = PlantDims<PLANT>::ND + PlantDims<PLANT>::NC, NE = NS + NQ;
    const double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 9.0 / 10.0, qmin = 1.0 / 5.0, qmax = 10.0;
    const double abstol = A.abstol, reltol = A.reltol;
    double k1[NS], k2[NS], k3[NS], k4[NS], k5[NS], k6[NS], k7[NS], tmp[NS], un[NS];
    double t = t0;
    const double dtmax = t1 - t0;
    int nlog = 0;
    const int log_cap = A.log_cap;
    if (log_cap > 0) {
        A.log_t[cell * log_cap] = t;
        double* lu = A.log_u + (size_t)cell * log_cap * NE;
#pragma unroll
        for (int i = 0; i < NS; i++) lu[i] = x[i];
#pragma unroll
        for (int i = 0; i < NQ; i++) lu[NS + i] = q[i];
        nlog = 1;
    }
    plant_rhs<PLANT>(x, q, k1);
    double dt;
    {   // ode_determine_initdt
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
            double sk = abstol + fabs(x[i]) * reltol;
            double a = x[i] / sk, b = k1[i] / sk;
            s0 += a * a; s1 += b * b;
        }
#pragma unroll
        for (int i = 0; i < NQ; i++) {
            double sk = abstol + fabs(q[i]) * reltol;
            double a = q[i] / sk;
            s0 += a * a; s1 += 0.0;
        }
        double d0 = sqrt(s0 / NE), d1 = sqrt(s1 / NE);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100;
        if (dt0 > dtmax) dt0 = dtmax;
        if (dt0 < 10 * 2.220446049250313e-16) {
            dt = 1e-6;
        } else {
#pragma unroll
            for (int i = 0; i < NS; i++) un[i] = x[i] + dt0 * k1[i];
            plant_rhs<PLANT>(un, q, k2);
            double s2 = 0.0;
#pragma unroll
            for (int i = 0; i < NS; i++) {
                double sk = abstol + fabs(x[i]) * reltol;
                double a = (k2[i] - k1[i]) / sk;
                s2 += a * a;
            }
            double d2 = sqrt(s2 / NE) / dt0;
            double md = d1 > d2 ? d1 : d2;
            double dt1 = md <= 1e-15 ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2 + log10(md)) / 5);
            dt = fmin(fmin(100 * dt0, dt1), dtmax);
        }
    }
    double qold = 1e-4;
    // OrdinaryDiffEq stops a solve with retcode MaxIters after maxiters = 1e5 iterations, with Unstable / DtLessThanMin when the
    // error estimate or the step is not finite or the step underflows; here the period ends and the call reports CLR_ERR_UNSTABLE
    int guard = 0;
    while (t < t1) {
        if (++guard > 100000 || !(fabs(dt) > ro_eps_of(t)) || !isfinite(dt)) { atomicCAS(A.error, 0, (int)CLR_ERR_UNSTABLE); break; }
        if (dt > t1 - t) dt = t1 - t;
#pragma unroll
        for (int i = 0; i < NS; i++) tmp[i] = x[i] + dt * (TS_A21 * k1[i]);
        plant_rhs<PLANT>(tmp, q, k2);
#pragma unroll
        for (int i = 0; i < NS; i++) tmp[i] = x[i] + dt * (TS_A31 * k1[i] + TS_A32 * k2[i]);
        plant_rhs<PLANT>(tmp, q, k3);
#pragma unroll
        for (int i = 0; i < NS; i++) tmp[i] = x[i] + dt * (TS_A41 * k1[i] + TS_A42 * k2[i] + TS_A43 * k3[i]);
        plant_rhs<PLANT>(tmp, q, k4);
#pragma unroll
        for (int i = 0; i < NS; i++)
            tmp[i] = x[i] + dt * (TS_A51 * k1[i] + TS_A52 * k2[i] + TS_A53 * k3[i] + TS_A54 * k4[i]);
        plant_rhs<PLANT>(tmp, q, k5);
#pragma unroll
        for (int i = 0; i < NS; i++)
            tmp[i] = x[i] + dt * (TS_A61 * k1[i] + TS_A62 * k2[i] + TS_A63 * k3[i] + TS_A64 * k4[i] + TS_A65 * k5[i]);
        plant_rhs<PLANT>(tmp, q, k6);
#pragma unroll
        for (int i = 0; i < NS; i++)
            un[i] = x[i] + dt * (TS_A71 * k1[i] + TS_A72 * k2[i] + TS_A73 * k3[i] + TS_A74 * k4[i] + TS_A75 * k5[i] +
                                 TS_A76 * k6[i]);
        plant_rhs<PLANT>(un, q, k7);
        double se = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
            double e = TS_BT1 * k1[i] + TS_BT2 * k2[i] + TS_BT3 * k3[i] + TS_BT4 * k4[i] + TS_BT5 * k5[i] +
                       TS_BT6 * k6[i] + TS_BT7 * k7[i];
            e = dt * e;
            double r = e / (abstol + fmax(fabs(x[i]), fabs(un[i])) * reltol);
            se += r * r;
        }
        double EEst = sqrt(se / NE);
        if (!isfinite(EEst)) { atomicCAS(A.error, 0, (int)CLR_ERR_UNSTABLE); break; }
        double qq, q11 = 0.0;
        if (EEst == 0.0) {
            qq = 1.0 / qmax;
        } else {
            q11 = ro_pow(EEst, beta1, A.pow_mode);
            qq = q11 / ro_pow(qold, beta2, A.pow_mode);
            qq = fmax(1.0 / qmax, fmin(1.0 / qmin, qq / gamma));
        }
        if (EEst <= 1.0) {
            nacc++;
            qold = fmax(EEst, 1e-4);
            double dtnew = dt / qq;
            double ttmp = t + dt;
            if (fabs(ttmp - t1) < 100 * ro_eps_of(fmax(t, t1))) ttmp = t1;
            t = ttmp;
#pragma unroll
            for (int i = 0; i < NS; i++) { x[i] = un[i]; k1[i] = k7[i]; }
            dt = fmin(dtnew, dtmax);
            if (log_cap > 0) {
                if (nlog < log_cap) {
                    A.log_t[cell * log_cap + nlog] = t;
                    double* lu = A.log_u + ((size_t)cell * log_cap + nlog) * NE;
#pragma unroll
                    for (int i = 0; i < NS; i++) lu[i] = x[i];
#pragma unroll
                    for (int i = 0; i < NQ; i++) lu[NS + i] = q[i];
                    nlog++;
                } else {
                    atomicCAS(A.error, 0, (int)CLR_ERR_CAPACITY);
                }
            }
        } else {
            nrej++;
            dt = dt / fmin(1.0 / qmin, q11 / gamma);
        }
    }
    if (log_cap > 0) A.log_n[cell] = nlog;
}

template <int PLANT>
__device__ __forceinline__ void ro_rk4(double* x, const double* q, double t0, double t1, int substeps) {
    constexpr int NS = PlantDims<PLANT>::NS;
    double k1[NS], k2[NS], k3[NS], k4[NS], tmp[NS];
    const double dt = (t1 - t0) / substeps;
    for (int s = 0; s < substeps; s++) {
        plant_rhs<PLANT>(x, q, k1);
#pragma unroll
        for (int i = 0; i < NS; i++) tmp[i] = x[i] + (dt / 2) * k1[i];
        plant_rhs<PLANT>(tmp, q, k2);
#pragma unroll
        for (int i = 0; i < NS; i++) tmp[i] = x[i] + (dt / 2) * k2[i];
        plant_rhs<PLANT>(tmp, q, k3);
#pragma unroll
        for (int i = 0; i < NS; i++) tmp[i] = x[i] + dt * k3[i];
        plant_rhs<PLANT>(tmp, q, k4);
#pragma unroll
        for (int i = 0; i < NS; i++) x[i] = x[i] + (dt / 6) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    }
}

template <int PLANT, int CODE>
__global__ void __launch_bounds__(128, 4)
rollout_kernel(PointNet net, RolloutArgs A, int in_smem) {
    constexpr int NS = PlantDims<PLANT>::NS, ND = PlantDims<PLANT>::ND, NC = PlantDims<PLANT>::NC;
    extern __shared__ __align__(16) double ro_sm[];
    const double* w = ro_stage_weights(net, ro_sm, in_smem != 0, CODE > 2);
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < A.N; j += (long long)gridDim.x * blockDim.x) {
        double x[NS], q[ND + NC];
#pragma unroll
        for (int i = 0; i < NS; i++) x[i] = A.x0[j * NS + i];
        double t = A.t0;
        for (int it = 0; it < A.iters; it++) {
            constexpr int NX = NS > 8 ? NS : 8, NU = NC > 8 ? NC : 8;
            double xin[NX], u[NU];
#pragma unroll
            for (int i = 0; i < NX; i++) xin[i] = i < NS ? x[i] : 0.0;
            ro_eval<CODE>(net, w, xin, u);
            const long long cell = (long long)it * A.N + j;
#pragma unroll
            for (int i = 0; i < NC; i++) A.U[cell * NC + i] = u[i];
#pragma unroll
            for (int i = 0; i < ND; i++) q[i] = A.Wd[cell * ND + i];
#pragma unroll
            for (int i = 0; i < NC; i++) q[ND + i] = u[i];
            const double t1 = it < A.iters - 1 ? t + A.tau : A.T;
            int nacc = 0, nrej = 0;
            if (A.alg == CLR_ALG_TSIT5) ro_tsit5<PLANT>(x, q, t, t1, A, cell, nacc, nrej);
            else { ro_rk4<PLANT>(x, q, t, t1, A.substeps); nacc = A.substeps; }
            if (A.naccept) A.naccept[cell] = nacc;
            if (A.nreject) A.nreject[cell] = nrej;
#pragma unroll
            for (int i = 0; i < NS; i++) A.X_end[cell * NS + i] = x[i];
            t = t1;
        }
    }
}

// Register-blocked evaluation of a narrow two-stage network for R points at once: every weight record is
// loaded from shared memory once and used for R trajectories, which takes the shared-memory pipe (the limiter
// of the one-point version: 4 broadcast loads per neuron) off the critical path.  Same arithmetic per point
// as ro_mlp_small.
template <int NI, int NO, int R>
__device__ __forceinline__ void ro_mlp_small_blocked(const PointNet& net, const double* __restrict__ w, double (&va)[R][8],
                                                     double (&uo)[R][8]) {
    int l = 0;
    while (l < net.L) {
        const int m = net.dims[l + 1];
        if (l + 1 < net.L) {
            constexpr int ST = NI + NO + 2;
            const double* rec = w + net.soff[l];
            const double* b2 = w + net.soff[l + 1];
            const int a1 = net.act[l], a2 = net.act[l + 1];
            const int m2 = net.dims[l + 2];
            double vb[R][NO];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int o = 0; o < NO; o++) vb[r][o] = 0.0;
            const bool relu = a1 == CLR_ACT_RELU;
#pragma unroll 2
            for (int i = 0; i < m; i++) {
                const double2* r2 = reinterpret_cast<const double2*>(rec + i * ST);
                double2 wv[NI / 2], vv[NO / 2];
#pragma unroll
                for (int k = 0; k < NI / 2; k++) wv[k] = r2[k];
#pragma unroll
                for (int o = 0; o < NO / 2; o++) vv[o] = r2[NI / 2 + o];
                const double b1 = rec[i * ST + NI + NO];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NI / 2; k++) { s = fma(wv[k].x, va[r][2 * k], s); s = fma(wv[k].y, va[r][2 * k + 1], s); }
                    s += b1;
                    const double h = relu ? ro_relu_bits(s) : ro_act(a1, s);       // max(s, 0) on the integer pipe: the FP64 pipe is the limiter
#pragma unroll
                    for (int o = 0; o < NO / 2; o++) { vb[r][2 * o] = fma(vv[o].x, h, vb[r][2 * o]); vb[r][2 * o + 1] = fma(vv[o].y, h, vb[r][2 * o + 1]); }
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int o = 0; o < 8; o++) va[r][o] = (o < NO && o < m2) ? ro_act(a2, vb[r][o < NO ? o : 0] + b2[o < NO ? o : 0]) : 0.0;
            l += 2;
        } else {
            constexpr int ST1 = NI + 2;
            const double* rec = w + net.soff[l];
            const int a1 = net.act[l];
#pragma unroll
            for (int r = 0; r < R; r++) {
                double vb[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    vb[i] = 0.0;
                    if (i < m) {
                        const double2* r2 = reinterpret_cast<const double2*>(rec + i * ST1);
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < NI / 2; k++) { const double2 wv = r2[k]; s = fma(wv.x, va[r][2 * k], s); s = fma(wv.y, va[r][2 * k + 1], s); }
                        vb[i] = ro_act(a1, s + rec[i * ST1 + NI]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; i++) va[r][i] = vb[i];
            }
            l += 1;
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int k = 0; k < 8; k++) uo[r][k] = va[r][k];
}

// Narrow-network rollouts: a thread owns RO_R trajectories.  Per period the controller is evaluated for all of
// them at once (register blocked), then each one is integrated in turn; states and controls live in
// thread-private shared-memory slots in between, so the ODE code exists once.
#ifndef RO_R
#define RO_R 4
#endif
#ifndef RO_MINBLK
#define RO_MINBLK 4
#endif
template <int PLANT, int CODE>
__global__ void __launch_bounds__(128, RO_MINBLK)
rollout_blocked_kernel(PointNet net, RolloutArgs A, int in_smem) {
    constexpr int NS = PlantDims<PLANT>::NS, ND = PlantDims<PLANT>::ND, NC = PlantDims<PLANT>::NC;
    constexpr int NI = CODE >> 4, NO = CODE & 15;
    extern __shared__ __align__(16) double ro_sm[];
    const int tid = threadIdx.x;
    double* xs = ro_sm;                                   // [RO_R][NS][128]
    double* us = xs + RO_R * NS * 128;                    // [RO_R][NC][128]
    double* wsm = us + RO_R * NC * 128;
    const double* w = ro_stage_weights(net, wsm, in_smem != 0, true);
    const long long span = 128ll * RO_R;
    for (long long base = blockIdx.x * span; base < A.N; base += (long long)gridDim.x * span) {
#pragma unroll
        for (int r = 0; r < RO_R; r++) {
            const long long j = base + r * 128 + tid;
#pragma unroll
            for (int i = 0; i < NS; i++) xs[(r * NS + i) * 128 + tid] = j < A.N ? A.x0[j * NS + i] : 0.0;
        }
        double t = A.t0;
        for (int it = 0; it < A.iters; it++) {
            {
                double va[RO_R][8], uo[RO_R][8];
#pragma unroll
                for (int r = 0; r < RO_R; r++)
#pragma unroll
                    for (int i = 0; i < 8; i++) va[r][i] = i < NS ? xs[(r * NS + (i < NS ? i : 0)) * 128 + tid] : 0.0;
                ro_mlp_small_blocked<NI, NO, RO_R>(net, w, va, uo);
#pragma unroll
                for (int r = 0; r < RO_R; r++)
#pragma unroll
                    for (int i = 0; i < NC; i++) us[(r * NC + i) * 128 + tid] = uo[r][i];
            }
            const double t1 = it < A.iters - 1 ? t + A.tau : A.T;
            for (int r = 0; r < RO_R; r++) {
                const long long j = base + r * 128 + tid;
                if (j >= A.N) continue;
                double x[NS], q[ND + NC];
#pragma unroll
                for (int i = 0; i < NS; i++) x[i] = xs[(r * NS + i) * 128 + tid];
                const long long cell = (long long)it * A.N + j;
#pragma unroll
                for (int i = 0; i < NC; i++) { q[ND + i] = us[(r * NC + i) * 128 + tid]; A.U[cell * NC + i] = q[ND + i]; }
#pragma unroll
                for (int i = 0; i < ND; i++) q[i] = A.Wd[cell * ND + i];
                int nacc = 0, nrej = 0;
                if (A.alg == CLR_ALG_TSIT5) ro_tsit5<PLANT>(x, q, t, t1, A, cell, nacc, nrej);
                else { ro_rk4<PLANT>(x, q, t, t1, A.substeps); nacc = A.substeps; }
                if (A.naccept) A.naccept[cell] = nacc;
                if (A.nreject) A.nreject[cell] = nrej;
#pragma unroll
                for (int i = 0; i < NS; i++) { A.X_end[cell * NS + i] = x[i]; xs[(r * NS + i) * 128 + tid] = x[i]; }
            }
            t = t1;
        }
    }
}

// ---- split rollouts: one controller launch + one plant launch per control period -----------------------------------
// The fused kernels above evaluate a narrow controller from 128-bit broadcast loads of its packed records: four loads
// per hidden neuron and warp = 14 cycles of the SM's 128 B/clk load-return path against the 56 cycles the neuron's 28
// FP64 instructions take on one of the four sub-partitions -- with four sub-partitions the load path is as busy as the
// FP64 pipe, and the adaptive integration that follows in the same kernel pins the register budget at 16 warps per SM.
// Split, the controller kernel runs in warp-uniform control flow throughout, so ptxas keeps the record pointer in a
// uniform register and fetches the weights from the constant bank (ro_cw) into UNIFORM registers (LDCU: one copy per
// warp instead of 32, no load-return traffic) that feed the DFMAs directly; the plant kernel owns one trajectory per
// thread at its own, higher occupancy.  The state travels between the two through the X_end / U arrays the call
// returns anyway (64 B read back per trajectory and period).
#ifndef RO_CW_DOUBLES
#define RO_CW_DOUBLES 7168
#endif
#ifndef __CUDACC_RTC__
__constant__ double ro_cw[RO_CW_DOUBLES];

__device__ __noinline__ double ro_act_cold(int act, double s) { return ro_act(act, s); }

template <int NI, int NO, int R, bool RELU, bool AHEAD>
__device__ __forceinline__ void cw_pair(int off, int m, int a1, const double (&va)[R][8], double (&vb)[R][NO]) {
    constexpr int ST = NI + NO + 2;
    double nx[ST - 1];
    if (AHEAD) {
#pragma unroll
        for (int k = 0; k < ST - 1; k++) nx[k] = ro_cw[off + k];
    }
#pragma unroll 2
    for (int i = 0; i < m; i++) {
        double rc[ST - 1];
        if (AHEAD) {     // the record of neuron i + 1 is requested before neuron i is evaluated (after the last one: whatever
                         // follows in the bank, unused -- the host keeps one record of head-room behind every image)
            const double* rec = ro_cw + off + (i + 1) * ST;
#pragma unroll
            for (int k = 0; k < ST - 1; k++) { rc[k] = nx[k]; nx[k] = rec[k]; }
        } else {
            const double* rec = ro_cw + off + i * ST;
#pragma unroll
            for (int k = 0; k < ST - 1; k++) rc[k] = rec[k];
        }
        double s[R];
#pragma unroll
        for (int r = 0; r < R; r++) s[r] = 0.0;
#pragma unroll
        for (int k = 0; k < NI; k++) {
#pragma unroll
            for (int r = 0; r < R; r++) s[r] = fma(rc[k], va[r][k], s[r]);
        }
#pragma unroll
        for (int r = 0; r < R; r++) { s[r] += rc[NI + NO]; s[r] = RELU ? ro_relu_bits(s[r]) : ro_act_cold(a1, s[r]); }
#pragma unroll
        for (int o = 0; o < NO; o++) {
#pragma unroll
            for (int r = 0; r < R; r++) vb[r][o] = fma(rc[NI + o], s[r], vb[r][o]);
        }
    }
}

// layers L0, L0 + 1 (a pair) or the single last layer L0; L0 is a literal so that every table index is one too: ptxas then
// sees uniform addresses and keeps the record pointer in a uniform register
template <int NI, int NO, int R, int L0, bool AHEAD>
__device__ __forceinline__ void cw_stage(const PointNet& net, double (&va)[R][8]) {
    if (net.L >= L0 + 2) {
        const int m = net.dims[L0 + 1], m2 = net.dims[L0 + 2];
        const int a1 = net.act[L0], a2 = net.act[L0 + 1];
        const double* b2 = ro_cw + net.soff[L0 + 1];
        double vb[R][NO];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int o = 0; o < NO; o++) vb[r][o] = 0.0;
        if (a1 == CLR_ACT_RELU) cw_pair<NI, NO, R, true, AHEAD>(net.soff[L0], m, a1, va, vb);
        else cw_pair<NI, NO, R, false, AHEAD>(net.soff[L0], m, a1, va, vb);
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int o = 0; o < 8; o++) va[r][o] = (o < NO && o < m2) ? ro_act_cold(a2, vb[r][o < NO ? o : 0] + b2[o < NO ? o : 0]) : 0.0;
    } else if (net.L == L0 + 1) {
        constexpr int ST1 = NI + 2;
        const int m = net.dims[L0 + 1], a1 = net.act[L0];
        const double* rec = ro_cw + net.soff[L0];
#pragma unroll
        for (int r = 0; r < R; r++) {
            double vb[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                vb[i] = 0.0;
                if (i < m) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NI; k++) s = fma(rec[i * ST1 + k], va[r][k], s);
                    vb[i] = ro_act_cold(a1, s + rec[i * ST1 + NI]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) va[r][i] = vb[i];
        }
    }
}

// U[j] = network(X[j]) for R points per thread (same arithmetic per point as ro_mlp_small); X: N x nx, U: N x m_out.
// Networks of at most RO_CW_LAYERS layers (the host checks).
#define RO_CW_LAYERS 4
template <int NI, int NO, int R, int MINBLK, bool AHEAD>
__global__ void __launch_bounds__(128, MINBLK)
controller_cw_kernel(PointNet net, int nx, long long N, const double* __restrict__ X, double* __restrict__ U) {
    const long long base = blockIdx.x * (128ll * R) + threadIdx.x;
    double va[R][8];
#pragma unroll
    for (int r = 0; r < R; r++) {
        long long j = base + r * 128;
        if (j >= N) j = N - 1;
#pragma unroll
        for (int k = 0; k < 8; k++) va[r][k] = (k < NI && k < nx) ? X[j * nx + (k < NI ? k : 0)] : 0.0;
    }
    cw_stage<NI, NO, R, 0, AHEAD>(net, va);
    cw_stage<NI, NO, R, 2, AHEAD>(net, va);
    const int mo = net.m_out;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const long long j = base + r * 128;
        if (j < N) {
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < mo) U[j * mo + k] = va[r][k];
        }
    }
}
#endif  // !__CUDACC_RTC__

// One control period of the plant for every trajectory: x(t) from Xprev (x0 or the previous period's X_end), the
// control from U (written by the controller launch of this period), the disturbance from Wd.
#ifndef RO_ODE_MINBLK
#define RO_ODE_MINBLK 5
#endif
template <int PLANT>
__global__ void __launch_bounds__(128, RO_ODE_MINBLK)
rollout_ode_kernel(RolloutArgs A, int it, double t, double t1, const double* __restrict__ Xprev) {
    constexpr int NS = PlantDims<PLANT>::NS, ND = PlantDims<PLANT>::ND, NC = PlantDims<PLANT>::NC;
    const long long j = blockIdx.x * 128ll + threadIdx.x;
    if (j >= A.N) return;
    const long long cell = (long long)it * A.N + j;
    double x[NS], q[ND + NC];
#pragma unroll
    for (int i = 0; i < NS; i++) x[i] = Xprev[j * NS + i];
#pragma unroll
    for (int i = 0; i < ND; i++) q[i] = A.Wd[cell * ND + i];
#pragma unroll
    for (int i = 0; i < NC; i++) q[ND + i] = A.U[cell * NC + i];
    int nacc = 0, nrej = 0;
    if (A.alg == CLR_ALG_TSIT5) ro_tsit5<PLANT>(x, q, t, t1, A, cell, nacc, nrej);
    else { ro_rk4<PLANT>(x, q, t, t1, A.substeps); nacc = A.substeps; }
    if (A.naccept) A.naccept[cell] = nacc;
    if (A.nreject) A.nreject[cell] = nrej;
#pragma unroll
    for (int i = 0; i < NS; i++) A.X_end[cell * NS + i] = x[i];
}
