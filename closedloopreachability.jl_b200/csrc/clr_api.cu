// clr_api.cu -- host side of libclr_b200.so: the C ABI declared in include/clr_b200.h.
//
// Everything here is plumbing around the kernels in deepz.cuh / rollout.cuh: device buffers,
// the folding of pre/post-processing (src/problem.jl:5-77) into extra affine layers, the DMMA
// fragment layout of the weights, the launch sequence (main pass -> retry pass -> big-cell pass)
// and the host <-> device copies of the blocking (non CLR_FLAG_DEVICE_IO) mode.
// There is no CPU fallback: without a device clr_create fails.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <string>
#include <vector>

#include "deepz_launch.h"
#include "rollout.cuh"
#include "reduce.cuh"
#include "project.cuh"
#include "split.cuh"
#include "tiny.cuh"

#define CLR_BOUNCE_BYTES (1u << 20)  // pinned bounce buffer of the small-call path
#define CLR_BOUNCE_IN_MAX (256u << 10)
#define CLR_PILOT_CELLS 2048        // cells of the pilot run that orders the neurons
#define CLR_REORDER_MIN_CELLS 16384 // calls with fewer cells keep the network's own neuron order

namespace {

const int kPlantCodes[5] = {66, 72, 136, 1, 2};   // evaluator codes of the rollout kernels (rollout.cuh: ro_eval)

thread_local std::string g_create_error;

struct HostLayer {
    int n, m, act, collapse, net_layer, project = 0;
    std::vector<double> W;   // m x n column-major
    std::vector<double> b;   // m
};

struct Compiled {            // one (network, pre/post-processing) pair on the device
    std::vector<unsigned char> key;
    std::vector<void*> allocs;
    DevNet host{};
    DevNet* dev = nullptr;
    PointNet pnet{};
    int ksreg = 8;
    int stored_width = 0;
    int point_code = 1;
    int inplace = 1;
    std::vector<HostLayer> layers_folded;     // the layer list this image was built from
    Compiled* perm = nullptr;                 // the same network with reordered neurons (see compile_net_reordered)
    std::vector<unsigned char> pilot_key;     // the split `perm` was derived for: the pilot is skipped while it stays the same
};

}  // namespace

struct clr_net {
    clr_ctx* ctx;
    int device = 0;                           // kept apart from ctx: clr_net_free must work after clr_destroy(ctx)
    int L;
    std::vector<int> dims, act;
    std::vector<std::vector<double>> W, b;
    // device images per (pre/post-processing, input dimension), most recently used first: DeepZ and rollout calls with
    // different processing alternate without recompiling (and without the device synchronisation a cudaFree implies)
    std::vector<Compiled*> variants;
};
#define CLR_NET_VARIANTS 4

// a run-time compiled plant right-hand side (NVRTC -> cubin -> cudaLibrary), one module per evaluator code, built on demand
struct clr_plant {
    clr_ctx* ctx = nullptr;
    int ns = 0, nd = 0, nc = 0;
    std::string source;             // prelude + rollout.cuh + the user's right-hand side
    cudaLibrary_t lib[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};      // slot 5: rollout_ode_kernel (split rollouts)
    cudaKernel_t kern[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct clr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int sms = 0;
    int smem_optin = 0;
    int tile_cells = 0;
    DevCallState* d_state = nullptr;
    DevCallState* h_state = nullptr;     // pinned
    long long* d_retry = nullptr;
    long long* d_big = nullptr;
    int64_t list_cap = 0;
    double* d_ws = nullptr;
    size_t ws_bytes = 0;
    // scratch for blocking-mode I/O
    void* d_io[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t io_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // small-call path of the blocking mode: one pinned bounce buffer and its device twin; the inputs of a call (and
    // the hull seed) go up in ONE copy, the boxes and the hull come back in ONE copy
    unsigned char* h_bounce = nullptr;   // pinned, mapped: m_bounce is its address as the device sees it
    unsigned char* d_bounce = nullptr;
    unsigned char* m_bounce = nullptr;
    int zero_copy = 1;                   // CLR_B200_ZEROCOPY=0: small calls of narrow networks copy their inputs / results like every other call
    size_t bounce_used = 0, bounce_flushed = 0;
    bool bounce_open = false;
    double* d_hull = nullptr;            // 2 * m_out
    int hull_cap = 0;
    // kept output zonotopes
    int* d_keepN = nullptr;
    double* d_keepC = nullptr;
    double* d_keepG = nullptr;
    size_t keepN_bytes = 0, keepC_bytes = 0, keepG_bytes = 0;
    int64_t keep_cells = 0;
    int keep_cap = 0, keep_m = 0;
    int* d_err = nullptr;
    int* d_chk = nullptr;                // [max, min] generator counts of device-resident batches
    // pilot run of the neuron reordering (see reorder_for_grid)
    DevCallState* d_state2 = nullptr;
    int* d_live = nullptr;
    int* h_live = nullptr;               // pinned
    long long* d_pilot = nullptr;
    int reorder = 1;                     // clr_set_reorder
    int wide_mode = 0;                   // CLR_B200_WIDE=1: the several-threads-per-neuron relaxation in the lean kernel too (experiments)
    bool debug = false;                  // CLR_B200_DEBUG: one line per DeepZ launch on stderr
    int tiny = 1;                        // clr_set_tiny: narrow networks take the one-warp-per-cell kernel
    int reduce_reg = 1;                  // CLR_B200_REDUCE_REG=0: the order reduction with the shared-memory sort for every width (A/B, tests)
    int ro_split = 1;                    // CLR_B200_RO_SPLIT=0: narrow controllers stay in the fused rollout kernel (experiments)
    // optional per-pass timing of the last DeepZ call (clr_set_profiling)
    int profiling = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float pass_ms[3] = {0.f, 0.f, 0.f};
    int64_t last_retry = 0, last_big = 0;
};

namespace {

int fail(clr_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, CLR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                                         \
    } while (0)

int ensure(clr_ctx* ctx, void** p, size_t* have, size_t need) {
    if (need <= *have && *p) return 0;
    if (*p) CU(cudaFree(*p));
    *p = nullptr; *have = 0;
    size_t sz = need + need / 4 + 256;
    CU(cudaMalloc(p, sz));
    *have = sz;
    return 0;
}

// reserve `bytes` in the bounce buffer (host pointer returned, device twin at the same offset); nullptr when it does not fit
void* bounce_take(clr_ctx* ctx, size_t bytes, size_t* off) {
    const size_t o = (ctx->bounce_used + 255) & ~(size_t)255;
    if (!ctx->bounce_open || o + bytes > CLR_BOUNCE_BYTES) return nullptr;
    ctx->bounce_used = o + bytes;
    *off = o;
    return ctx->h_bounce + o;
}
// everything staged so far goes to the device (one copy for all inputs of a small call)
int bounce_flush(clr_ctx* ctx) {
    if (ctx->bounce_used > ctx->bounce_flushed) {
        const size_t a = ctx->bounce_flushed & ~(size_t)255;
        CU(cudaMemcpyAsync(ctx->d_bounce + a, ctx->h_bounce + a, ctx->bounce_used - a, cudaMemcpyHostToDevice, ctx->stream));
        ctx->bounce_flushed = ctx->bounce_used;
    }
    return 0;
}

__global__ void fill_hull_kernel(double* hull, int m) {
    int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < m) { hull[i] = INFINITY; hull[m + i] = -INFINITY; }
}

// clr_measure_fp64_peak: 8 independent DMMA chains per warp, operands in registers
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a0, double b0) {
    double d[8][2];
#pragma unroll
    for (int c = 0; c < 8; c++) { d[c][0] = threadIdx.x; d[c][1] = c; }
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d[c][0]), "+d"(d[c][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) s += d[c][0] + d[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// generator counts of device-resident zonotope batches: out[0] = max, out[1] = min (a negative minimum = decreasing offsets)
__global__ void max_cols_kernel(const long long* col_off, const int* ncols, long long N, int* out) {
    int best = 0, worst = 0x7fffffff;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const long long d = col_off ? col_off[i + 1] - col_off[i] : (long long)ncols[i];
        const int p = d > 0x7fffffffLL ? 0x7fffffff : (d < -0x7fffffffLL ? -0x7fffffff : (int)d);
        best = p > best ? p : best;
        worst = p < worst ? p : worst;
    }
    atomicMax(&out[0], best);
    atomicMin(&out[1], worst);
    if (col_off && blockIdx.x == 0 && threadIdx.x == 0 && col_off[0] != 0) atomicMin(&out[1], -1);
}

void free_compiled(Compiled* c) {
    if (!c) return;
    free_compiled(c->perm);
    for (void* p : c->allocs) cudaFree(p);
    delete c;
}

// The layer list the kernels run: [pre] + network + [post]   (src/solve.jl:211-213)
int fold_layers(clr_ctx* ctx, const clr_net* net, const clr_prepost* pp, int n_in, std::vector<HostLayer>& out) {
    int cur = n_in;
    if (pp && pp->pre_rows > 0) {
        if (!pp->preA || !pp->preb) return fail(ctx, CLR_ERR_ARG, "pre_rows > 0 needs preA and preb");
        HostLayer h;
        h.n = n_in; h.m = pp->pre_rows; h.act = CLR_ACT_ID; h.collapse = pp->pre_rows == 1; h.net_layer = -1;
        h.W.assign(pp->preA, pp->preA + (size_t)h.m * h.n);
        h.b.assign(pp->preb, pp->preb + h.m);
        out.push_back(std::move(h));
        cur = pp->pre_rows;
    }
    if (cur != net->dims[0])
        return fail(ctx, CLR_ERR_DIM, "input dimension %d does not match the network input dimension %d", cur, net->dims[0]);
    for (int l = 0; l < net->L; l++) {
        HostLayer h;
        h.n = net->dims[l]; h.m = net->dims[l + 1]; h.act = net->act[l]; h.collapse = h.m == 1; h.net_layer = l;
        h.W = net->W[l]; h.b = net->b[l];
        out.push_back(std::move(h));
    }
    const int m = net->dims[net->L];
    const int kind = pp ? pp->post_kind : CLR_POST_NONE;
    if (kind == CLR_POST_SHIFT || kind == CLR_POST_SCALE) {
        if ((kind == CLR_POST_SHIFT && !pp->postb) || (kind == CLR_POST_SCALE && !pp->postA))
            return fail(ctx, CLR_ERR_ARG, "post-processing value missing");
        HostLayer h;
        h.n = m; h.m = m; h.act = CLR_ACT_ID; h.collapse = 0; h.net_layer = -1;
        h.W.assign((size_t)m * m, 0.0);
        h.b.assign(m, kind == CLR_POST_SHIFT ? pp->postb[0] : 0.0);
        for (int i = 0; i < m; i++) h.W[(size_t)i * m + i] = kind == CLR_POST_SHIFT ? 1.0 : pp->postA[0];
        out.push_back(std::move(h));
    } else if (kind == CLR_POST_MATRIX) {
        if (!pp->postA || pp->post_rows < 1) return fail(ctx, CLR_ERR_ARG, "post matrix missing");
        HostLayer h;
        h.n = m; h.m = pp->post_rows; h.act = CLR_ACT_ID; h.collapse = h.m == 1; h.net_layer = -1;
        h.W.assign(pp->postA, pp->postA + (size_t)h.m * m);
        h.b.assign(h.m, 0.0);
        out.push_back(std::move(h));
    } else if (kind == CLR_POST_PROJECT) {
        if (!pp->post_dims || pp->post_rows < 1) return fail(ctx, CLR_ERR_ARG, "post projection dims missing");
        HostLayer h;
        h.n = m; h.m = pp->post_rows; h.act = CLR_ACT_ID; h.collapse = 0; h.net_layer = -1; h.project = 1;
        h.W.assign((size_t)h.m * m, 0.0);
        h.b.assign(h.m, 0.0);
        for (int i = 0; i < h.m; i++) {
            int d = pp->post_dims[i];
            if (d < 0 || d >= m) return fail(ctx, CLR_ERR_DIM, "projection dimension %d out of range", d);
            h.W[(size_t)d * h.m + i] = 1.0;
        }
        out.push_back(std::move(h));
    } else if (kind != CLR_POST_NONE) {
        return fail(ctx, CLR_ERR_ARG, "unknown post_kind %d", kind);
    }
    return 0;
}

std::vector<unsigned char> prepost_key(const clr_net* net, const clr_prepost* pp, int n_in) {
    std::vector<unsigned char> k;
    auto put = [&](const void* p, size_t n) { const unsigned char* c = (const unsigned char*)p; k.insert(k.end(), c, c + n); };
    put(&n_in, sizeof(int));
    if (!pp) return k;
    put(&pp->pre_rows, sizeof(int));
    put(&pp->post_kind, sizeof(int));
    put(&pp->post_rows, sizeof(int));
    if (pp->pre_rows > 0 && pp->preA && pp->preb) {
        put(pp->preA, sizeof(double) * (size_t)pp->pre_rows * n_in);
        put(pp->preb, sizeof(double) * pp->pre_rows);
    }
    const int m = net->dims[net->L];
    if (pp->post_kind == CLR_POST_SHIFT && pp->postb) put(pp->postb, sizeof(double));
    if (pp->post_kind == CLR_POST_SCALE && pp->postA) put(pp->postA, sizeof(double));
    if (pp->post_kind == CLR_POST_MATRIX && pp->postA) put(pp->postA, sizeof(double) * (size_t)pp->post_rows * m);
    if (pp->post_kind == CLR_POST_PROJECT && pp->post_dims) put(pp->post_dims, sizeof(int) * pp->post_rows);
    return k;
}

template <typename T>
int upload(clr_ctx* ctx, Compiled* c, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    size_t bytes = sizeof(T) * (v.empty() ? 1 : v.size());
    CU(cudaMalloc(&p, bytes));
    c->allocs.push_back(p);
    if (!v.empty()) CU(cudaMemcpyAsync(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out = (const T*)p;
    return 0;
}

// Device images of a folded layer list: DMMA fragments, biases, pointwise images.
int compile_layers(clr_ctx* ctx, const std::vector<HostLayer>& layers, int n_in, const std::vector<unsigned char>& key,
                   Compiled** outc) {
    int rc;
    if ((int)layers.size() > CLR_DEV_MAX_LAYERS) return fail(ctx, CLR_ERR_ARG, "too many layers");
    Compiled* c = new Compiled();
    c->key = key;
    DevNet& D = c->host;
    D.L = (int)layers.size();
    D.n_in = n_in;
    D.m_out = layers.back().m;
    D.max_rows = n_in;
    D.max_pitch = clr_pitch(n_in);
    int maxKsReg = 1, maxRB = 1;
    PointNet& P = c->pnet;
    P.L = D.L; P.n_in = n_in; P.m_out = D.m_out;
    P.dims[0] = n_in;
    std::vector<double> image;
    for (int l = 0; l < D.L; l++) {
        const HostLayer& h = layers[l];
        if (h.m > CLR_MAX_ROWS || h.n > CLR_MAX_ROWS * 8) {
            free_compiled(c);
            return fail(ctx, CLR_ERR_ARG, "layer %d is %d x %d; this build supports at most %d rows", l, h.m, h.n, CLR_MAX_ROWS);
        }
        DevLayer& d = D.layer[l];
        d.n = h.n; d.m = h.m;
        d.ksteps = (h.n + 3) / 4;
        d.rblocks = (h.m + 7) / 8;
        d.act = h.act; d.collapse = h.collapse; d.net_layer = h.net_layer;
        d.remove_zero = h.act != CLR_ACT_ID || h.project;
        D.max_rows = std::max(D.max_rows, h.m);
        D.max_pitch = std::max(D.max_pitch, clr_pitch(h.m));
        if (d.ksteps <= 32) maxKsReg = std::max(maxKsReg, d.ksteps);
        maxRB = std::max(maxRB, d.rblocks);
        // DMMA A fragments: [rb][ks][lane] = W[rb*8 + lane/4][ks*4 + lane%4]
        std::vector<double> af((size_t)d.rblocks * d.ksteps * 32, 0.0);
        for (int rb = 0; rb < d.rblocks; rb++)
            for (int ks = 0; ks < d.ksteps; ks++)
                for (int lane = 0; lane < 32; lane++) {
                    int r = rb * 8 + (lane >> 2), k = ks * 4 + (lane & 3);
                    if (r < h.m && k < h.n) af[((size_t)rb * d.ksteps + ks) * 32 + lane] = h.W[(size_t)k * h.m + r];
                }
        if ((rc = upload(ctx, c, af, &d.Afrag))) { free_compiled(c); return rc; }
        if ((rc = upload(ctx, c, h.b, &d.bias))) { free_compiled(c); return rc; }
        if ((rc = upload(ctx, c, h.W, &d.Wcm))) { free_compiled(c); return rc; }
        // pointwise image: row-major W, column-major W, bias
        P.dims[l + 1] = h.m; P.act[l] = h.act;
        P.off_rm[l] = (int)image.size();
        for (int i = 0; i < h.m; i++) for (int k = 0; k < h.n; k++) image.push_back(h.W[(size_t)k * h.m + i]);
        P.off_cm[l] = (int)image.size();
        image.insert(image.end(), h.W.begin(), h.W.end());
        P.off_b[l] = (int)image.size();
        image.insert(image.end(), h.b.begin(), h.b.end());
        P.Wcm[l] = d.Wcm; P.b[l] = d.bias; P.Wrm[l] = nullptr;
    }
    for (int l = 0; l < D.L; l++) { D.layer[l].pitch_in = D.max_pitch; D.layer[l].pitch_out = D.max_pitch; }
    P.total_doubles = (int)image.size();
    if ((rc = upload(ctx, c, image, &P.image))) { free_compiled(c); return rc; }
    c->stored_width = ro_stored_width(P.dims, P.L);
    P.small_ni = P.small_no = P.stotal = 0;
    P.simage = nullptr;
    c->point_code = c->stored_width <= 128 ? 1 : 2;
    if (c->stored_width <= 8) {
        // packed image for narrow networks (ro_mlp_small): widths of the materialised vectors
        int win = n_in, wout = 0;
        for (int l = 0; l < P.L;) {
            if (l + 1 < P.L) { wout = std::max(wout, P.dims[l + 2]); win = std::max(win, P.dims[l + 2]); l += 2; }
            else { wout = std::max(wout, P.dims[l + 1]); win = std::max(win, P.dims[l + 1]); l += 1; }
        }
        const int NI = win <= 4 ? 4 : 8;
        const int NO = (NI == 4 && wout <= 2) ? 2 : 8;
        P.small_ni = NI; P.small_no = NO;
        c->point_code = NI * 16 + NO;
        std::vector<double> simg;
        for (int l = 0; l < P.L;) {
            const HostLayer& h1 = layers[l];
            if (l + 1 < P.L) {
                const HostLayer& h2 = layers[l + 1];
                const int ST = NI + NO + 2;
                P.soff[l] = (int)simg.size();
                simg.resize(simg.size() + (size_t)h1.m * ST, 0.0);
                double* rec = simg.data() + P.soff[l];
                for (int i = 0; i < h1.m; i++) {
                    for (int k = 0; k < h1.n; k++) rec[i * ST + k] = h1.W[(size_t)k * h1.m + i];
                    for (int o = 0; o < h2.m; o++) rec[i * ST + NI + o] = h2.W[(size_t)i * h2.m + o];
                    rec[i * ST + NI + NO] = h1.b[i];
                }
                P.soff[l + 1] = (int)simg.size();
                simg.resize(simg.size() + NO + (NO & 1), 0.0);
                for (int o = 0; o < h2.m; o++) simg[P.soff[l + 1] + o] = h2.b[o];
                l += 2;
            } else {
                const int ST1 = NI + 2;
                P.soff[l] = (int)simg.size();
                simg.resize(simg.size() + (size_t)h1.m * ST1, 0.0);
                double* rec = simg.data() + P.soff[l];
                for (int i = 0; i < h1.m; i++) {
                    for (int k = 0; k < h1.n; k++) rec[i * ST1 + k] = h1.W[(size_t)k * h1.m + i];
                    rec[i * ST1 + NI] = h1.b[i];
                }
                l += 1;
            }
        }
        P.stotal = (int)simg.size();
        static std::atomic<long long> next_serial{1};
        P.sserial = next_serial.fetch_add(1);
        if ((rc = upload(ctx, c, simg, &P.simage))) { free_compiled(c); return rc; }
    }
    const int ks_opts[4] = {8, 16, 25, 32};
    c->ksreg = 32;
    for (int o = 0; o < 4; o++) if (maxKsReg <= ks_opts[o]) { c->ksreg = ks_opts[o]; break; }
    c->inplace = maxRB <= DZ_THREADS / 32;   // every warp owns at most one row block
    {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, sizeof(DevNet));
        if (e != cudaSuccess) { free_compiled(c); return fail(ctx, CLR_ERR_CUDA, "cudaMalloc(DevNet): %s", cudaGetErrorString(e)); }
        c->allocs.push_back(p);
        c->dev = (DevNet*)p;
        e = cudaMemcpy(p, &D, sizeof(DevNet), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { free_compiled(c); return fail(ctx, CLR_ERR_CUDA, "cudaMemcpy(DevNet): %s", cudaGetErrorString(e)); }
    }
    *outc = c;
    return 0;
}

int compile_net(clr_ctx* ctx, clr_net* net, const clr_prepost* pp, int n_in, Compiled** outc) {
    std::vector<unsigned char> key = prepost_key(net, pp, n_in);
    for (size_t i = 0; i < net->variants.size(); i++)
        if (net->variants[i]->key == key) {
            Compiled* c = net->variants[i];
            net->variants.erase(net->variants.begin() + (long)i);
            net->variants.insert(net->variants.begin(), c);
            *outc = c;
            return 0;
        }
    std::vector<HostLayer> layers;
    int rc = fold_layers(ctx, net, pp, n_in, layers);
    if (rc) return rc;
    Compiled* c = nullptr;
    if ((rc = compile_layers(ctx, layers, n_in, key, &c))) return rc;
    c->layers_folded = layers;
    net->variants.insert(net->variants.begin(), c);
    if (net->variants.size() > CLR_NET_VARIANTS) { free_compiled(net->variants.back()); net->variants.pop_back(); }
    *outc = c;
    return 0;
}

// The same network with the neurons of every interior layer reordered (rows of W_l and b_l, columns of W_{l+1}):
// an equivalent parametrisation; `order[l][i]` = original index of the neuron placed at position i of layer l.
int compile_net_reordered(clr_ctx* ctx, Compiled* base, int n_in, const std::vector<std::vector<int>>& order, Compiled** outc) {
    std::vector<unsigned char> key = base->key;
    for (const auto& o : order) {
        const unsigned char* p = (const unsigned char*)o.data();
        key.insert(key.end(), p, p + sizeof(int) * o.size());
        key.push_back(0xFF);
    }
    if (base->perm && base->perm->key == key) { *outc = base->perm; return 0; }
    std::vector<HostLayer> layers = base->layers_folded;
    for (size_t l = 0; l + 1 < layers.size(); l++) {
        const std::vector<int>& o = order[l];
        if (o.empty()) continue;
        HostLayer& A = layers[l];
        HostLayer& B = layers[l + 1];
        const std::vector<double> W = A.W, b = A.b, W2 = B.W;
        for (int i = 0; i < A.m; i++) {
            A.b[i] = b[o[i]];
            for (int k = 0; k < A.n; k++) A.W[(size_t)k * A.m + i] = W[(size_t)k * A.m + o[i]];
            for (int r = 0; r < B.m; r++) B.W[(size_t)i * B.m + r] = W2[(size_t)o[i] * B.m + r];
        }
    }
    Compiled* c = nullptr;
    int rc = compile_layers(ctx, layers, n_in, key, &c);
    if (rc) return rc;
    free_compiled(base->perm);
    base->perm = c;
    *outc = c;
    return 0;
}

// worst-case columns of one cell anywhere in the network, given p0 input generators
int worst_cols(const DevNet& D, int p0) {
    int p = p0, worst = 1 + p0;
    for (int l = 0; l < D.L; l++) {
        const DevLayer& d = D.layer[l];
        if (d.collapse) p = 1;
        if (d.act != CLR_ACT_ID) p += d.m;
        worst = std::max(worst, 1 + p);
    }
    return worst;
}

}  // namespace

namespace {

// Tile metadata of one launch (see DeepzSched): the appended-generator masks and the position table are sized for the network
// and the shared memory at hand instead of for the largest supported tile (64 cells x 512 rows x 2048 columns) -- every KB
// saved is column store.  ncol_need > 0 fixes the position table (global-store pass: cells wider than an SM's store).
// Returns the bytes of dynamic shared memory in front of the column store.
#define CLR_ASTAGE_MAX (48u << 10)   // largest W-fragment image staged in shared memory by TMA (wide-K layers)
size_t dz_geom(const DevNet& D, bool stats, size_t smem_total, int ncol_need, DeepzSched& sp, int ksreg = 32) {
    const size_t fixed0 = dz_meta_bytes(stats);
    // TMA staging area of the largest wide-K fragment image that fits the budget (DeepzSched::astage_*), behind the fixed metadata
    size_t astage = 0;
    for (int l = 0; l < D.L; l++) {
        const size_t img = (size_t)D.layer[l].rblocks * D.layer[l].ksteps * 256;
        if (D.layer[l].ksteps > ksreg && img <= CLR_ASTAGE_MAX) astage = std::max(astage, img);
    }
    sp.astage_off = (int)fixed0; sp.astage_bytes = (int)astage;
    const size_t fixed = fixed0 + astage;
    sp.mask_words = std::max(1, (D.max_rows + 31) / 32);
    sp.mask_stride = sp.mask_words | 1;
    const long long avail_cols = smem_total > fixed ? (long long)((smem_total - fixed) / 8 / (size_t)D.max_pitch) : 0;
    sp.tile_max = (int)std::max<long long>(4, std::min<long long>(CLR_TILE_CELLS_MAX, avail_cols / (sp.grow_cols + 2)));
    sp.ncol_max = ncol_need > 0 ? std::min(CLR_NCOL_MAX, ncol_need)
                                : (int)std::min<long long>(CLR_NCOL_MAX, ((avail_cols + 40 + 7) / 8) * 8);
    size_t off = fixed + sizeof(unsigned) * 2 * (size_t)sp.tile_max * sp.mask_stride + sizeof(unsigned short) * ((size_t)sp.ncol_max + 32);
    off = (off + 15) & ~(size_t)15;
    sp.store_off = (int)off;
    return off;
}

int launch_deepz(clr_ctx* ctx, bool global, bool stats, const Compiled* c, int grid, int threads, size_t smem,
                 const DeepzInput& in, const DeepzOutput& out, const DeepzSched& sp) {
    cudaError_t e;
    const int ksreg = c->ksreg;
    const DevNet* net = c->dev;
    bool full = false;     // sigmoid / tanh layers or streamed products: the full instantiation
    for (int l = 0; l < c->host.L; l++) {
        const DevLayer& d = c->host.layer[l];
        full = full || d.act == CLR_ACT_SIGMOID || d.act == CLR_ACT_TANH || d.ksteps > ksreg;
    }
    full = full || !c->inplace;                 // two-buffer products (a layer wider than 128 neurons)
    // the several-threads-per-neuron relaxation: measured to pay in the global-store pass only (one wide cell per tile, every
    // access an L2 round trip: 136 -> 124 ms in the dense regime of C4), not in shared-memory tiles (dense main pass 305 ms either
    // way, C3 Quadrotor 4 % slower); CLR_B200_WIDE=1 switches it on for the lean kernel too (experiments)
    const bool wide = global || ctx->wide_mode == 1;
    if (ctx->debug) fprintf(stderr, "[clr] deepz launch: pass %d global %d stats %d full %d wide %d fuse %d grid %d\n",
                            sp.pass, (int)global, (int)stats, (int)full, (int)wide, sp.fuse_tail, grid);
    switch (ksreg) {   // one translation unit per KSREG (deepz_ks*.cu), so that the instantiations compile in parallel
        case 8: e = launch_deepz_ks8(global, stats, full, wide, grid, threads, smem, ctx->stream, net, in, out, ctx->d_state, sp); break;
        case 16: e = launch_deepz_ks16(global, stats, full, wide, grid, threads, smem, ctx->stream, net, in, out, ctx->d_state, sp); break;
        case 25: e = launch_deepz_ks25(global, stats, full, wide, grid, threads, smem, ctx->stream, net, in, out, ctx->d_state, sp); break;
        default: e = launch_deepz_ks32(global, stats, full, wide, grid, threads, smem, ctx->stream, net, in, out, ctx->d_state, sp); break;
    }
    if (e != cudaSuccess) return fail(ctx, CLR_ERR_CUDA, "DeepZ kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    return 0;
}

// Neuron reordering for box-grid calls.  With ReLU layers about half of the neurons are inactive on the small region a
// split covers, and inactive neurons leave all-zero rows, i.e. k-steps of the next layer product that multiply zeros.
// A pilot run over CLR_PILOT_CELLS cells spread over the WHOLE split (so that every shard of the split derives the same
// order) counts in how many cells each neuron stays non-zero; the neurons of every interior layer are then sorted by
// that count (an equivalent reparametrisation of the network: rows of W_l, b_l and columns of W_{l+1}), which moves
// the dead rows to the end where deepz_kernel skips their k-steps (S.kAct).  Results change only by the summation
// order inside the next product (~1e-16 relative).
int reorder_for_grid(clr_ctx* ctx, Compiled* c, const DeepzInput& in_full, long double total_cells, int n_in,
                     const std::vector<unsigned char>& split_key, Compiled** outc) {
    *outc = c;
    const DevNet& D = c->host;
    bool any = false;
    for (int l = 0; l + 1 < D.L; l++) any = any || (D.layer[l].act == CLR_ACT_RELU && D.layer[l].m >= 16);
    if (!any) return 0;
    // the same split as last time (partition and tables): the order is known, no pilot launch, no synchronisation
    if (!split_key.empty() && c->pilot_key == split_key) { if (c->perm) *outc = c->perm; return 0; }
    const long long S = (long long)std::min<long double>((long double)CLR_PILOT_CELLS, total_cells);
    std::vector<long long> ids((size_t)S);
    for (long long k = 0; k < S; k++) ids[(size_t)k] = (long long)((total_cells * k) / S);
    CU(cudaMemcpyAsync(ctx->d_pilot, ids.data(), sizeof(long long) * (size_t)S, cudaMemcpyHostToDevice, ctx->stream));
    { int rcf = bounce_flush(ctx); if (rcf) return rcf; }     // the split tables of a blocking call
    DevCallState hs{};
    hs.n_retry = (unsigned long long)S;
    CU(cudaMemcpyAsync(ctx->d_state2, &hs, sizeof(DevCallState), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_live, 0, sizeof(int) * CLR_DEV_MAX_LAYERS * CLR_MAX_ROWS, ctx->stream));
    DeepzInput in = in_full;
    in.cell_begin = 0;                    // the list holds ids of the whole split
    in.cyc_blk = 0; in.cyc_stride = 0;
    in.N = S;
    DeepzOutput out{};
    out.liveCount = ctx->d_live;
    DeepzSched sp{};
    const size_t smem_cta = ((size_t)ctx->smem_optin + 1024) / DZ_CTAS - 1024;      // shared memory of one CTA
    const size_t meta = dz_geom(D, false, smem_cta, 0, sp, c->ksreg);
    sp.cap = (int)((smem_cta - meta) / 8);
    sp.tile_cells = 4;                    // fixed tiles: which cells a tile drops must not depend on the scheduling,
                                          // or the order (and the last bits of the results) would vary from call to call
    sp.inplace = c->inplace;
    sp.pass = 1;                          // list-driven pass: the sample ids are the "retry list"
    sp.retry_list = ctx->d_pilot;
    sp.big_list = ctx->d_big;             // overflowing cells are simply not sampled (the list is reset by the real run)
    DevCallState* saved = ctx->d_state;
    ctx->d_state = ctx->d_state2;
    int rc = launch_deepz(ctx, false, false, c, (int)std::min<long long>((long long)ctx->sms * DZ_CTAS, S), DZ_THREADS, meta + (size_t)sp.cap * 8,
                          in, out, sp);
    ctx->d_state = saved;
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->h_live, ctx->d_live, sizeof(int) * CLR_DEV_MAX_LAYERS * CLR_MAX_ROWS, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<std::vector<int>> order((size_t)D.L);
    bool changed = false;
    for (int l = 0; l + 1 < D.L; l++) {
        const DevLayer& d = D.layer[l];
        if (d.act != CLR_ACT_RELU || d.m < 16) continue;
        const int* cnt = ctx->h_live + (size_t)l * CLR_MAX_ROWS;
        std::vector<int> o(d.m);
        for (int i = 0; i < d.m; i++) o[i] = i;
        std::stable_sort(o.begin(), o.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
        for (int i = 0; i < d.m; i++) changed = changed || o[i] != i;
        order[(size_t)l] = std::move(o);
    }
    c->pilot_key = split_key;
    if (!changed) { free_compiled(c->perm); c->perm = nullptr; return 0; }
    return compile_net_reordered(ctx, c, n_in, order, outc);
}

// Shared driver of clr_deepz_boxgrid / clr_deepz_zonotopes once `in` points at device inputs.
int run_deepz(clr_ctx* ctx, clr_net* net, Compiled* c, DeepzInput& in, int p0max, double* out_lo, double* out_hi,
              double* hull_lo, double* hull_hi, clr_stats* stats, int keep_cap, int flags) {
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    const bool want_stats = stats != nullptr && !(flags & CLR_FLAG_NO_STATS);
    const DevNet& D = c->host;
    const int m = D.m_out;
    const long long N = in.N;
    if (keep_cap < 0) return fail(ctx, CLR_ERR_ARG, "keep_cap < 0");
    if (want_stats) memset(stats, 0, sizeof(*stats));
    if (N == 0) {
        if (stats) { memset(stats, 0, sizeof(*stats)); }
        if (!dev_io && hull_lo) for (int i = 0; i < m; i++) { hull_lo[i] = INFINITY; hull_hi[i] = -INFINITY; }
        ctx->keep_cells = 0;
        ctx->bounce_open = false;
        return 0;
    }
    // device outputs
    DeepzOutput out{};
    out.keep_cap = keep_cap;
    if (dev_io) {
        out.out_lo = out_lo; out.out_hi = out_hi;
    }
    if (m > ctx->hull_cap) {
        if (ctx->d_hull) CU(cudaFree(ctx->d_hull));
        ctx->d_hull = nullptr;
        CU(cudaMalloc((void**)&ctx->d_hull, sizeof(double) * 2 * m));
        ctx->hull_cap = m;
    }
    const bool want_hull = hull_lo != nullptr && hull_hi != nullptr;
    // Small blocking calls (the reference's own problem sizes: one to a few thousand cells per control step) are bound by
    // the number of driver calls, not by the kernel: the hull seed travels with the inputs, and the hull and the boxes
    // come back in one copy from one contiguous region [hull | lo | hi] of the bounce buffer.
    size_t fast_off = 0, fast_bytes = 0;
    bool fast = false;
    // Narrow networks (every layer at most 32 neurons wide: ACC, VerticalCAS, InvertedPendulum ...) without statistics or kept
    // zonotopes take the one-small-CTA-per-cell kernel (tiny.cuh): the tiled DMMA kernel's phases are a ~40 us latency chain there.
    const bool kstats_early = want_stats || keep_cap > 0;
    const size_t smem_cta_early = ((size_t)ctx->smem_optin + 1024) / DZ_CTAS - 1024;
    bool tiny = !kstats_early && keep_cap == 0 && ctx->tile_cells == 0 && ctx->tiny && D.n_in <= TN_ROWS;
    for (int l = 0; l < D.L && tiny; l++) tiny = D.layer[l].m <= TN_ROWS && D.layer[l].n <= TN_ROWS;
    const int tiny_cols = worst_cols(D, p0max) + 1;
    const size_t tsm = ((size_t)tiny_cols + TN_WARPS) * TN_ROWS * sizeof(double);      // the cell's store + the partial radius sums
    tiny = tiny && tsm <= smem_cta_early;
    bool zc = false;      // zero copy: the kernel reads its inputs from and writes its boxes to the pinned bounce buffer itself
    // The scheduler state travels the same way: its zeroed image goes up with the inputs (no memset call) and comes back with
    // the results (no copy of its own): one region [state | hull | lo | hi] -- four driver calls per step in all.
    const size_t sb = (sizeof(DevCallState) + 255) & ~(size_t)255;
    struct StateSwap { clr_ctx* c; DevCallState* saved; ~StateSwap() { c->d_state = saved; } } swap_guard{ctx, ctx->d_state};
    if (!dev_io && ctx->bounce_open && keep_cap == 0) {
        const size_t hb = want_hull ? sizeof(double) * 2 * (size_t)m : 0;
        const size_t ob = (out_lo && out_hi) ? sizeof(double) * (size_t)m * (size_t)N : 0;
        unsigned char* base = (unsigned char*)bounce_take(ctx, sb + hb + 2 * ob + 16, &fast_off);
        if (base) {
            fast = true;
            fast_bytes = sb + hb + 2 * ob;
            memset(base, 0, sizeof(DevCallState));
            double* hseed = (double*)(base + sb);
            for (int i = 0; want_hull && i < m; i++) { hseed[i] = INFINITY; hseed[m + i] = -INFINITY; }
            ctx->bounce_used = fast_off + sb + hb;     // the state and the seed have to go up
            ctx->d_state = (DevCallState*)(ctx->d_bounce + fast_off);
            double* dbase = (double*)(ctx->d_bounce + fast_off + sb);
            if (want_hull) { out.hull_lo = dbase; out.hull_hi = dbase + m; }
            if (ob) { out.out_lo = dbase + 2 * (want_hull ? m : 0); out.out_hi = out.out_lo + (size_t)m * N; }
        }
    }
    // The reference's own problem sizes on narrow networks (ONE zonotope per control step for ACC): the call is three copies and a
    // 15 us kernel.  The pinned bounce buffer is mapped into the device's address space, so the kernel reads the staged inputs from
    // it and writes the boxes into it over PCIe itself -- no H2D, no D2H: one launch and one synchronisation.  The hull is the
    // min / max of the boxes, taken on the host; the only state the tiny kernel touches is the error word.
    zc = tiny && fast && ctx->zero_copy && ctx->m_bounce && out_lo && out_hi && N <= 8192;
    if (zc) {
        auto remap = [&](const void* p) -> const void* {
            const unsigned char* b = (const unsigned char*)p;
            return (p && b >= ctx->d_bounce && b < ctx->d_bounce + CLR_BOUNCE_BYTES) ? ctx->m_bounce + (b - ctx->d_bounce) : p;
        };
        in.centers = (const double*)remap(in.centers); in.radii = (const double*)remap(in.radii);
        in.C = (const double*)remap(in.C); in.G = (const double*)remap(in.G);
        in.col_off = (const long long*)remap(in.col_off); in.ncols = (const int*)remap(in.ncols);
        out.out_lo = (double*)remap(out.out_lo); out.out_hi = (double*)remap(out.out_hi);
        out.hull_lo = nullptr; out.hull_hi = nullptr;
        ctx->d_state = (DevCallState*)(ctx->m_bounce + fast_off);
        ctx->bounce_flushed = ctx->bounce_used;          // nothing goes up
    }
    { int rcf = bounce_flush(ctx); if (rcf) return rcf; }
    ctx->bounce_open = false;
    if (fast) {
        // nothing else to set up
    } else if (want_hull) {
        fill_hull_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_hull, m);
        CU(cudaGetLastError());
        ctx->launches++;
        out.hull_lo = ctx->d_hull; out.hull_hi = ctx->d_hull + m;
    } else {
        out.hull_lo = nullptr; out.hull_hi = nullptr;
    }
    if (!dev_io && !fast) {
        if (out_lo && out_hi) {
            int rc;
            if ((rc = ensure(ctx, &ctx->d_io[4], &ctx->io_bytes[4], sizeof(double) * (size_t)m * N))) return rc;
            if ((rc = ensure(ctx, &ctx->d_io[5], &ctx->io_bytes[5], sizeof(double) * (size_t)m * N))) return rc;
            out.out_lo = (double*)ctx->d_io[4]; out.out_hi = (double*)ctx->d_io[5];
        } else {
            out.out_lo = nullptr; out.out_hi = nullptr;
        }
    }
    if (keep_cap > 0) {
        int rc;
        if ((rc = ensure(ctx, (void**)&ctx->d_keepN, &ctx->keepN_bytes, sizeof(int) * (size_t)N))) return rc;
        if ((rc = ensure(ctx, (void**)&ctx->d_keepC, &ctx->keepC_bytes, sizeof(double) * (size_t)m * N))) return rc;
        if ((rc = ensure(ctx, (void**)&ctx->d_keepG, &ctx->keepG_bytes, sizeof(double) * (size_t)m * keep_cap * N))) return rc;
        out.keepN = ctx->d_keepN; out.keepC = ctx->d_keepC; out.keepG = ctx->d_keepG;
    }
    ctx->keep_cells = 0;
    // scheduler state and lists
    if (N > ctx->list_cap) {
        if (ctx->d_retry) CU(cudaFree(ctx->d_retry));
        if (ctx->d_big) CU(cudaFree(ctx->d_big));
        ctx->d_retry = ctx->d_big = nullptr; ctx->list_cap = 0;
        CU(cudaMalloc((void**)&ctx->d_retry, sizeof(long long) * (size_t)N));
        CU(cudaMalloc((void**)&ctx->d_big, sizeof(long long) * (size_t)N));
        ctx->list_cap = N;
    }
    if (!fast) CU(cudaMemsetAsync(ctx->d_state, 0, sizeof(DevCallState), ctx->stream));

    const int threads = DZ_THREADS;
    // kept zonotopes need the zero-generator bookkeeping, which only the STATS instantiations carry
    const bool kstats = want_stats || keep_cap > 0;
    const size_t smem_total = ((size_t)ctx->smem_optin + 1024) / DZ_CTAS - 1024;      // shared memory of one CTA
    DeepzSched sp{};
    for (int l = 0; l < D.L; l++)      // every sigmoid / tanh neuron appends a generator (unless its bounds coincide)
        if (D.layer[l].act == CLR_ACT_SIGMOID || D.layer[l].act == CLR_ACT_TANH) sp.grow_cols += D.layer[l].m;
    const size_t meta = dz_geom(D, kstats, smem_total, 0, sp, c->ksreg);
    if (smem_total < meta + 8 * 1024) return fail(ctx, CLR_ERR_CUDA, "device has too little shared memory");
    sp.cap = (int)((smem_total - meta) / 8);
    sp.tile_cells = ctx->tile_cells;
    sp.inplace = c->inplace;
    {
        // the last layer applied column by column in the output stage (deepz_kernel): a final Id layer with at most 4 inputs and
        // rows (the folded post-processing) that the reference does not follow by remove_zero_columns; kept zonotopes need the
        // real columns
        const DevLayer& lt = D.layer[D.L - 1];
        sp.fuse_tail = D.L >= 2 && lt.act == CLR_ACT_ID && lt.m <= 4 && !lt.remove_zero && keep_cap == 0;
    }
    sp.retry_list = ctx->d_retry;
    sp.big_list = ctx->d_big;
    sp.workspace = nullptr;
    int grid = (int)std::min<long long>((long long)ctx->sms * DZ_CTAS, N);
    int rc;
    sp.pass = 0;
    ctx->pass_ms[0] = ctx->pass_ms[1] = ctx->pass_ms[2] = 0.f;
    if (ctx->profiling) CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (tiny) {
        if (tsm > 48 * 1024) CU(cudaFuncSetAttribute(deepz_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
        const long long per_sm = std::max<long long>(1, std::min<long long>(8, (long long)(smem_total / (tsm + 1024))));
        const int tgrid = (int)std::max<long long>(1, std::min<long long>(N, (long long)ctx->sms * per_sm));
        deepz_tiny_kernel<<<tgrid, TN_WARPS * 32, tsm, ctx->stream>>>(c->dev, in, out, ctx->d_state, tiny_cols);
        CU(cudaGetLastError());
        ctx->launches++;
    } else if ((rc = launch_deepz(ctx, false, kstats, c, grid, threads, meta + (size_t)sp.cap * 8, in, out, sp))) return rc;
    if (ctx->profiling) CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    sp.pass = 1;
    bool fast_done = false;
    if (zc) {
        CU(cudaStreamSynchronize(ctx->stream));
        memcpy(ctx->h_state, ctx->h_bounce + fast_off, sizeof(DevCallState));
        fast_done = true;
        if (want_hull) {                                    // min / max of the boxes (what the device-side hull holds)
            double* hb = (double*)(ctx->h_bounce + fast_off + sb);
            const double* lo = hb + 2 * m; const double* hi = lo + (size_t)m * N;
            for (long long j = 0; j < N; j++)
                for (int i = 0; i < m; i++) { hb[i] = fmin(hb[i], lo[j * m + i]); hb[m + i] = fmax(hb[m + i], hi[j * m + i]); }
        }
    } else if (fast) {
        // one round trip: scheduler state and results together; the retry pass only runs when a tile overflowed
        CU(cudaMemcpyAsync(ctx->h_bounce + fast_off, ctx->d_bounce + fast_off, fast_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->profiling) CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        memcpy(ctx->h_state, ctx->h_bounce + fast_off, sizeof(DevCallState));
        fast_done = ctx->h_state->n_retry == 0 && ctx->h_state->n_big == 0;
    }
    if (!zc && (!fast || ctx->h_state->n_retry > 0)) {
        if (!tiny && (rc = launch_deepz(ctx, false, kstats, c, grid, threads, meta + (size_t)sp.cap * 8, in, out, sp))) return rc;
        if (ctx->profiling) CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        // how many cells did not fit in shared memory at all?
        CU(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(DevCallState), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->profiling) {
        CU(cudaEventElapsedTime(&ctx->pass_ms[0], ctx->ev[0], ctx->ev[1]));
        CU(cudaEventElapsedTime(&ctx->pass_ms[1], ctx->ev[1], ctx->ev[2]));
    }
    ctx->last_retry = (int64_t)ctx->h_state->n_retry;
    ctx->last_big = (int64_t)ctx->h_state->n_big;
    if (ctx->h_state->n_big > 0) {
        const int wc = worst_cols(D, p0max);
        DeepzSched sg = sp;
        sg.pass = 2;
        const int stageW = std::max(clr_pad4(D.max_rows), 2 * D.n_in);
        sg.cap = (wc + 8) * D.max_pitch * (c->inplace ? 1 : 2) + stageW + 64;
        int gridg = (int)std::min<unsigned long long>((unsigned long long)ctx->sms * DZ_CTAS, ctx->h_state->n_big);
        // sized for a full grid whatever n_big is: the number of such cells varies from call to call (cells at the edge of the
        // capacity), and growing the workspace means a cudaFree + cudaMalloc in the middle of the step (7 ms per step in the mixed regime)
        size_t need = sizeof(double) * (size_t)ctx->sms * DZ_CTAS * sg.cap;
        if ((rc = ensure(ctx, (void**)&ctx->d_ws, &ctx->ws_bytes, need))) return rc;
        sg.workspace = ctx->d_ws;
        const size_t metag = dz_geom(D, kstats, smem_total, wc + 8 + 40, sg, c->ksreg);     // one wide cell per tile
        if (metag > smem_total) return fail(ctx, CLR_ERR_CAPACITY, "a cell with %d columns exceeds the tile metadata", wc);
        if (ctx->profiling) CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        if ((rc = launch_deepz(ctx, true, kstats, c, gridg, threads, metag, in, out, sg))) return rc;
        if (ctx->profiling) CU(cudaEventRecord(ctx->ev[3], ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(DevCallState), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->profiling) CU(cudaEventElapsedTime(&ctx->pass_ms[2], ctx->ev[2], ctx->ev[3]));
    }
    if (ctx->h_state->error == CLR_ERR_CAPACITY && keep_cap > 0)
        return fail(ctx, CLR_ERR_CAPACITY, "an output zonotope has more than keep_cap = %d generators (max_p_out = %d)", keep_cap,
                    ctx->h_state->max_p_out);
    if (ctx->h_state->error) return fail(ctx, ctx->h_state->error, "device-side error %d in the DeepZ kernel", ctx->h_state->error);
    if (want_stats) {
        const DevCallState& s = *ctx->h_state;
        stats->n_cells = N;
        stats->near_threshold = (int64_t)s.near_threshold;
        stats->max_p_out = s.max_p_out;
        for (int l = 0; l < net->L; l++) {
            stats->sum_p_in[l] = (int64_t)s.sum_p_in[l];
            stats->sum_unstable[l] = (int64_t)s.sum_unstable[l];
            stats->flops += 2.0 * net->dims[l + 1] * net->dims[l] * ((double)s.sum_p_in[l] + (double)N);
        }
    }
    if (keep_cap > 0) { ctx->keep_cells = N; ctx->keep_cap = keep_cap; ctx->keep_m = m; }
    // results
    if (dev_io) {
        if (want_hull) {
            CU(cudaMemcpyAsync(hull_lo, ctx->d_hull, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
            CU(cudaMemcpyAsync(hull_hi, ctx->d_hull + m, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    } else if (fast) {
        if (!fast_done) {
            CU(cudaMemcpyAsync(ctx->h_bounce + fast_off, ctx->d_bounce + fast_off, fast_bytes, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        const double* hb = (const double*)(ctx->h_bounce + fast_off + sb);
        if (want_hull) { memcpy(hull_lo, hb, sizeof(double) * m); memcpy(hull_hi, hb + m, sizeof(double) * m); hb += 2 * m; }
        if (out_lo && out_hi) {
            memcpy(out_lo, hb, sizeof(double) * (size_t)m * N);
            memcpy(out_hi, hb + (size_t)m * N, sizeof(double) * (size_t)m * N);
        }
    } else {
        if (out_lo && out_hi) {
            CU(cudaMemcpyAsync(out_lo, out.out_lo, sizeof(double) * (size_t)m * N, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(out_hi, out.out_hi, sizeof(double) * (size_t)m * N, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (want_hull) {
            CU(cudaMemcpyAsync(hull_lo, ctx->d_hull, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(hull_hi, ctx->d_hull + m, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int stage_in(clr_ctx* ctx, int slot, const void* src, size_t bytes, bool dev_io, const void** out) {
    if (dev_io) { *out = src; return 0; }
    if (ctx->bounce_open && bytes <= CLR_BOUNCE_IN_MAX && ctx->bounce_used + bytes + 256 <= CLR_BOUNCE_IN_MAX) {
        size_t off;
        void* h = bounce_take(ctx, bytes ? bytes : 8, &off);
        if (h) {
            if (bytes) memcpy(h, src, bytes);
            *out = ctx->d_bounce + off;
            return 0;
        }
    }
    int rc = ensure(ctx, &ctx->d_io[slot], &ctx->io_bytes[slot], bytes ? bytes : 8);
    if (rc) return rc;
    if (bytes) CU(cudaMemcpyAsync(ctx->d_io[slot], src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *out = ctx->d_io[slot];
    return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int32_t clr_create(int32_t device, clr_ctx** out) {
    clr_ctx* ctx = nullptr;   // errors before the ctx exists go to the thread-local message
    if (!out) return fail(nullptr, CLR_ERR_ARG, "clr_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, CLR_ERR_CUDA, "no CUDA device available (%s); libclr_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(nullptr, CLR_ERR_ARG, "device %d out of range (0..%d)", device, count - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, CLR_ERR_CUDA, "device %d is sm_%d%d; libclr_b200 is built for sm_100a only", device, prop.major, prop.minor);
    clr_ctx* c = new clr_ctx();
    c->device = device;
    c->sms = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (const char* wv = getenv("CLR_B200_WIDE")) c->wide_mode = atoi(wv) != 0;
    c->debug = getenv("CLR_B200_DEBUG") != nullptr;
    if (const char* v = getenv("CLR_B200_ZEROCOPY")) c->zero_copy = atoi(v) != 0;
    if (const char* v = getenv("CLR_B200_REDUCE_REG")) c->reduce_reg = atoi(v) != 0;
    if (const char* v = getenv("CLR_B200_RO_SPLIT")) c->ro_split = atoi(v) != 0;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_state, sizeof(DevCallState));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_state, sizeof(DevCallState));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_chk, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h_bounce, CLR_BOUNCE_BYTES, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&c->m_bounce, c->h_bounce, 0);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_bounce, CLR_BOUNCE_BYTES);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_state2, sizeof(DevCallState));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_live, sizeof(int) * CLR_DEV_MAX_LAYERS * CLR_MAX_ROWS);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_live, sizeof(int) * CLR_DEV_MAX_LAYERS * CLR_MAX_ROWS);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_pilot, sizeof(long long) * CLR_PILOT_CELLS);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&c->ev[i]);
    if (e != cudaSuccess) {
        int rc = fail(nullptr, CLR_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
        delete c;
        return rc;
    }
    *out = c;
    return 0;
}

void clr_destroy(clr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_state);
    cudaFreeHost(ctx->h_state);
    cudaFree(ctx->d_retry);
    cudaFree(ctx->d_big);
    cudaFree(ctx->d_ws);
    for (int i = 0; i < 8; i++) cudaFree(ctx->d_io[i]);
    cudaFree(ctx->d_hull);
    cudaFree(ctx->d_keepN);
    cudaFree(ctx->d_keepC);
    cudaFree(ctx->d_keepG);
    cudaFree(ctx->d_err);
    cudaFree(ctx->d_chk);
    cudaFreeHost(ctx->h_bounce);
    cudaFree(ctx->d_bounce);
    cudaFree(ctx->d_state2);
    cudaFree(ctx->d_live);
    cudaFreeHost(ctx->h_live);
    cudaFree(ctx->d_pilot);
    for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* clr_last_error(const clr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void* clr_stream(clr_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int32_t clr_synchronize(clr_ctx* ctx) {
    if (!ctx) return CLR_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int64_t clr_launch_count(const clr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t clr_device_alloc(clr_ctx* ctx, int64_t bytes, void** out) {
    if (!ctx || !out || bytes < 0) return fail(ctx, CLR_ERR_ARG, "clr_device_alloc: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(out, (size_t)(bytes ? bytes : 8)));
    return 0;
}
int32_t clr_device_free(clr_ctx* ctx, void* p) {
    if (!ctx) return CLR_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(p));
    return 0;
}
int32_t clr_memcpy_h2d(clr_ctx* ctx, void* dst, const void* src, int64_t bytes) {
    if (!ctx) return CLR_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int32_t clr_memcpy_d2h(clr_ctx* ctx, void* dst, const void* src, int64_t bytes) {
    if (!ctx) return CLR_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int32_t clr_host_register(clr_ctx* ctx, void* p, int64_t bytes) {
    if (!ctx) return CLR_ERR_ARG;
    if (!p || bytes <= 0) return fail(ctx, CLR_ERR_ARG, "clr_host_register: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, CLR_ERR_CUDA, "cudaHostRegister(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e)); }
    return 0;
}
int32_t clr_host_unregister(clr_ctx* ctx, void* p) {
    if (!ctx) return CLR_ERR_ARG;
    if (!p) return 0;
    CU(cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostUnregister(p);
    if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, CLR_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return 0;
}
int32_t clr_last_device_error(clr_ctx* ctx, int32_t* code) {
    if (!ctx || !code) return CLR_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    int err = 0;
    CU(cudaMemcpyAsync(&err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *code = err;
    return 0;
}
int32_t clr_set_tile_cells(clr_ctx* ctx, int32_t cells) {
    if (!ctx || cells < 0 || cells > CLR_TILE_CELLS_MAX) return fail(ctx, CLR_ERR_ARG, "tile cells must be in 0..%d", CLR_TILE_CELLS_MAX);
    ctx->tile_cells = cells;
    return 0;
}

int32_t clr_set_reorder(clr_ctx* ctx, int32_t on) {
    if (!ctx) return CLR_ERR_ARG;
    ctx->reorder = on != 0;
    return 0;
}
int32_t clr_set_tiny(clr_ctx* ctx, int32_t on) {
    if (!ctx) return CLR_ERR_ARG;
    ctx->tiny = on != 0;
    return 0;
}
int32_t clr_set_profiling(clr_ctx* ctx, int32_t on) {
    if (!ctx) return CLR_ERR_ARG;
    ctx->profiling = on != 0;
    return 0;
}
int32_t clr_last_pass_info(clr_ctx* ctx, double* pass_ms, int64_t* retried, int64_t* big) {
    if (!ctx) return CLR_ERR_ARG;
    if (pass_ms) for (int i = 0; i < 3; i++) pass_ms[i] = ctx->pass_ms[i];
    if (retried) *retried = ctx->last_retry;
    if (big) *big = ctx->last_big;
    return 0;
}

int32_t clr_measure_fp64_peak(clr_ctx* ctx, double ms, double* tflops) {
    if (!ctx || !tflops) return CLR_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    const int grid = ctx->sms * 2, threads = 256;
    int rc = ensure(ctx, &ctx->d_io[0], &ctx->io_bytes[0], sizeof(double) * (size_t)grid * threads);
    if (rc) return rc;
    cudaEvent_t e0 = ctx->ev[0], e1 = ctx->ev[1];
    auto run = [&](int iters, float* t) -> cudaError_t {
        cudaError_t e = cudaEventRecord(e0, ctx->stream);
        if (e != cudaSuccess) return e;
        fp64_peak_kernel<<<grid, threads, 0, ctx->stream>>>((double*)ctx->d_io[0], iters, 1.0000001, 0.9999999);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if ((e = cudaEventRecord(e1, ctx->stream)) != cudaSuccess) return e;
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) return e;
        return cudaEventElapsedTime(t, e0, e1);
    };
    float t = 0.f;
    CU(run(2000, &t));                                   // warm-up, and the rate for sizing the timed launch
    CU(run(2000, &t));
    const double per_iter = (double)t / 2000.0;
    const int iters = (int)std::min(5.0e7, std::max(2000.0, (ms > 0 ? ms : 20.0) / std::max(per_iter, 1e-9)));
    CU(run(iters, &t));
    ctx->launches += 3;
    const double flop = 512.0 * 8.0 * (double)iters * (double)grid * (threads / 32);
    *tflops = flop / ((double)t * 1e-3) / 1e12;
    return 0;
}

// DZ_TIMING builds only: cycles per phase of the last DeepZ call, summed over CTAs (not part of the public header)
int32_t clr_debug_timers(clr_ctx* ctx, unsigned long long* out8) {
    if (!ctx || !out8) return CLR_ERR_ARG;
    for (int i = 0; i < 8; i++) out8[i] = ctx->h_state->timers[i];
    return 0;
}

int32_t clr_net_upload(clr_ctx* ctx, int32_t n_layers, const int32_t* dims, const int32_t* act,
                       const double* const* W, const double* const* b, clr_net** out) {
    if (!ctx) return CLR_ERR_ARG;
    if (!out || !dims || !act || !W || !b) return fail(ctx, CLR_ERR_ARG, "clr_net_upload: NULL argument");
    *out = nullptr;
    if (n_layers < 1 || n_layers > CLR_MAX_LAYERS) return fail(ctx, CLR_ERR_ARG, "n_layers must be in 1..%d", CLR_MAX_LAYERS);
    for (int l = 0; l <= n_layers; l++) if (dims[l] < 1) return fail(ctx, CLR_ERR_DIM, "dims[%d] = %d", l, dims[l]);
    for (int l = 0; l < n_layers; l++) {
        if (act[l] < CLR_ACT_ID || act[l] > CLR_ACT_TANH) return fail(ctx, CLR_ERR_ARG, "unknown activation %d in layer %d", act[l], l);
        if (!W[l] || !b[l]) return fail(ctx, CLR_ERR_ARG, "layer %d has NULL weights", l);
        if (dims[l + 1] > CLR_MAX_ROWS) return fail(ctx, CLR_ERR_ARG, "layer %d has %d neurons; at most %d supported", l, dims[l + 1], CLR_MAX_ROWS);
    }
    clr_net* net = new clr_net();
    net->ctx = ctx;
    net->device = ctx->device;
    net->L = n_layers;
    net->dims.assign(dims, dims + n_layers + 1);
    net->act.assign(act, act + n_layers);
    for (int l = 0; l < n_layers; l++) {
        net->W.emplace_back(W[l], W[l] + (size_t)dims[l + 1] * dims[l]);
        net->b.emplace_back(b[l], b[l] + dims[l + 1]);
    }
    *out = net;
    return 0;
}

void clr_net_free(clr_net* net) {
    if (!net) return;
    cudaSetDevice(net->device);
    for (Compiled* c : net->variants) free_compiled(c);
    delete net;
}

int32_t clr_deepz_boxgrid(clr_ctx* ctx, const clr_net* net_c, int32_t n, const int32_t* partition,
                          const double* centers, const double* radii, int64_t cell_begin,
                          int64_t cell_count, const clr_prepost* pp, double* out_lo, double* out_hi,
                          double* hull_lo, double* hull_hi, clr_stats* stats, int32_t keep_cap,
                          int32_t flags) {
    return clr_deepz_boxgrid_cyclic(ctx, net_c, n, partition, centers, radii, cell_begin, cell_count, 0, 0, pp, out_lo, out_hi,
                                    hull_lo, hull_hi, stats, keep_cap, flags);
}

int32_t clr_deepz_boxgrid_cyclic(clr_ctx* ctx, const clr_net* net_c, int32_t n, const int32_t* partition,
                                 const double* centers, const double* radii, int64_t cell_begin, int64_t cell_count,
                                 int64_t block, int64_t stride, const clr_prepost* pp, double* out_lo, double* out_hi,
                                 double* hull_lo, double* hull_hi, clr_stats* stats, int32_t keep_cap, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    clr_net* net = const_cast<clr_net*>(net_c);
    if (!net || !partition || !centers || !radii) return fail(ctx, CLR_ERR_ARG, "clr_deepz_boxgrid: NULL argument");
    if (n < 1 || n > 64) return fail(ctx, CLR_ERR_DIM, "box dimension %d not in 1..64", n);
    CU(cudaSetDevice(ctx->device));
    long double total = 1;
    int tab = 0;
    DeepzInput in{};
    for (int d = 0; d < n; d++) {
        if (partition[d] < 1) return fail(ctx, CLR_ERR_ARG, "partition[%d] = %d: each dimension needs at least one block", d, partition[d]);
        in.part[d] = partition[d]; in.base[d] = tab;
        in.stride[d] = (long long)std::min<long double>(total, 9.0e18L);
        tab += partition[d];
        total *= partition[d];
    }
    if (block < 0 || stride < 0 || (block > 0 && stride < block))
        return fail(ctx, CLR_ERR_ARG, "block-cyclic ids need 0 < block <= stride (got block %lld, stride %lld)", (long long)block, (long long)stride);
    long double last = (long double)cell_begin + (long double)cell_count;        // one past the last id of the call
    if (block > 0 && cell_count > 0)
        last = (long double)cell_begin + (long double)((cell_count - 1) / block) * (long double)stride + (long double)((cell_count - 1) % block) + 1;
    if (cell_begin < 0 || cell_count < 0 || last > total)
        return fail(ctx, CLR_ERR_ARG, "cell range [%lld, %lld) outside the split", (long long)cell_begin, (long long)last);
    Compiled* c = nullptr;
    int rc = compile_net(ctx, net, pp, n, &c);
    if (rc) return rc;
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    ctx->bounce_open = !dev_io; ctx->bounce_used = ctx->bounce_flushed = 0;
    const void *dc, *dr;
    if ((rc = stage_in(ctx, 0, centers, sizeof(double) * tab, dev_io, &dc))) return rc;
    if ((rc = stage_in(ctx, 1, radii, sizeof(double) * tab, dev_io, &dr))) return rc;
    in.kind = 0; in.n = n; in.N = cell_count; in.cell_begin = cell_begin;
    in.cyc_blk = block; in.cyc_stride = stride;
    in.centers = (const double*)dc; in.radii = (const double*)dr;
    if (ctx->reorder && keep_cap == 0 && cell_count >= CLR_REORDER_MIN_CELLS) {
        if (cell_count > ctx->list_cap) {       // the pilot borrows the large-cell list of the real run
            if (ctx->d_retry) CU(cudaFree(ctx->d_retry));
            if (ctx->d_big) CU(cudaFree(ctx->d_big));
            ctx->d_retry = ctx->d_big = nullptr; ctx->list_cap = 0;
            CU(cudaMalloc((void**)&ctx->d_retry, sizeof(long long) * (size_t)cell_count));
            CU(cudaMalloc((void**)&ctx->d_big, sizeof(long long) * (size_t)cell_count));
            ctx->list_cap = cell_count;
        }
        // key of the split: partition + tables (host tables by value; device-resident tables by address -- a stale order is
        // still a valid order, it only skips fewer k-steps)
        std::vector<unsigned char> sk;
        auto put = [&](const void* p, size_t nb) { const unsigned char* q = (const unsigned char*)p; sk.insert(sk.end(), q, q + nb); };
        put(partition, sizeof(int32_t) * n);
        if (dev_io) { put(&centers, sizeof(centers)); put(&radii, sizeof(radii)); }
        else { put(centers, sizeof(double) * tab); put(radii, sizeof(double) * tab); }
        if ((rc = reorder_for_grid(ctx, c, in, total, n, sk, &c))) return rc;
    }
    return run_deepz(ctx, net, c, in, n, out_lo, out_hi, hull_lo, hull_hi, stats, keep_cap, flags);
}

int32_t clr_deepz_zonotopes(clr_ctx* ctx, const clr_net* net_c, int32_t n, int64_t N, const double* C,
                            const int64_t* col_off, const double* G, const clr_prepost* pp,
                            double* out_lo, double* out_hi, double* hull_lo, double* hull_hi,
                            clr_stats* stats, int32_t keep_cap, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    clr_net* net = const_cast<clr_net*>(net_c);
    if (!net || N < 0 || n < 1) return fail(ctx, CLR_ERR_ARG, "clr_deepz_zonotopes: bad argument");
    if (N > 0 && (!C || !col_off)) return fail(ctx, CLR_ERR_ARG, "clr_deepz_zonotopes: NULL centres / offsets");
    CU(cudaSetDevice(ctx->device));
    Compiled* c = nullptr;
    int rc = compile_net(ctx, net, pp, n, &c);
    if (rc) return rc;
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    DeepzInput in{};
    in.kind = 1; in.n = n; in.N = N;
    int p0max = 0;
    ctx->bounce_open = !dev_io; ctx->bounce_used = ctx->bounce_flushed = 0;
    if (N > 0) {
        const void *dC, *dO, *dG;
        if (dev_io) {
            int mm[2] = {0, 0x7fffffff};
            CU(cudaMemcpyAsync(ctx->d_chk, mm, sizeof(mm), cudaMemcpyHostToDevice, ctx->stream));
            max_cols_kernel<<<std::min<long long>(1024, (N + 255) / 256), 256, 0, ctx->stream>>>((const long long*)col_off, nullptr, N, ctx->d_chk);
            CU(cudaGetLastError());
            ctx->launches++;
            CU(cudaMemcpyAsync(mm, ctx->d_chk, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (mm[1] < 0) return fail(ctx, CLR_ERR_ARG, "col_off must start at 0 and be non-decreasing");
            if (mm[0] > CLR_NCOL_MAX - 8) return fail(ctx, CLR_ERR_ARG, "a cell has %d generators; at most %d supported", mm[0], CLR_NCOL_MAX - 8);
            p0max = mm[0];
            dC = C; dO = col_off; dG = G;
        } else {
            if (col_off[0] != 0) return fail(ctx, CLR_ERR_ARG, "col_off[0] must be 0");
            for (int64_t i = 0; i < N; i++) {
                int64_t p = col_off[i + 1] - col_off[i];
                if (p < 0) return fail(ctx, CLR_ERR_ARG, "col_off must be non-decreasing (cell %lld)", (long long)i);
                if (p > CLR_NCOL_MAX - 8) return fail(ctx, CLR_ERR_ARG, "cell %lld has %lld generators; at most %d supported", (long long)i, (long long)p, CLR_NCOL_MAX - 8);
                p0max = std::max<int>(p0max, (int)p);
            }
            if (col_off[N] > 0 && !G) return fail(ctx, CLR_ERR_ARG, "clr_deepz_zonotopes: NULL generators");
            if ((rc = stage_in(ctx, 0, C, sizeof(double) * (size_t)n * N, false, &dC))) return rc;
            if ((rc = stage_in(ctx, 1, col_off, sizeof(int64_t) * (size_t)(N + 1), false, &dO))) return rc;
            if ((rc = stage_in(ctx, 2, G, sizeof(double) * (size_t)n * col_off[N], false, &dG))) return rc;
        }
        in.C = (const double*)dC; in.col_off = (const long long*)dO; in.G = (const double*)dG;
    }
    return run_deepz(ctx, net, c, in, p0max, out_lo, out_hi, hull_lo, hull_hi, stats, keep_cap, flags);
}

int32_t clr_deepz_zonotopes_padded(clr_ctx* ctx, const clr_net* net_c, int32_t n, int64_t N, const double* C, const double* G,
                                   const int32_t* ncols, int32_t cap, const clr_prepost* pp, double* out_lo, double* out_hi,
                                   double* hull_lo, double* hull_hi, clr_stats* stats, int32_t keep_cap, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    clr_net* net = const_cast<clr_net*>(net_c);
    if (!net || N < 0 || n < 1 || cap < 0) return fail(ctx, CLR_ERR_ARG, "clr_deepz_zonotopes_padded: bad argument");
    if (N > 0 && (!C || !ncols || (cap > 0 && !G))) return fail(ctx, CLR_ERR_ARG, "clr_deepz_zonotopes_padded: NULL argument");
    if (cap > CLR_NCOL_MAX - 8) return fail(ctx, CLR_ERR_ARG, "cap = %d generators per cell; at most %d supported", cap, CLR_NCOL_MAX - 8);
    CU(cudaSetDevice(ctx->device));
    Compiled* c = nullptr;
    int rc = compile_net(ctx, net, pp, n, &c);
    if (rc) return rc;
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    DeepzInput in{};
    in.kind = 1; in.n = n; in.N = N; in.pad_cap = cap;
    ctx->bounce_open = !dev_io; ctx->bounce_used = ctx->bounce_flushed = 0;
    if (N > 0) {
        const void *dC, *dG, *dN;
        if (!dev_io) {
            for (int64_t i = 0; i < N; i++)
                if (ncols[i] < 0 || ncols[i] > cap) return fail(ctx, CLR_ERR_ARG, "ncols[%lld] = %d outside 0..cap = %d", (long long)i, ncols[i], cap);
        } else {
            int mm[2] = {0, 0x7fffffff};
            CU(cudaMemcpyAsync(ctx->d_chk, mm, sizeof(mm), cudaMemcpyHostToDevice, ctx->stream));
            max_cols_kernel<<<std::min<long long>(1024, (N + 255) / 256), 256, 0, ctx->stream>>>(nullptr, (const int*)ncols, N, ctx->d_chk);
            CU(cudaGetLastError());
            ctx->launches++;
            CU(cudaMemcpyAsync(mm, ctx->d_chk, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (mm[1] < 0 || mm[0] > cap) return fail(ctx, CLR_ERR_ARG, "ncols outside 0..cap = %d (min %d, max %d)", cap, mm[1], mm[0]);
        }
        if ((rc = stage_in(ctx, 0, C, sizeof(double) * (size_t)n * N, dev_io, &dC))) return rc;
        if ((rc = stage_in(ctx, 1, ncols, sizeof(int32_t) * (size_t)N, dev_io, &dN))) return rc;
        if ((rc = stage_in(ctx, 2, G, sizeof(double) * (size_t)n * cap * N, dev_io, &dG))) return rc;
        in.C = (const double*)dC; in.ncols = (const int*)dN; in.G = (const double*)dG; in.col_off = nullptr;
    }
    return run_deepz(ctx, net, c, in, cap, out_lo, out_hi, hull_lo, hull_hi, stats, keep_cap, flags);
}

int32_t clr_deepz_fetch_zonotopes(clr_ctx* ctx, int32_t* out_ncols, double* out_c, double* out_G) {
    if (!ctx) return CLR_ERR_ARG;
    if (ctx->keep_cells <= 0) return fail(ctx, CLR_ERR_STATE, "no kept zonotopes: run a DeepZ call with keep_cap > 0 first");
    if (!out_ncols || !out_c || !out_G) return fail(ctx, CLR_ERR_ARG, "clr_deepz_fetch_zonotopes: NULL argument");
    CU(cudaSetDevice(ctx->device));
    const size_t N = (size_t)ctx->keep_cells, m = (size_t)ctx->keep_m, cap = (size_t)ctx->keep_cap;
    CU(cudaMemcpyAsync(out_ncols, ctx->d_keepN, sizeof(int) * N, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_c, ctx->d_keepC, sizeof(double) * m * N, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_G, ctx->d_keepG, sizeof(double) * m * cap * N, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // generators past ncols are unspecified on the device: zero them for the caller
    for (size_t i = 0; i < N; i++) {
        size_t k = (size_t)out_ncols[i];
        if (k < cap) memset(out_G + (i * cap + k) * m, 0, sizeof(double) * (cap - k) * m);
    }
    return 0;
}

int32_t clr_deepz_fetch_reduced(clr_ctx* ctx, int32_t order, double* out_c, double* out_G, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    if (ctx->keep_cells <= 0) return fail(ctx, CLR_ERR_STATE, "no kept zonotopes: run a DeepZ call with keep_cap > 0 first");
    if (!out_c || !out_G) return fail(ctx, CLR_ERR_ARG, "clr_deepz_fetch_reduced: NULL argument");
    if (order < 1) return fail(ctx, CLR_ERR_ARG, "the target order should be at least 1, but it is %d", order);
    CU(cudaSetDevice(ctx->device));
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    const size_t N = (size_t)ctx->keep_cells, m = (size_t)ctx->keep_m;
    const size_t bc = sizeof(double) * m * N, bg = sizeof(double) * m * m * (size_t)order * N;
    ReduceArgs A;
    A.N = (long long)N; A.m = (int)m; A.cap = ctx->keep_cap; A.order = order;
    A.keepN = ctx->d_keepN; A.keepC = ctx->d_keepC; A.keepG = ctx->d_keepG; A.error = ctx->d_err; A.cap2 = 0;
    if (dev_io) { A.outC = out_c; A.outG = out_G; }
    else {
        int rc;
        if ((rc = ensure(ctx, &ctx->d_io[6], &ctx->io_bytes[6], bc))) return rc;
        if ((rc = ensure(ctx, &ctx->d_io[7], &ctx->io_bytes[7], bg))) return rc;
        A.outC = (double*)ctx->d_io[6]; A.outG = (double*)ctx->d_io[7];
    }
    int cap2 = 1;                       // the index array is sorted over the next power of two
    while (cap2 < A.cap) cap2 <<= 1;
    A.cap2 = cap2;
    CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    const long long want = ((long long)N + RD_WARPS - 1) / RD_WARPS;
    // zonotopes of at most 256 generators whose staged copy fits: generators staged in shared memory, sort in registers
    const int E = std::max(1, cap2 / 32);
    const size_t smem_reg = (size_t)RD_WARPS * (2 * (size_t)A.cap * m * sizeof(double) + (size_t)32 * E * sizeof(int));
    if (E <= 8 && smem_reg <= (size_t)ctx->smem_optin / 2 && ctx->reduce_reg) {
        void (*kern)(ReduceArgs) = E == 1 ? reduce_order_reg_kernel<1> : E == 2 ? reduce_order_reg_kernel<2>
                                 : E == 4 ? reduce_order_reg_kernel<4> : reduce_order_reg_kernel<8>;
        if (smem_reg > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_reg));
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)ctx->smem_optin / (smem_reg + 1024)));
        const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)ctx->sms * per_sm));
        kern<<<grid, RD_WARPS * 32, smem_reg, ctx->stream>>>(A);
    } else {
        const size_t smem = (size_t)RD_WARPS * ((size_t)A.cap * sizeof(double) + (size_t)cap2 * sizeof(int));
        if (smem > (size_t)ctx->smem_optin) return fail(ctx, CLR_ERR_CAPACITY, "keep_cap = %d is too large for the order reduction", A.cap);
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(reduce_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)ctx->smem_optin / (smem + 1024)));
        const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)ctx->sms * per_sm));
        reduce_order_kernel<<<grid, RD_WARPS * 32, smem, ctx->stream>>>(A);
    }
    CU(cudaGetLastError());
    ctx->launches++;
    int err = 0;
    CU(cudaMemcpyAsync(&err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (!dev_io) {
        CU(cudaMemcpyAsync(out_c, A.outC, bc, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(out_G, A.outG, bg, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (err)
        return fail(ctx, CLR_ERR_DIM, "order-%d reduction needs at least %d generators per zonotope (the reference's reduce_order "
                    "indexes past the end there)", order, (int)m * (order - 1));
    return 0;
}

int32_t clr_split_zonotopes(clr_ctx* ctx, int32_t n, int64_t N, const double* C, const double* G, const int32_t* ncols, int32_t cap,
                            int32_t n_gens, const int32_t* gens, const int32_t* nparts, double* out_C, double* out_G,
                            int32_t* out_ncols, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    if (n < 1 || N < 0 || cap < 0 || n_gens < 0 || (n_gens > 0 && (!gens || !nparts)))
        return fail(ctx, CLR_ERR_ARG, "clr_split_zonotopes: bad argument");
    int H = 0;
    for (int i = 0; i < n_gens; i++) {
        if (nparts[i] < 0) return fail(ctx, CLR_ERR_ARG, "nparts[%d] = %d", i, nparts[i]);
        if (gens[i] < 0 || gens[i] >= cap) return fail(ctx, CLR_ERR_ARG, "generator index %d outside the capacity %d", gens[i], cap);
        H += nparts[i];
    }
    if (H > 30) return fail(ctx, CLR_ERR_ARG, "%d halvings: 2^%d children per zonotope are not supported", H, H);
    if (N == 0) return 0;
    if (!C || !ncols || (cap > 0 && !G) || !out_C || !out_G || !out_ncols) return fail(ctx, CLR_ERR_ARG, "clr_split_zonotopes: NULL buffer");
    CU(cudaSetDevice(ctx->device));
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    const size_t total = (size_t)N << H;
    ctx->bounce_open = false;
    int rc;
    const void *dC, *dG, *dN, *dmeta;
    std::vector<int> meta(2 * (size_t)std::max(n_gens, 1));
    for (int i = 0; i < n_gens; i++) { meta[i] = gens[i]; meta[n_gens + i] = nparts[i]; }
    if ((rc = stage_in(ctx, 0, C, sizeof(double) * (size_t)n * N, dev_io, &dC))) return rc;
    if ((rc = stage_in(ctx, 1, ncols, sizeof(int32_t) * (size_t)N, dev_io, &dN))) return rc;
    if ((rc = stage_in(ctx, 2, G, sizeof(double) * (size_t)n * cap * N, dev_io, &dG))) return rc;
    if ((rc = stage_in(ctx, 3, meta.data(), sizeof(int) * meta.size(), false, &dmeta))) return rc;
    SplitArgs A;
    A.N = N; A.n = n; A.cap = cap; A.H = H; A.ng = n_gens;
    A.C = (const double*)dC; A.G = (const double*)dG; A.ncols = (const int*)dN;
    A.gens = (const int*)dmeta; A.nparts = (const int*)dmeta + n_gens; A.error = ctx->d_err;
    const size_t bC = sizeof(double) * (size_t)n * total, bG = sizeof(double) * (size_t)n * cap * total, bN = sizeof(int) * total;
    if (dev_io) { A.outC = out_C; A.outG = out_G; A.outN = out_ncols; }
    else {
        if ((rc = ensure(ctx, &ctx->d_io[5], &ctx->io_bytes[5], bC))) return rc;
        if ((rc = ensure(ctx, &ctx->d_io[6], &ctx->io_bytes[6], bG ? bG : 8))) return rc;
        if ((rc = ensure(ctx, &ctx->d_io[7], &ctx->io_bytes[7], bN))) return rc;
        A.outC = (double*)ctx->d_io[5]; A.outG = (double*)ctx->d_io[6]; A.outN = (int*)ctx->d_io[7];
    }
    CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    const size_t threads = total * 32;
    split_zonotopes_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(A);
    CU(cudaGetLastError());
    ctx->launches++;
    if (!dev_io) {
        int err = 0;
        CU(cudaMemcpyAsync(&err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(out_C, A.outC, bC, cudaMemcpyDeviceToHost, ctx->stream));
        if (bG) CU(cudaMemcpyAsync(out_G, A.outG, bG, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(out_ncols, A.outN, bN, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (err) return fail(ctx, CLR_ERR_ARG, "a listed generator index is not smaller than the generator count of its zonotope");
    }
    return 0;
}

int32_t clr_sign_split(clr_ctx* ctx, int64_t N, const double* lo, const double* hi, double* out_lo, double* out_hi,
                       int64_t* out_parent, int64_t* out_count, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    if (N < 0 || !out_count) return fail(ctx, CLR_ERR_ARG, "clr_sign_split: bad argument");
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    if (N == 0) { if (!dev_io) *out_count = 0; return 0; }
    if (!lo || !hi || !out_lo || !out_hi || !out_parent) return fail(ctx, CLR_ERR_ARG, "clr_sign_split: NULL buffer");
    CU(cudaSetDevice(ctx->device));
    ctx->bounce_open = false;
    int rc;
    const void *dlo, *dhi;
    if ((rc = stage_in(ctx, 0, lo, sizeof(double) * (size_t)N, dev_io, &dlo))) return rc;
    if ((rc = stage_in(ctx, 1, hi, sizeof(double) * (size_t)N, dev_io, &dhi))) return rc;
    const long long nb = (N + SS_BLOCK - 1) / SS_BLOCK;
    if ((rc = ensure(ctx, &ctx->d_io[3], &ctx->io_bytes[3], sizeof(long long) * (size_t)(nb + 1)))) return rc;
    long long* bs = (long long*)ctx->d_io[3];
    double *olo = out_lo, *ohi = out_hi;
    long long* opar = (long long*)out_parent;
    long long* ocnt = dev_io ? (long long*)out_count : bs + nb;
    if (!dev_io) {
        if ((rc = ensure(ctx, &ctx->d_io[5], &ctx->io_bytes[5], sizeof(double) * 2 * (size_t)N))) return rc;
        if ((rc = ensure(ctx, &ctx->d_io[6], &ctx->io_bytes[6], sizeof(double) * 2 * (size_t)N))) return rc;
        if ((rc = ensure(ctx, &ctx->d_io[7], &ctx->io_bytes[7], sizeof(long long) * 2 * (size_t)N))) return rc;
        olo = (double*)ctx->d_io[5]; ohi = (double*)ctx->d_io[6]; opar = (long long*)ctx->d_io[7];
    }
    sign_count_kernel<<<(unsigned)nb, SS_BLOCK, 0, ctx->stream>>>(N, (const double*)dlo, (const double*)dhi, bs);
    sign_scan_kernel<<<1, SS_BLOCK, 0, ctx->stream>>>(nb, bs, ocnt);
    sign_scatter_kernel<<<(unsigned)nb, SS_BLOCK, 0, ctx->stream>>>(N, (const double*)dlo, (const double*)dhi, bs, olo, ohi, opar);
    CU(cudaGetLastError());
    ctx->launches += 3;
    if (!dev_io) {
        long long cnt = 0;
        CU(cudaMemcpyAsync(&cnt, ocnt, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemcpyAsync(out_lo, olo, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(out_hi, ohi, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(out_parent, opar, sizeof(long long) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        *out_count = cnt;
    }
    return 0;
}

int32_t clr_project_oa(clr_ctx* ctx, int32_t n, int32_t order_t, int32_t n_monomials, const int32_t* exponents, int64_t N,
                       const double* coeffs, const double* rem, const double* dt, const int32_t* vars, int32_t n_keep,
                       int32_t remove_zero_generators, double* out_c, double* out_G, int32_t* out_ncols, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    if (n < 1 || n > 64 || order_t < 0 || n_monomials < 1 || N < 0 || n_keep < 1 || n_keep > n)
        return fail(ctx, CLR_ERR_DIM, "clr_project_oa: bad dimensions (n = %d, order_t = %d, monomials = %d, kept = %d)", n, order_t,
                    n_monomials, n_keep);
    if (!exponents || !vars || (N > 0 && (!coeffs || !rem || !dt || !out_c || !out_G || !out_ncols)))
        return fail(ctx, CLR_ERR_ARG, "clr_project_oa: NULL argument");
    for (int k = 0; k < n_keep; k++)
        if (vars[k] < 0 || vars[k] >= n) return fail(ctx, CLR_ERR_DIM, "clr_project_oa: variable %d outside 0..%d", vars[k], n - 1);
    if (N == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    ctx->bounce_open = false;
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    // classify the monomials once: constant / linear in variable j / nonlinear (all exponents even or not)
    const int M = n_monomials;
    std::vector<int> meta((size_t)M + n_keep);
    int n_const = 0;
    for (int m = 0; m < M; m++) {
        int deg = 0, lin = -1; bool even = true;
        for (int j = 0; j < n; j++) {
            const int e = exponents[(size_t)m * n + j];
            if (e < 0) return fail(ctx, CLR_ERR_ARG, "clr_project_oa: negative exponent");
            deg += e; if (e == 1) lin = j; if (e & 1) even = false;
        }
        meta[(size_t)m] = deg == 0 ? -1 : (deg == 1 ? lin : (even ? -2 : -3));
        n_const += deg == 0;
    }
    if (n_const != 1) return fail(ctx, CLR_ERR_ARG, "clr_project_oa: the exponent table needs exactly one constant monomial (has %d)", n_const);
    for (int k = 0; k < n_keep; k++) meta[(size_t)M + k] = vars[k];
    // LazySets' remove_redundant_generators (CLR_FLAG_MERGE_REDUNDANT) needs every component: parallel columns are found in n dimensions
    const bool merge = (flags & CLR_FLAG_MERGE_REDUNDANT) != 0;
    if (merge && n > PJ_NMAX) return fail(ctx, CLR_ERR_DIM, "CLR_FLAG_MERGE_REDUNDANT supports at most %d variables", PJ_NMAX);
    const int n_eval = merge ? n : n_keep;
    if (merge) { meta.resize((size_t)M + n_keep + n); for (int v = 0; v < n; v++) meta[(size_t)M + n_keep + v] = v; }
    const size_t bcoef = sizeof(double) * (size_t)N * n * (order_t + 1) * M, brem = sizeof(double) * (size_t)N * n * 2;
    const size_t bC = sizeof(double) * (size_t)N * n_keep, bG = sizeof(double) * (size_t)N * 2 * n * n_keep;
    const size_t bscr = sizeof(double) * (size_t)N * (size_t)((flags & CLR_FLAG_MERGE_REDUNDANT) ? n : n_keep) * (n + 2);
    int rc;
    const void *dcoef, *drem, *ddt, *dmeta;
    if ((rc = stage_in(ctx, 0, coeffs, bcoef, dev_io, &dcoef))) return rc;
    if ((rc = stage_in(ctx, 1, rem, brem, dev_io, &drem))) return rc;
    if ((rc = stage_in(ctx, 2, dt, sizeof(double) * (size_t)N, dev_io, &ddt))) return rc;
    if ((rc = stage_in(ctx, 3, meta.data(), sizeof(int) * meta.size(), false, &dmeta))) return rc;
    ProjArgs A;
    A.N = (long long)N; A.n = n; A.KT = order_t + 1; A.M = M; A.nkeep = n_eval; A.remove_zero = remove_zero_generators != 0;
    A.cls = (const int*)dmeta; A.vars = merge ? (const int*)dmeta + M + n_keep : (const int*)dmeta + M;
    A.merge = merge; A.nproj = n_keep; A.pvars = (const int*)dmeta + M;
    A.coeffs = (const double*)dcoef; A.rem = (const double*)drem; A.dt = (const double*)ddt;
    if ((rc = ensure(ctx, (void**)&ctx->d_ws, &ctx->ws_bytes, bscr))) return rc;
    A.scratch = ctx->d_ws;
    CU(cudaMemsetAsync(A.scratch, 0, bscr, ctx->stream));
    if (dev_io) { A.outC = out_c; A.outG = out_G; A.outN = out_ncols; }
    else {
        if ((rc = ensure(ctx, &ctx->d_io[5], &ctx->io_bytes[5], bC))) return rc;
        if ((rc = ensure(ctx, &ctx->d_io[6], &ctx->io_bytes[6], bG))) return rc;
        if ((rc = ensure(ctx, &ctx->d_io[7], &ctx->io_bytes[7], sizeof(int) * (size_t)N))) return rc;
        A.outC = (double*)ctx->d_io[5]; A.outG = (double*)ctx->d_io[6]; A.outN = (int*)ctx->d_io[7];
    }
    const long long units = (long long)N * n_eval;
    const int grid1 = (int)std::max<long long>(1, std::min<long long>((units + PJ_WARPS - 1) / PJ_WARPS, (long long)ctx->sms * 16));
    project_eval_kernel<<<grid1, PJ_WARPS * 32, 0, ctx->stream>>>(A);
    CU(cudaGetLastError());
    if (merge) project_pack_merge_kernel<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(A);
    else project_pack_kernel<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(A);
    CU(cudaGetLastError());
    ctx->launches += 2;
    if (!dev_io) {
        CU(cudaMemcpyAsync(out_c, A.outC, bC, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(out_G, A.outG, bG, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(out_ncols, A.outN, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

}  // extern "C"

namespace {
template <int CODE>
int launch_points(clr_ctx* ctx, const PointNet& P, int nx, int64_t N, const double* X, double* U) {
    size_t wbytes = sizeof(double) * (size_t)(CODE > 2 ? P.stotal : P.total_doubles);
    int in_smem = wbytes <= (size_t)ctx->smem_optin - 1024;
    size_t smem = in_smem ? wbytes : 0;
    auto kern = forward_points_kernel<CODE>;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = in_smem ? std::max(1, std::min(8, (int)((size_t)ctx->smem_optin / (smem + 1024)))) : 8;
    int grid = (int)std::min<int64_t>((N + 127) / 128, (int64_t)ctx->sms * per_sm);
    kern<<<grid, 128, smem, ctx->stream>>>(P, nx, (long long)N, X, U, in_smem);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

int launch_points_code(clr_ctx* ctx, int code, const PointNet& P, int nx, int64_t N, const double* X, double* U) {
    switch (code) {
        case 66: return launch_points<66>(ctx, P, nx, N, X, U);
        case 72: return launch_points<72>(ctx, P, nx, N, X, U);
        case 136: return launch_points<136>(ctx, P, nx, N, X, U);
        case 1: return launch_points<1>(ctx, P, nx, N, X, U);
        default: return launch_points<2>(ctx, P, nx, N, X, U);
    }
}

template <int PLANT, int CODE>
int launch_rollout_t(clr_ctx* ctx, const PointNet& P, const RolloutArgs& A) {
    size_t wbytes = sizeof(double) * (size_t)(CODE > 2 ? P.stotal : P.total_doubles);
    int in_smem = wbytes <= (size_t)ctx->smem_optin - 1024;
    size_t smem = in_smem ? wbytes : 0;
    auto kern = rollout_kernel<PLANT, CODE>;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = in_smem ? std::max(1, std::min(4, (int)((size_t)ctx->smem_optin / (smem + 1024)))) : 4;
    int grid = (int)std::min<int64_t>((A.N + 127) / 128, (int64_t)ctx->sms * per_sm);
    kern<<<grid, 128, smem, ctx->stream>>>(P, A, in_smem);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

// narrow networks whose packed image fits beside the state slots: RO_R trajectories per thread
template <int PLANT, int CODE>
int launch_rollout_blocked(clr_ctx* ctx, const PointNet& P, const RolloutArgs& A) {
    constexpr int NS = PlantDims<PLANT>::NS, NC = PlantDims<PLANT>::NC;
    const size_t slots = sizeof(double) * RO_R * (NS + NC) * 128;
    const size_t wbytes = sizeof(double) * (size_t)P.stotal;
    if (slots + wbytes > (size_t)ctx->smem_optin / RO_MINBLK - 1024) return launch_rollout_t<PLANT, CODE>(ctx, P, A);
    const size_t smem = slots + wbytes;
    auto kern = rollout_blocked_kernel<PLANT, CODE>;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t span = 128 * RO_R;
    int grid = (int)std::min<int64_t>((A.N + span - 1) / span, (int64_t)ctx->sms * RO_MINBLK);
    kern<<<grid, 128, smem, ctx->stream>>>(P, A, 1);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

// ---- split rollouts (rollout.cuh: controller_cw_kernel + rollout_ode_kernel) -------------------------------------------
// The constant-bank copy of a packed controller (ro_cw) is one per device: a call that needs another image replaces it
// behind the last kernel that read the old one (events order the streams of different contexts on one device).
struct ConstBank {
    std::mutex mu;
    long long serial = 0;            // PointNet::sserial of the resident image, 0 = none
    cudaEvent_t ready = nullptr;     // recorded behind the copy
    cudaEvent_t last_use = nullptr;  // recorded behind the most recent kernel that reads the bank
};
#define CLR_CBANK_DEVICES 64
ConstBank g_cbank[CLR_CBANK_DEVICES];

// with B.mu held: make P's packed image the resident one, ordered on ctx->stream
int cbank_acquire(clr_ctx* ctx, ConstBank& B, const PointNet& P) {
    if (!B.ready) {
        CU(cudaEventCreateWithFlags(&B.ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&B.last_use, cudaEventDisableTiming));
    }
    if (B.serial != P.sserial) {
        if (B.serial) CU(cudaStreamWaitEvent(ctx->stream, B.last_use, 0));
        B.serial = 0;
        CU(cudaMemcpyToSymbolAsync(ro_cw, P.simage, sizeof(double) * (size_t)P.stotal, 0, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaEventRecord(B.ready, ctx->stream));
        B.serial = P.sserial;
    } else {
        CU(cudaStreamWaitEvent(ctx->stream, B.ready, 0));
    }
    return 0;
}

template <int NI, int NO>
int launch_ctrl_cw_n(clr_ctx* ctx, const PointNet& P, int nx, int64_t N, const double* X, double* U) {
    // 4 points per thread, 4 blocks per SM, records fetched one neuron ahead: the best of the variants measured on B200
    // (2 / 6 / 8 points per thread, 5 / 6 blocks per SM with their spills: profiles/r02/rollout_split_variants_b200.txt)
    constexpr int R = 4;
    const int64_t span = 128 * R;
    controller_cw_kernel<NI, NO, R, 4, true><<<(unsigned)((N + span - 1) / span), 128, 0, ctx->stream>>>(P, nx, (long long)N, X, U);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

int launch_ctrl_cw(clr_ctx* ctx, int code, const PointNet& P, int nx, int64_t N, const double* X, double* U) {
    switch (code) {
        case 66: return launch_ctrl_cw_n<4, 2>(ctx, P, nx, N, X, U);
        case 72: return launch_ctrl_cw_n<4, 8>(ctx, P, nx, N, X, U);
        default: return launch_ctrl_cw_n<8, 8>(ctx, P, nx, N, X, U);
    }
}

bool split_eligible(const clr_ctx* ctx, int code, const PointNet& P, int64_t N) {
    return ctx->ro_split && code > 2 && P.L <= RO_CW_LAYERS && P.stotal + 32 <= RO_CW_DOUBLES && ctx->device < CLR_CBANK_DEVICES && N <= (int64_t)1 << 38;
}

// `ode`: launches one control period of the plant (built-in template or the run-time compiled kernel)
template <class OdeLaunch>
int rollout_split(clr_ctx* ctx, int code, const PointNet& P, const RolloutArgs& A, int ns, int nc, OdeLaunch ode) {
    ConstBank& B = g_cbank[ctx->device];
    std::lock_guard<std::mutex> lock(B.mu);
    int rc = cbank_acquire(ctx, B, P);
    if (rc) return rc;
    double t = A.t0;
    for (int it = 0; it < A.iters; it++) {
        const double* xprev = it == 0 ? A.x0 : A.X_end + (size_t)(it - 1) * A.N * ns;
        if ((rc = launch_ctrl_cw(ctx, code, P, ns, A.N, xprev, A.U + (size_t)it * A.N * nc))) return rc;
        const double t1 = it < A.iters - 1 ? t + A.tau : A.T;
        if ((rc = ode(it, t, t1, xprev))) return rc;
        t = t1;
    }
    CU(cudaEventRecord(B.last_use, ctx->stream));
    return 0;
}

template <int PLANT>
int launch_rollout_split(clr_ctx* ctx, int code, const PointNet& P, const RolloutArgs& A) {
    constexpr int NS = PlantDims<PLANT>::NS, NC = PlantDims<PLANT>::NC;
    const unsigned grid = (unsigned)((A.N + 127) / 128);
    return rollout_split(ctx, code, P, A, NS, NC, [&](int it, double t, double t1, const double* xprev) -> int {
        rollout_ode_kernel<PLANT><<<grid, 128, 0, ctx->stream>>>(A, it, t, t1, xprev);
        CU(cudaGetLastError());
        ctx->launches++;
        return 0;
    });
}

// run-time compiled plant: same kernels, launched through the cudaLibrary handle of the plant
int plant_build(clr_ctx* ctx, clr_plant* p, int slot, bool load);

int launch_rollout_user(clr_ctx* ctx, const clr_plant* plant, int code, const PointNet& P, const RolloutArgs& A) {
    if (split_eligible(ctx, code, P, A.N)) {
        int brc = plant_build(ctx, const_cast<clr_plant*>(plant), 5, true);
        if (brc) return brc;
        void* fn = (void*)plant->kern[5];
        const unsigned grid = (unsigned)((A.N + 127) / 128);
        return rollout_split(ctx, code, P, A, plant->ns, plant->nc, [&](int it, double t, double t1, const double* xprev) -> int {
            RolloutArgs Ac = A;
            void* args[5] = {&Ac, &it, &t, &t1, &xprev};
            CU(cudaLaunchKernel(fn, dim3(grid), dim3(128), args, 0, ctx->stream));
            ctx->launches++;
            return 0;
        });
    }
    int slot = 4;
    for (int i = 0; i < 5; i++) if (kPlantCodes[i] == code) slot = i;
    int brc = plant_build(ctx, const_cast<clr_plant*>(plant), slot, true);
    if (brc) return brc;
    const bool narrow = code > 2;
    const size_t wbytes = sizeof(double) * (size_t)(narrow ? P.stotal : P.total_doubles);
    size_t smem;
    int in_smem = 1, grid;
    if (narrow) {
        const size_t slots = sizeof(double) * RO_R * (size_t)(plant->ns + plant->nc) * 128;
        if (slots + wbytes > (size_t)ctx->smem_optin / RO_MINBLK - 1024)
            return fail(ctx, CLR_ERR_ARG, "controller too large for the blocked rollout of a run-time compiled plant");
        smem = slots + wbytes;
        const int64_t span = 128 * RO_R;
        grid = (int)std::min<int64_t>((A.N + span - 1) / span, (int64_t)ctx->sms * RO_MINBLK);
    } else {
        in_smem = wbytes <= (size_t)ctx->smem_optin - 1024;
        smem = in_smem ? wbytes : 0;
        const int per_sm = in_smem ? std::max(1, std::min(4, (int)((size_t)ctx->smem_optin / (smem + 1024)))) : 4;
        grid = (int)std::min<int64_t>((A.N + 127) / 128, (int64_t)ctx->sms * per_sm);
    }
    void* fn = (void*)plant->kern[slot];
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PointNet Pc = P;
    RolloutArgs Ac = A;
    void* args[3] = {&Pc, &Ac, &in_smem};
    CU(cudaLaunchKernel(fn, dim3(grid), dim3(128), args, smem, ctx->stream));
    ctx->launches++;
    return 0;
}

template <int PLANT>
int launch_rollout_p(clr_ctx* ctx, int code, const PointNet& P, const RolloutArgs& A) {
    if (split_eligible(ctx, code, P, A.N)) return launch_rollout_split<PLANT>(ctx, code, P, A);
    switch (code) {
        case 66: return launch_rollout_blocked<PLANT, 66>(ctx, P, A);
        case 72: return launch_rollout_blocked<PLANT, 72>(ctx, P, A);
        case 136: return launch_rollout_blocked<PLANT, 136>(ctx, P, A);
        case 1: return launch_rollout_t<PLANT, 1>(ctx, P, A);
        default: return launch_rollout_t<PLANT, 2>(ctx, P, A);
    }
}
}  // namespace

extern "C" {

}  // extern "C"

// ---- run-time compiled plants (the reference's plant is an arbitrary user function) -------------------------
namespace {

const char* const kRolloutSource =
#include "build/rollout_src.inc"
    ;

struct Nvrtc {
    void* so = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*AddNameExpression)(void*, const char*) = nullptr;
    int (*GetLoweredName)(void*, const char*, const char**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    if (n.so) return n;
    const char* cands[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (const char* c : cands) { n.so = dlopen(c, RTLD_NOW | RTLD_LOCAL); if (n.so) break; }
    if (!n.so) return n;
#define CLR_SYM(field, name) *(void**)(&n.field) = dlsym(n.so, name)
    CLR_SYM(CreateProgram, "nvrtcCreateProgram"); CLR_SYM(DestroyProgram, "nvrtcDestroyProgram");
    CLR_SYM(CompileProgram, "nvrtcCompileProgram"); CLR_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    CLR_SYM(GetProgramLog, "nvrtcGetProgramLog"); CLR_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    CLR_SYM(GetCUBIN, "nvrtcGetCUBIN"); CLR_SYM(AddNameExpression, "nvrtcAddNameExpression");
    CLR_SYM(GetLoweredName, "nvrtcGetLoweredName"); CLR_SYM(GetErrorString, "nvrtcGetErrorString");
#undef CLR_SYM
    n.ok = n.CreateProgram && n.DestroyProgram && n.CompileProgram && n.GetProgramLogSize && n.GetProgramLog &&
           n.GetCUBINSize && n.GetCUBIN && n.AddNameExpression && n.GetLoweredName && n.GetErrorString;
    return n;
}

// NVRTC-compile the plant's source for evaluator slot `slot` (kPlantCodes) and, with `load`, make the kernel launchable
int plant_build(clr_ctx* ctx, clr_plant* p, int slot, bool load) {
    if (p->kern[slot]) return 0;
    Nvrtc& n = nvrtc();
    if (!n.ok) return fail(ctx, CLR_ERR_STATE, "libnvrtc.so.12 could not be loaded: run-time compiled plants are unavailable");
    void* prog = nullptr;
    int rc = n.CreateProgram(&prog, p->source.c_str(), "clr_user_plant.cu", 0, nullptr, nullptr);
    if (rc) return fail(ctx, CLR_ERR_STATE, "nvrtcCreateProgram: %s", n.GetErrorString(rc));
    char expr[96];
    if (slot == 5) snprintf(expr, sizeof(expr), "rollout_ode_kernel<1000>");
    else snprintf(expr, sizeof(expr), kPlantCodes[slot] > 2 ? "rollout_blocked_kernel<1000, %d>" : "rollout_kernel<1000, %d>", kPlantCodes[slot]);
    n.AddNameExpression(prog, expr);
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo"};
    rc = n.CompileProgram(prog, 4, opts);
    if (rc) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls ? ls : 1, '\0');
        if (ls) n.GetProgramLog(prog, &log[0]);
        n.DestroyProgram(&prog);
        if (log.size() > 400) log = log.substr(log.size() - 400);   // the user's body is at the end of the source
        return fail(ctx, CLR_ERR_ARG, "plant right-hand side does not compile: %s", log.c_str());
    }
    if (!load) { n.DestroyProgram(&prog); return 0; }
    size_t cs = 0;
    n.GetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    n.GetCUBIN(prog, cubin.data());
    const char* low = nullptr;
    std::string name = (n.GetLoweredName(prog, expr, &low) == 0 && low) ? low : "";
    n.DestroyProgram(&prog);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaLibraryLoadData(&p->lib[slot], cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&p->kern[slot], p->lib[slot], name.c_str());
    if (e != cudaSuccess) {
        if (p->lib[slot]) { cudaLibraryUnload(p->lib[slot]); p->lib[slot] = nullptr; }
        p->kern[slot] = nullptr;
        return fail(ctx, CLR_ERR_CUDA, "loading the compiled plant failed: %s", cudaGetErrorString(e));
    }
    return 0;
}

}  // namespace

extern "C" {

int32_t clr_plant_compile(clr_ctx* ctx, int32_t n_state, int32_t n_dist, int32_t n_ctrl, const char* rhs_body,
                          clr_plant** out) {
    if (!out || !rhs_body) return fail(ctx, CLR_ERR_ARG, "clr_plant_compile: NULL argument");
    *out = nullptr;
    if (n_state < 1 || n_state > 32 || n_dist < 0 || n_ctrl < 1 || n_dist + n_ctrl > 32)
        return fail(ctx, CLR_ERR_DIM, "plant dimensions out of range (1..32 states, at most 32 disturbances + controls)");
    char defs[768];
    snprintf(defs, sizeof(defs),
             "#define CLR_DEV_MAX_LAYERS %d\n"
             "enum { CLR_ACT_ID = 0, CLR_ACT_RELU = 1, CLR_ACT_SIGMOID = 2, CLR_ACT_TANH = 3 };\n"
             "enum { CLR_ALG_TSIT5 = 0, CLR_ALG_RK4 = 1 };\n"
             "enum { CLR_PLANT_UNICYCLE = 1, CLR_PLANT_TORA = 2, CLR_PLANT_ACC = 3, CLR_PLANT_INVPEND = 4 };\n"
             "#define CLR_ERR_CAPACITY (-2)\n#define CLR_ERR_UNSTABLE (-7)\n"
             "#define CLR_USER_NS %d\n#define CLR_USER_ND %d\n#define CLR_USER_NC %d\n"
             "__device__ void clr_user_rhs(const double* x, const double* q, double* dx);\n",
             CLR_DEV_MAX_LAYERS, n_state, n_dist, n_ctrl);
    clr_plant* p = new clr_plant();
    p->ctx = ctx; p->ns = n_state; p->nd = n_dist; p->nc = n_ctrl;
    p->source = defs;
    p->source += kRolloutSource;
    p->source += "\n__device__ void clr_user_rhs(const double* x, const double* q, double* dx) {\n";
    p->source += rhs_body;
    p->source += "\n}\n";
    // compile one evaluator now: the caller learns about errors in the right-hand side here, not at the first rollout
    int rc = plant_build(ctx, p, 3 /* code 1: generic evaluator */, ctx != nullptr);
    if (rc) { delete p; return rc; }
    *out = p;
    return 0;
}

void clr_plant_free(clr_plant* p) {
    if (!p) return;
    if (p->ctx) cudaSetDevice(p->ctx->device);
    for (int i = 0; i < 6; i++) if (p->lib[i]) cudaLibraryUnload(p->lib[i]);
    delete p;
}

}  // extern "C"

extern "C" {

int32_t clr_forward_points(clr_ctx* ctx, const clr_net* net_c, const clr_prepost* pp, int32_t nx,
                           int64_t N, const double* X, double* U, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    clr_net* net = const_cast<clr_net*>(net_c);
    if (!net || nx < 1 || N < 0) return fail(ctx, CLR_ERR_ARG, "clr_forward_points: bad argument");
    if (N > 0 && (!X || !U)) return fail(ctx, CLR_ERR_ARG, "clr_forward_points: NULL buffer");
    CU(cudaSetDevice(ctx->device));
    Compiled* c = nullptr;
    int rc = compile_net(ctx, net, pp, nx, &c);
    if (rc) return rc;
    if (N == 0) return 0;
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    const int m = c->host.m_out;
    const void* dX;
    double* dU = U;
    ctx->bounce_open = false;
    if ((rc = stage_in(ctx, 0, X, sizeof(double) * (size_t)nx * N, dev_io, &dX))) return rc;
    if (!dev_io) {
        if ((rc = ensure(ctx, &ctx->d_io[4], &ctx->io_bytes[4], sizeof(double) * (size_t)m * N))) return rc;
        dU = (double*)ctx->d_io[4];
    }
    if (split_eligible(ctx, c->point_code, c->pnet, N)) {
        ConstBank& B = g_cbank[ctx->device];
        std::lock_guard<std::mutex> lock(B.mu);
        if ((rc = cbank_acquire(ctx, B, c->pnet))) return rc;
        if ((rc = launch_ctrl_cw(ctx, c->point_code, c->pnet, nx, N, (const double*)dX, dU))) return rc;
        CU(cudaEventRecord(B.last_use, ctx->stream));
    } else {
        rc = launch_points_code(ctx, c->point_code, c->pnet, nx, N, (const double*)dX, dU);
        if (rc) return rc;
    }
    if (!dev_io) {
        CU(cudaMemcpyAsync(U, dU, sizeof(double) * (size_t)m * N, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

static int32_t rollout_impl(clr_ctx* ctx, const clr_net* net_c, const clr_prepost* pp, int32_t plant, const clr_plant* user,
                    int32_t n_state, int32_t n_dist, int32_t n_ctrl, int64_t N, int32_t iters,
                    const double* x0, const double* Wd, double t0, double tau, double T, int32_t alg,
                    double abstol, double reltol, int32_t substeps, int32_t pow_mode, double* X_end,
                    double* U, int32_t* naccept, int32_t* nreject, int32_t log_cap, int32_t* log_n,
                    double* log_t, double* log_u, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    clr_net* net = const_cast<clr_net*>(net_c);
    if (!net || N < 0 || iters < 0) return fail(ctx, CLR_ERR_ARG, "clr_rollout: bad argument");
    int ns, nd, nc;
    if (user) { ns = user->ns; nd = user->nd; nc = user->nc; plant = CLR_PLANT_UNICYCLE; }
    else switch (plant) {
        case CLR_PLANT_UNICYCLE: ns = 4; nd = 1; nc = 2; break;
        case CLR_PLANT_TORA: ns = 4; nd = 0; nc = 1; break;
        case CLR_PLANT_ACC: ns = 6; nd = 0; nc = 1; break;
        case CLR_PLANT_INVPEND: ns = 2; nd = 0; nc = 1; break;
        default: return fail(ctx, CLR_ERR_ARG, "unknown plant id %d", plant);
    }
    if (user && user->ctx != ctx) return fail(ctx, CLR_ERR_STATE, "the plant was compiled for another context (or without one: syntax check only)");
    if (n_state != ns || n_dist != nd || n_ctrl != nc)
        return fail(ctx, CLR_ERR_DIM, "plant %d has %d states, %d disturbances, %d controls; got %d, %d, %d", plant, ns, nd, nc, n_state, n_dist, n_ctrl);
    if (alg != CLR_ALG_TSIT5 && alg != CLR_ALG_RK4) return fail(ctx, CLR_ERR_ARG, "unknown ODE algorithm %d", alg);
    if (alg == CLR_ALG_RK4 && substeps < 1) return fail(ctx, CLR_ERR_ARG, "substeps must be >= 1");
    if (log_cap < 0 || (log_cap > 0 && (!log_n || !log_t || !log_u))) return fail(ctx, CLR_ERR_ARG, "log buffers missing");
    CU(cudaSetDevice(ctx->device));
    Compiled* c = nullptr;
    int rc = compile_net(ctx, net, pp, n_state, &c);
    if (rc) return rc;
    if (c->host.m_out != n_ctrl) return fail(ctx, CLR_ERR_DIM, "controller output dimension %d != n_ctrl %d", c->host.m_out, n_ctrl);
    if (N == 0 || iters == 0) return 0;
    if (!x0 || !X_end || !U || (nd > 0 && !Wd)) return fail(ctx, CLR_ERR_ARG, "clr_rollout: NULL buffer");
    const bool dev_io = flags & CLR_FLAG_DEVICE_IO;
    const size_t cells = (size_t)N * iters, ne = (size_t)ns + nd + nc;
    RolloutArgs A{};
    A.N = N; A.iters = iters; A.t0 = t0; A.tau = tau; A.T = T; A.alg = alg; A.abstol = abstol; A.reltol = reltol;
    A.substeps = substeps; A.pow_mode = pow_mode; A.log_cap = log_cap; A.error = ctx->d_err;
    CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    const void *dx0, *dWd = nullptr;
    ctx->bounce_open = false;
    if ((rc = stage_in(ctx, 0, x0, sizeof(double) * (size_t)ns * N, dev_io, &dx0))) return rc;
    if (nd > 0 && (rc = stage_in(ctx, 1, Wd, sizeof(double) * nd * cells, dev_io, &dWd))) return rc;
    A.x0 = (const double*)dx0; A.Wd = (const double*)dWd;
    // outputs: one scratch allocation in blocking mode
    unsigned char* base = nullptr;
    size_t oX = 0, oU = 0, oA = 0, oR = 0, oLn = 0, oLt = 0, oLu = 0, total = 0;
    auto take = [&](size_t bytes) { size_t o = total; total += (bytes + 255) & ~(size_t)255; return o; };
    if (dev_io) {
        A.X_end = X_end; A.U = U; A.naccept = naccept; A.nreject = nreject; A.log_n = log_n; A.log_t = log_t; A.log_u = log_u;
    } else {
        oX = take(sizeof(double) * ns * cells);
        oU = take(sizeof(double) * nc * cells);
        if (naccept) oA = take(sizeof(int) * cells);
        if (nreject) oR = take(sizeof(int) * cells);
        if (log_cap > 0) {
            oLn = take(sizeof(int) * cells);
            oLt = take(sizeof(double) * log_cap * cells);
            oLu = take(sizeof(double) * ne * log_cap * cells);
        }
        if ((rc = ensure(ctx, &ctx->d_io[4], &ctx->io_bytes[4], total))) return rc;
        base = (unsigned char*)ctx->d_io[4];
        A.X_end = (double*)(base + oX); A.U = (double*)(base + oU);
        A.naccept = naccept ? (int*)(base + oA) : nullptr;
        A.nreject = nreject ? (int*)(base + oR) : nullptr;
        if (log_cap > 0) { A.log_n = (int*)(base + oLn); A.log_t = (double*)(base + oLt); A.log_u = (double*)(base + oLu); }
    }
    if (user) rc = launch_rollout_user(ctx, user, c->point_code, c->pnet, A);
    else switch (plant) {
        case CLR_PLANT_UNICYCLE: rc = launch_rollout_p<CLR_PLANT_UNICYCLE>(ctx, c->point_code, c->pnet, A); break;
        case CLR_PLANT_TORA: rc = launch_rollout_p<CLR_PLANT_TORA>(ctx, c->point_code, c->pnet, A); break;
        case CLR_PLANT_ACC: rc = launch_rollout_p<CLR_PLANT_ACC>(ctx, c->point_code, c->pnet, A); break;
        default: rc = launch_rollout_p<CLR_PLANT_INVPEND>(ctx, c->point_code, c->pnet, A); break;
    }
    if (rc) return rc;
    if (!dev_io) {
        CU(cudaMemcpyAsync(X_end, A.X_end, sizeof(double) * ns * cells, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(U, A.U, sizeof(double) * nc * cells, cudaMemcpyDeviceToHost, ctx->stream));
        if (naccept) CU(cudaMemcpyAsync(naccept, A.naccept, sizeof(int) * cells, cudaMemcpyDeviceToHost, ctx->stream));
        if (nreject) CU(cudaMemcpyAsync(nreject, A.nreject, sizeof(int) * cells, cudaMemcpyDeviceToHost, ctx->stream));
        if (log_cap > 0) {
            CU(cudaMemcpyAsync(log_n, A.log_n, sizeof(int) * cells, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(log_t, A.log_t, sizeof(double) * log_cap * cells, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(log_u, A.log_u, sizeof(double) * ne * log_cap * cells, cudaMemcpyDeviceToHost, ctx->stream));
        }
        int err = 0;
        CU(cudaMemcpyAsync(&err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (err == CLR_ERR_UNSTABLE)
            return fail(ctx, err, "an integration gave up (non-finite error estimate or step, step below eps(t), or more than 1e5 "
                        "iterations in one control period): OrdinaryDiffEq's retcodes Unstable / DtLessThanMin / MaxIters");
        if (err) return fail(ctx, err, "step log overflow: a period took more than log_cap - 1 = %d accepted steps", log_cap - 1);
    }
    return 0;
}


int32_t clr_rollout(clr_ctx* ctx, const clr_net* net, const clr_prepost* pp, int32_t plant,
                    int32_t n_state, int32_t n_dist, int32_t n_ctrl, int64_t N, int32_t iters,
                    const double* x0, const double* Wd, double t0, double tau, double T, int32_t alg,
                    double abstol, double reltol, int32_t substeps, int32_t pow_mode, double* X_end,
                    double* U, int32_t* naccept, int32_t* nreject, int32_t log_cap, int32_t* log_n,
                    double* log_t, double* log_u, int32_t flags) {
    return rollout_impl(ctx, net, pp, plant, nullptr, n_state, n_dist, n_ctrl, N, iters, x0, Wd, t0, tau, T, alg, abstol, reltol,
                        substeps, pow_mode, X_end, U, naccept, nreject, log_cap, log_n, log_t, log_u, flags);
}

int32_t clr_rollout_plant(clr_ctx* ctx, const clr_net* net, const clr_prepost* pp, const clr_plant* plant, int64_t N,
                          int32_t iters, const double* x0, const double* Wd, double t0, double tau, double T, int32_t alg,
                          double abstol, double reltol, int32_t substeps, int32_t pow_mode, double* X_end, double* U,
                          int32_t* naccept, int32_t* nreject, int32_t log_cap, int32_t* log_n, double* log_t,
                          double* log_u, int32_t flags) {
    if (!ctx) return CLR_ERR_ARG;
    if (!plant) return fail(ctx, CLR_ERR_ARG, "clr_rollout_plant: NULL plant");
    return rollout_impl(ctx, net, pp, 0, plant, plant->ns, plant->nd, plant->nc, N, iters, x0, Wd, t0, tau, T, alg, abstol,
                        reltol, substeps, pow_mode, X_end, U, naccept, nreject, log_cap, log_n, log_t, log_u, flags);
}

}  // extern "C"

#include "multi.inc"
