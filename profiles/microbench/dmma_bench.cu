// Microbenchmark: FP64 tensor (DMMA via mma.sync) vs DFMA throughput on sm_100a.
// Evidence for the kernel design in DESIGN.md (no library calls; timed with CUDA events).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* d, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---- pure issue throughput: CH independent accumulator chains per warp
template <int CH>
__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
    double d[CH][2];
#pragma unroll
    for (int c = 0; c < CH; c++) { d[c][0] = threadIdx.x; d[c][1] = c; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) dmma884(d[c][0], d[c][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += d[c][0] + d[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dmma1684(double* out, int iters, double a0, double b0) {
    double d[CH][4];
#pragma unroll
    for (int c = 0; c < CH; c++) { d[c][0] = threadIdx.x; d[c][1] = c; d[c][2] = 1; d[c][3] = 2; }
    double a[2] = {a0 + threadIdx.x * 1e-9, a0}; double b = b0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) dmma1684(d[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dmma1688(double* out, int iters, double a0, double b0) {
    double d[CH][4];
#pragma unroll
    for (int c = 0; c < CH; c++) { d[c][0] = threadIdx.x; d[c][1] = c; d[c][2] = 1; d[c][3] = 2; }
    double a[4] = {a0 + threadIdx.x * 1e-9, a0, a0 * 2, a0 * 3}; double b[2] = {b0, b0 * 0.5};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) dmma1688(d[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
    double d[CH][4];
#pragma unroll
    for (int c = 0; c < CH; c++) { d[c][0] = threadIdx.x; d[c][1] = c; d[c][2] = 1; d[c][3] = 2; }
    double a[8]; double b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = a0 * (i + 1) + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = b0 * (i + 1);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) dmma16816(d[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double d[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) d[c] = threadIdx.x + c;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) d[c] = fma(d[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += d[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- W-in-registers pattern: warp owns 8 rows, K=100 (25 A frags in regs), B from smem
// X tile in smem: [NCOL][100] doubles (pitch 100: 100 mod 16 == 4 -> conflict free)
template <int NB>   // NB = column blocks (8 cols each) processed concurrently per warp
__global__ void k_wreg_smem(double* out, int iters, const double* __restrict__ Wg, int ncol) {
    extern __shared__ double xs[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ncol * 100; i += blockDim.x) xs[i] = 1e-3 * (i % 97);
    __syncthreads();
    double a[25];
#pragma unroll
    for (int k = 0; k < 25; k++) a[k] = Wg[(warp * 8 + (lane >> 2)) * 100 + k * 4 + (lane & 3)];
    double acc = 0;
    for (int it = 0; it < iters; it++) {
        for (int cb = 0; cb < ncol / 8; cb += NB) {
            double d[NB][2];
#pragma unroll
            for (int j = 0; j < NB; j++) { d[j][0] = 0; d[j][1] = 0; }
#pragma unroll
            for (int k = 0; k < 25; k++) {
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    double b = xs[((cb + j) * 8 + (lane >> 2)) * 100 + k * 4 + (lane & 3)];
                    dmma884(d[j][0], d[j][1], a[k], b);
                }
            }
#pragma unroll
            for (int j = 0; j < NB; j++) acc += fabs(d[j][0]) + fabs(d[j][1]);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// 16-row variant: warp owns 16 rows (two m8 tiles), 50 A regs, B frag reused by both tiles
template <int NB>
__global__ void k_wreg16_smem(double* out, int iters, const double* __restrict__ Wg, int ncol) {
    extern __shared__ double xs[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ncol * 100; i += blockDim.x) xs[i] = 1e-3 * (i % 97);
    __syncthreads();
    double a0[25], a1[25];
#pragma unroll
    for (int k = 0; k < 25; k++) {
        a0[k] = Wg[((warp * 16 + (lane >> 2)) % 100) * 100 + k * 4 + (lane & 3)];
        a1[k] = Wg[((warp * 16 + 8 + (lane >> 2)) % 100) * 100 + k * 4 + (lane & 3)];
    }
    double acc = 0;
    for (int it = 0; it < iters; it++) {
        for (int cb = 0; cb < ncol / 8; cb += NB) {
            double d[NB][4];
#pragma unroll
            for (int j = 0; j < NB; j++) { d[j][0] = 0; d[j][1] = 0; d[j][2] = 0; d[j][3] = 0; }
#pragma unroll
            for (int k = 0; k < 25; k++) {
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    double b = xs[((cb + j) * 8 + (lane >> 2)) * 100 + k * 4 + (lane & 3)];
                    dmma884(d[j][0], d[j][1], a0[k], b);
                    dmma884(d[j][2], d[j][3], a1[k], b);
                }
            }
#pragma unroll
            for (int j = 0; j < NB; j++) acc += fabs(d[j][0]) + fabs(d[j][1]) + fabs(d[j][2]) + fabs(d[j][3]);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}


// ---- do DMMA and DFMA share a pipe?  warps 0..7 of a CTA run DMMA chains, warps 8..15 DFMA chains (mode bit 0 / bit 1)
__global__ void k_mixed(double* out, int iters_mma, int iters_fma, int mode, double a0, double b0) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp < 8) {
        if (!(mode & 1)) return;
        double d[4][2];
#pragma unroll
        for (int c = 0; c < 4; c++) { d[c][0] = threadIdx.x; d[c][1] = c; }
        double a = a0 + threadIdx.x * 1e-9, b = b0;
        for (int i = 0; i < iters_mma; i++) {
#pragma unroll
            for (int c = 0; c < 4; c++) dmma884(d[c][0], d[c][1], a, b);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) s += d[c][0] + d[c][1];
    } else {
        if (!(mode & 2)) return;
        double d[8];
#pragma unroll
        for (int c = 0; c < 8; c++) d[c] = threadIdx.x + c;
        double a = a0 + threadIdx.x * 1e-9, b = b0;
        for (int i = 0; i < iters_fma; i++) {
#pragma unroll
            for (int c = 0; c < 8; c++) d[c] = fma(d[c], a, b);
        }
#pragma unroll
        for (int c = 0; c < 8; c++) s += d[c];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f, int reps = 3) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 4));
    double* W; CK(cudaMalloc(&W, sizeof(double) * 104 * 100));
    CK(cudaMemset(W, 0, sizeof(double) * 104 * 100));
    const int iters = 4096;
    // pure issue throughput, one CTA per SM x ctas_per_sm, varying warps
    int warps_list[] = {4, 8, 13, 16, 26, 32};
    for (int w : warps_list) {
        int threads = w * 32; int grid = sms * (w <= 16 ? 2 : 1);
        double nw = (double)grid * w;
#define RUN(NAME, KERNEL, CH, FLOPS) { float ms = timeit([&] { KERNEL<CH><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }); \
        printf("%-12s warps/cta %2d ctas %3d chains %d : %8.3f ms  %7.2f TFLOP/s\n", NAME, w, grid, CH, ms, nw * iters * CH * (double)(FLOPS) / ms / 1e9); }
        RUN("dmma884", k_dmma884, 1, 512); RUN("dmma884", k_dmma884, 2, 512); RUN("dmma884", k_dmma884, 4, 512); RUN("dmma884", k_dmma884, 8, 512);
        RUN("dmma1684", k_dmma1684, 1, 1024); RUN("dmma1684", k_dmma1684, 4, 1024);
        RUN("dmma1688", k_dmma1688, 1, 2048); RUN("dmma1688", k_dmma1688, 4, 2048);
        RUN("dmma16816", k_dmma16816, 1, 4096); RUN("dmma16816", k_dmma16816, 4, 4096);
        RUN("dfma", k_dfma, 1, 64); RUN("dfma", k_dfma, 4, 64); RUN("dfma", k_dfma, 8, 64);
    }
    // DMMA and DFMA warps side by side: separate pipes (time = max) or a shared one (time = sum)?
    {
        const int im = 4096, ifm = 4096 * 4;   // 8 warps x 4096 x 4 DMMA (512 flop) vs 8 warps x 16384 x 8 DFMA (64 flop): equal flops
        for (int mode = 1; mode <= 3; mode++) {
            float ms = timeit([&] { k_mixed<<<sms, 512>>>(out, im, ifm, mode, 1.0000001, 1e-9); });
            double fl = (double)sms * 8 * ((mode & 1 ? (double)im * 4 * 512 : 0) + (mode & 2 ? (double)ifm * 8 * 64 : 0));
            printf("mixed mode %d (1 = DMMA warps, 2 = DFMA warps, 3 = both): %8.3f ms %7.2f TFLOP/s\n", mode, ms, fl / ms / 1e9);
        }
    }
    // W-in-registers + B from smem
    {
        int ncol = 128; size_t smem = (size_t)ncol * 100 * 8;
        CK(cudaFuncSetAttribute(k_wreg_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_wreg_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_wreg_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_wreg16_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_wreg16_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int it2 = 200;
        for (int ctas = 1; ctas <= 2; ctas++) {
            int grid = sms * ctas;
            double fl = (double)grid * 13 * it2 * (ncol / 8) * 25 * 512.0;
            float ms;
            ms = timeit([&] { k_wreg_smem<1><<<grid, 13 * 32, smem>>>(out, it2, W, ncol); });
            printf("wreg8  smemB NB=1 ctas/sm %d: %8.3f ms %7.2f TFLOP/s\n", ctas, ms, fl / ms / 1e9);
            ms = timeit([&] { k_wreg_smem<2><<<grid, 13 * 32, smem>>>(out, it2, W, ncol); });
            printf("wreg8  smemB NB=2 ctas/sm %d: %8.3f ms %7.2f TFLOP/s\n", ctas, ms, fl / ms / 1e9);
            ms = timeit([&] { k_wreg_smem<4><<<grid, 13 * 32, smem>>>(out, it2, W, ncol); });
            printf("wreg8  smemB NB=4 ctas/sm %d: %8.3f ms %7.2f TFLOP/s\n", ctas, ms, fl / ms / 1e9);
            double fl16 = (double)grid * 7 * it2 * (ncol / 8) * 25 * 1024.0;
            ms = timeit([&] { k_wreg16_smem<1><<<grid, 7 * 32, smem>>>(out, it2, W, ncol); });
            printf("wreg16 smemB NB=1 ctas/sm %d (7 warps): %8.3f ms %7.2f TFLOP/s\n", ctas, ms, fl16 / ms / 1e9);
            ms = timeit([&] { k_wreg16_smem<2><<<grid, 7 * 32, smem>>>(out, it2, W, ncol); });
            printf("wreg16 smemB NB=2 ctas/sm %d (7 warps): %8.3f ms %7.2f TFLOP/s\n", ctas, ms, fl16 / ms / 1e9);
        }
    }
    return 0;
}
