// FP32-pipe microbenchmarks for B200 (sm_100a): the measured denominator for the
// Chamfer roofline (MEASURED_PEAKS.json carries no FP32 figure), plus instruction-mix
// experiments that bound the nearest-neighbour inner loops (FFMA2 / FADD2 / FMNMX3).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp32_peak fp32_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---- (1) scalar FFMA chains -------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b, unsigned long long* clk) {
    float v[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = threadIdx.x * 1e-3f + i;
    unsigned long long t0 = gtimer(); long long c0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < CH; i++) v[i] = fmaf(v[i], a, b);
    }
    long long c1 = clock64(); unsigned long long t1 = gtimer();
    float s = 0; for (int i = 0; i < CH; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = (unsigned long long)(c1 - c0); clk[1] = t1 - t0; }
}
// ---- (2) packed FFMA2 chains -----------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b) {
    float2 v[CH]; float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < CH; i++) v[i] = __ffma2_rn(v[i], aa, bb);
    }
    float s = 0; for (int i = 0; i < CH; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- (3) FMNMX3 chains (ALU pipe) -------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) k_min3(float* out, int iters, const float* __restrict__ in) {
    float v[CH]; float x = in[threadIdx.x], y = in[threadIdx.x + 256];
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < CH; i++) { v[i] = fminf(fminf(v[i], x), y); x += 1.0f; }
    }
    float s = 0; for (int i = 0; i < CH; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- (4) mix: NF FFMA2 + NM FMNMX3 per group, the screen loop's 3:1 mix -------
template <int NF, int NM>
__global__ void __launch_bounds__(256) k_mix(float* out, int iters, float a, float b) {
    float2 v[8]; float m[4]; float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = make_float2(threadIdx.x * 1e-3f + i, i);
#pragma unroll
    for (int i = 0; i < 4; i++) m[i] = 1e30f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < NF; i++) v[i % 8] = __ffma2_rn(v[i % 8], aa, bb);
#pragma unroll
            for (int i = 0; i < NM; i++) m[i % 4] = fminf(fminf(m[i % 4], v[i % 8].x), v[i % 8].y);
        }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += v[i].x + v[i].y; for (int i = 0; i < 4; i++) s += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1) { float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms; }

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clockRate(kHz) %d\n", p.name, sms, p.clockRate);
    float* out; CK(cudaMalloc(&out, 148 * 16 * 1024 * sizeof(float)));
    float* in; CK(cudaMalloc(&in, 4096)); CK(cudaMemset(in, 0, 4096));
    unsigned long long* clk; CK(cudaMalloc(&clk, 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    for (int bps : {2, 4, 8}) {
        int grid = sms * bps;
        auto run = [&](const char* name, auto launch, double flop_per_thread, double inst_per_thread) {
            for (int w = 0; w < 2; w++) launch(grid);
            CK(cudaDeviceSynchronize());
            float best = 1e30f;
            for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); launch(grid); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); best = std::min(best, time_ms(0, e0, e1)); }
            double thr = (double)grid * 256;
            unsigned long long h[2] = {0, 0}; CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
            double mhz = h[1] ? 1e3 * (double)h[0] / (double)h[1] : 0;
            printf("%-22s blocks/SM %d  %.3f ms  %.2f TFLOP/s  %.3f warp-inst/clk/SM(@1965MHz)  clk(obs) %.0f MHz\n", name, bps, best,
                   flop_per_thread * thr / best * 1e-9, inst_per_thread * thr / 32.0 / (best * 1e-3 * 1965e6) / sms, mhz);
        };
        run("ffma x8 chains", [&](int g) { k_ffma<8><<<g, 256>>>(out, iters, 1.0001f, 0.5f, clk); }, 2.0 * 8 * 8 * iters, 8.0 * 8 * iters);
        run("ffma2 x8 chains", [&](int g) { k_ffma2<8><<<g, 256>>>(out, iters, 1.0001f, 0.5f); }, 4.0 * 8 * 8 * iters, 8.0 * 8 * iters);
        run("fmnmx3 x8 chains", [&](int g) { k_min3<8><<<g, 256>>>(out, iters, in); }, 0, 2.0 * 8 * 8 * iters);
        run("mix 6 ffma2 + 2 min3", [&](int g) { k_mix<6, 2><<<g, 256>>>(out, iters, 1.0001f, 0.5f); }, 4.0 * 6 * 4 * iters, 8.0 * 4 * iters);
        run("mix 6 ffma2 + 4 min3", [&](int g) { k_mix<6, 4><<<g, 256>>>(out, iters, 1.0001f, 0.5f); }, 4.0 * 6 * 4 * iters, 10.0 * 4 * iters);
    }
    CK(cudaGetLastError());
    return 0;
}
