// Variants of the 3-FFMA screening inner loop, to find the schedule that keeps the FMA pipe busiest.
// (Timing experiment only: results are written but not checked.)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
#define INF __int_as_float(0x7f800000)

struct Grp { float4 x, y, z, w; };
__device__ __forceinline__ Grp ld(const float* X, const float* Y, const float* Z, const float* W, int k) {
    Grp g; g.x = *(const float4*)(X + k); g.y = *(const float4*)(Y + k); g.z = *(const float4*)(Z + k); g.w = *(const float4*)(W + k); return g;
}
// LAYERED: issue all first-layer FMAs of the group, then the second, then the third
template <int R, bool LAYERED>
__device__ __forceinline__ void eval(const Grp& g, const float (&qx)[R], const float (&qy)[R], const float (&qz)[R], float (&cm)[R]) {
    if (LAYERED) {
        float2 t0[R], t1[R];
#pragma unroll
        for (int r = 0; r < R; r++) { float2 b = make_float2(qz[r], qz[r]);
            t0[r] = __ffma2_rn(make_float2(g.z.x, g.z.y), b, make_float2(g.w.x, g.w.y)); t1[r] = __ffma2_rn(make_float2(g.z.z, g.z.w), b, make_float2(g.w.z, g.w.w)); }
#pragma unroll
        for (int r = 0; r < R; r++) { float2 b = make_float2(qy[r], qy[r]);
            t0[r] = __ffma2_rn(make_float2(g.y.x, g.y.y), b, t0[r]); t1[r] = __ffma2_rn(make_float2(g.y.z, g.y.w), b, t1[r]); }
#pragma unroll
        for (int r = 0; r < R; r++) { float2 b = make_float2(qx[r], qx[r]);
            t0[r] = __ffma2_rn(make_float2(g.x.x, g.x.y), b, t0[r]); t1[r] = __ffma2_rn(make_float2(g.x.z, g.x.w), b, t1[r]); }
#pragma unroll
        for (int r = 0; r < R; r++) { cm[r] = fminf(fminf(cm[r], t0[r].x), t0[r].y); cm[r] = fminf(fminf(cm[r], t1[r].x), t1[r].y); }
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
            float2 bx = make_float2(qx[r], qx[r]), by = make_float2(qy[r], qy[r]), bz = make_float2(qz[r], qz[r]);
            float2 t0 = __ffma2_rn(make_float2(g.z.x, g.z.y), bz, make_float2(g.w.x, g.w.y));
            float2 t1 = __ffma2_rn(make_float2(g.z.z, g.z.w), bz, make_float2(g.w.z, g.w.w));
            t0 = __ffma2_rn(make_float2(g.y.x, g.y.y), by, t0); t1 = __ffma2_rn(make_float2(g.y.z, g.y.w), by, t1);
            t0 = __ffma2_rn(make_float2(g.x.x, g.x.y), bx, t0); t1 = __ffma2_rn(make_float2(g.x.z, g.x.w), bx, t1);
            cm[r] = fminf(fminf(cm[r], t0.x), t0.y); cm[r] = fminf(fminf(cm[r], t1.x), t1.y);
        }
    }
}

__device__ __forceinline__ float2 vfma2(float2 a, float b, float2 c) {
    float2 d;
    asm volatile("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mov.b64 rc, {%5, %6}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
                 : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b), "f"(c.x), "f"(c.y));
    return d;
}
// half-major: for each candidate pair, all R queries back to back per layer -> the pair operand can sit in the reuse cache
template <int R, bool VOL>
__device__ __forceinline__ void eval_hm(const Grp& g, const float (&qx)[R], const float (&qy)[R], const float (&qz)[R], float (&cm)[R]) {
    const float2 zp[2] = {make_float2(g.z.x, g.z.y), make_float2(g.z.z, g.z.w)};
    const float2 yp[2] = {make_float2(g.y.x, g.y.y), make_float2(g.y.z, g.y.w)};
    const float2 xp[2] = {make_float2(g.x.x, g.x.y), make_float2(g.x.z, g.x.w)};
    const float2 wp[2] = {make_float2(g.w.x, g.w.y), make_float2(g.w.z, g.w.w)};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        float2 t[R];
#pragma unroll
        for (int r = 0; r < R; r++) t[r] = VOL ? vfma2(zp[h], qz[r], wp[h]) : __ffma2_rn(zp[h], make_float2(qz[r], qz[r]), wp[h]);
#pragma unroll
        for (int r = 0; r < R; r++) t[r] = VOL ? vfma2(yp[h], qy[r], t[r]) : __ffma2_rn(yp[h], make_float2(qy[r], qy[r]), t[r]);
#pragma unroll
        for (int r = 0; r < R; r++) t[r] = VOL ? vfma2(xp[h], qx[r], t[r]) : __ffma2_rn(xp[h], make_float2(qx[r], qx[r]), t[r]);
#pragma unroll
        for (int r = 0; r < R; r++) cm[r] = fminf(fminf(cm[r], t[r].x), t[r].y);
    }
}
template <int R, int G, int T, int MODE, int MINB>
__global__ void __launch_bounds__(T, MINB) screen_hm(const float* __restrict__ q, const float* __restrict__ c, float* out, int M) {
    extern __shared__ __align__(16) float sm[];
    float *X = sm, *Y = sm + M, *Z = sm + 2 * M, *W = sm + 3 * M;
    for (int i = threadIdx.x; i < 4 * M; i += T) sm[i] = c[i];
    __syncthreads();
    float qx[R], qy[R], qz[R], best[R], s2[R]; int bc[R];
#pragma unroll
    for (int r = 0; r < R; r++) { int j = blockIdx.x * T * R + r * T + threadIdx.x;
        qx[r] = -2.f * q[j * 3]; qy[r] = -2.f * q[j * 3 + 1]; qz[r] = -2.f * q[j * 3 + 2]; best[r] = INF; s2[r] = INF; bc[r] = 0; }
    for (int c0 = 0; c0 < M; c0 += G) {
        float cm[R];
#pragma unroll
        for (int r = 0; r < R; r++) cm[r] = INF;
#pragma unroll
        for (int k = 0; k < G; k += 4) {
            Grp cur = ld(X, Y, Z, W, c0 + k);
            eval_hm<R, MODE == 1>(cur, qx, qy, qz, cm);
        }
#pragma unroll
        for (int r = 0; r < R; r++) { bool p = cm[r] < best[r]; s2[r] = fminf(s2[r], fmaxf(cm[r], best[r])); best[r] = fminf(cm[r], best[r]); bc[r] = p ? c0 : bc[r]; }
    }
#pragma unroll
    for (int r = 0; r < R; r++) { int j = blockIdx.x * T * R + r * T + threadIdx.x; out[j] = best[r] + bc[r] + s2[r]; }
}

template <int R, int G, int T, bool LAYERED, bool PREFETCH, int MINB>
__global__ void __launch_bounds__(T, MINB) screen_v(const float* __restrict__ q, const float* __restrict__ c, float* out, int M) {
    extern __shared__ __align__(16) float sm[];
    float *X = sm, *Y = sm + M, *Z = sm + 2 * M, *W = sm + 3 * M;
    for (int i = threadIdx.x; i < 4 * M; i += T) sm[i] = c[i];
    __syncthreads();
    float qx[R], qy[R], qz[R], best[R], s2[R]; int bc[R];
#pragma unroll
    for (int r = 0; r < R; r++) { int j = blockIdx.x * T * R + r * T + threadIdx.x;
        qx[r] = -2.f * q[j * 3]; qy[r] = -2.f * q[j * 3 + 1]; qz[r] = -2.f * q[j * 3 + 2]; best[r] = INF; s2[r] = INF; bc[r] = 0; }
    Grp nxt = ld(X, Y, Z, W, 0);
    for (int c0 = 0; c0 < M; c0 += G) {
        float cm[R];
#pragma unroll
        for (int r = 0; r < R; r++) cm[r] = INF;
#pragma unroll
        for (int k = 0; k < G; k += 4) {
            if (PREFETCH) {
                Grp cur = nxt;
                int kn = c0 + k + 4; kn = kn < M ? kn : 0;
                nxt = ld(X, Y, Z, W, kn);
                eval<R, LAYERED>(cur, qx, qy, qz, cm);
            } else {
                Grp cur = ld(X, Y, Z, W, c0 + k);
                eval<R, LAYERED>(cur, qx, qy, qz, cm);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) { bool p = cm[r] < best[r]; s2[r] = fminf(s2[r], fmaxf(cm[r], best[r])); best[r] = fminf(cm[r], best[r]); bc[r] = p ? c0 : bc[r]; }
    }
#pragma unroll
    for (int r = 0; r < R; r++) { int j = blockIdx.x * T * R + r * T + threadIdx.x; out[j] = best[r] + bc[r] + s2[r]; }
}

template <typename K> void bench(const char* name, K kern, int R, int T, const float* q, const float* c, float* out, int M, long nq) {
    int grid = (int)(nq / (T * R)); size_t smem = (size_t)4 * M * 4;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; w++) kern<<<grid, T, smem>>>(q, c, out, M);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { CK(cudaEventRecord(e0)); kern<<<grid, T, smem>>>(q, c, out, M); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    double pairs = (double)nq * M;
    printf("%-44s regs %3d occ %2d (%2d warps/SM)  %.3f ms  %.2f Tpair/s\n", name, fa.numRegs, occ, occ * T / 32, best, pairs / best * 1e-9);
}
int main() {
    const int M = 2048; const long nq = 640L * 2 * 2048;
    std::vector<float> hq(nq * 3), hc(4 * M);
    srand(1); for (auto& v : hq) v = rand() / (float)RAND_MAX; for (auto& v : hc) v = rand() / (float)RAND_MAX;
    float *q, *c, *out; CK(cudaMalloc(&q, hq.size() * 4)); CK(cudaMalloc(&c, hc.size() * 4)); CK(cudaMalloc(&out, nq * 4));
    CK(cudaMemcpy(q, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(c, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
#define B(R, G, T, L, P, MB) bench("R" #R " G" #G " T" #T " layered=" #L " prefetch=" #P " minb=" #MB, screen_v<R, G, T, L, P, MB>, R, T, q, c, out, M, nq)
#define H(R, G, T, MODE, MB) bench("halfmajor R" #R " G" #G " T" #T " volatile=" #MODE " minb=" #MB, screen_hm<R, G, T, MODE, MB>, R, T, q, c, out, M, nq)
    H(4, 32, 128, 0, 1); H(4, 32, 128, 1, 1); H(8, 32, 128, 0, 1); H(8, 32, 128, 1, 1); H(8, 32, 64, 1, 1); H(6, 32, 128, 1, 1); H(4, 16, 128, 1, 1); H(8, 16, 128, 1, 1);
    H(4, 32, 256, 1, 1); H(8, 32, 256, 1, 1); H(16, 32, 64, 1, 1); H(12, 32, 64, 1, 1);
    B(4, 16, 128, false, false, 1); B(4, 16, 128, true, false, 1); B(4, 16, 128, false, true, 1); B(4, 16, 128, true, true, 1);
    B(4, 32, 128, false, false, 1); B(4, 32, 128, true, false, 1); B(4, 32, 128, false, true, 1); B(4, 32, 128, true, true, 1);
    B(4, 32, 256, true, true, 1); B(4, 32, 256, false, false, 1); B(2, 32, 256, true, true, 1); B(2, 32, 256, false, false, 1);
    B(8, 32, 128, true, true, 1); B(8, 32, 64, true, true, 1); B(6, 32, 128, true, true, 1); B(4, 64, 128, true, true, 1);
    B(4, 32, 128, true, true, 7); B(2, 32, 128, true, true, 7); B(3, 32, 128, true, true, 1);
    return 0;
}
