// Prototype (timing only): evaluate each unordered (row, column) pair ONCE in the exact difference form and feed both
// directions -- the row minimum stays in registers as in nn_kernel<EXACT>, the column minimum goes through a
// thread-local min over the R rows, a warp redux.min on the float bits (d >= 0, so uint order == float order) and one
// 64-bit shared-memory atomicMin of (d_bits << 32 | row_group) per (warp, column).  Reported in ordered-pair
// equivalents (2 per evaluation) so that it compares directly with nn_proto / bench.py.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
#define INF __int_as_float(0x7f800000)

template <int R, int G, int T, int COLMODE>
__global__ void __launch_bounds__(T) sym_k(const float* __restrict__ q, const float* __restrict__ c, float* out, unsigned long long* colout, int M) {
    extern __shared__ __align__(16) float sm[];
    float *X = sm, *Y = sm + M, *Z = sm + 2 * M;
    unsigned long long* colkey = reinterpret_cast<unsigned long long*>(sm + 3 * M);
    for (int i = threadIdx.x; i < 3 * M; i += T) sm[i] = c[i];
    for (int i = threadIdx.x; i < M; i += T) colkey[i] = ~0ull;
    __syncthreads();
    float nqx[R], nqy[R], nqz[R], best[R]; int bc[R];
#pragma unroll
    for (int r = 0; r < R; r++) { int j = blockIdx.x * T * R + r * T + threadIdx.x;
        nqx[r] = -q[j * 3]; nqy[r] = -q[j * 3 + 1]; nqz[r] = -q[j * 3 + 2]; best[r] = INF; bc[r] = 0; }
    const int lane = threadIdx.x & 31;
    const unsigned long long group = blockIdx.x * (T / 32) + (threadIdx.x >> 5);
    for (int c0 = 0; c0 < M; c0 += G) {
        float cm[R];
#pragma unroll
        for (int r = 0; r < R; r++) cm[r] = INF;
#pragma unroll
        for (int k = 0; k < G; k += 4) {
            float4 x4 = *(const float4*)(X + c0 + k), y4 = *(const float4*)(Y + c0 + k), z4 = *(const float4*)(Z + c0 + k);
            float col[4] = {INF, INF, INF, INF};
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float2 bx = make_float2(nqx[r], nqx[r]), by = make_float2(nqy[r], nqy[r]), bz = make_float2(nqz[r], nqz[r]);
                float2 dx0 = __fadd2_rn(make_float2(x4.x, x4.y), bx), dx1 = __fadd2_rn(make_float2(x4.z, x4.w), bx);
                float2 dy0 = __fadd2_rn(make_float2(y4.x, y4.y), by), dy1 = __fadd2_rn(make_float2(y4.z, y4.w), by);
                float2 dz0 = __fadd2_rn(make_float2(z4.x, z4.y), bz), dz1 = __fadd2_rn(make_float2(z4.z, z4.w), bz);
                float2 t0 = __fmul2_rn(dy0, dy0), t1 = __fmul2_rn(dy1, dy1);
                t0 = __ffma2_rn(dx0, dx0, t0); t1 = __ffma2_rn(dx1, dx1, t1);
                t0 = __ffma2_rn(dz0, dz0, t0); t1 = __ffma2_rn(dz1, dz1, t1);
                cm[r] = fminf(fminf(cm[r], t0.x), t0.y); cm[r] = fminf(fminf(cm[r], t1.x), t1.y);
                if (COLMODE) { col[0] = fminf(col[0], t0.x); col[1] = fminf(col[1], t0.y); col[2] = fminf(col[2], t1.x); col[3] = fminf(col[3], t1.y); }
            }
            if (COLMODE == 1) {  // redux + one atomic per (warp, column)
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    unsigned m = __reduce_min_sync(0xffffffffu, __float_as_uint(col[u]));
                    if (lane == 0) atomicMin(&colkey[c0 + k + u], ((unsigned long long)m << 32) | group);
                }
            } else if (COLMODE == 2) {  // every lane issues the atomic (ptxas may aggregate)
#pragma unroll
                for (int u = 0; u < 4; u++) atomicMin(&colkey[c0 + k + u], ((unsigned long long)__float_as_uint(col[u]) << 32) | group);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) { bool p = cm[r] < best[r]; best[r] = p ? cm[r] : best[r]; bc[r] = p ? c0 : bc[r]; }
    }
    __syncthreads();
    if (COLMODE) for (int i = threadIdx.x; i < M; i += T) atomicMin(&colout[i], colkey[i]);  // cross-CTA merge (global, 64-bit)
#pragma unroll
    for (int r = 0; r < R; r++) { int j = blockIdx.x * T * R + r * T + threadIdx.x; out[j] = best[r] + bc[r]; }
}

template <typename K> void bench(const char* name, K kern, int R, int T, double ordered_per_eval, const float* q, const float* c, float* out, unsigned long long* col, int M, long nq) {
    int grid = (int)(nq / (T * R)); size_t smem = (size_t)3 * M * 4 + (size_t)M * 8;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; w++) kern<<<grid, T, smem>>>(q, c, out, col, M);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { CK(cudaEventRecord(e0)); kern<<<grid, T, smem>>>(q, c, out, col, M); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    double evals = (double)nq * M;
    printf("%-46s regs %3d occ %2d  %.3f ms  %.2f Teval/s  = %.2f Tpair/s ordered-equivalent\n", name, fa.numRegs, occ, best, evals / best * 1e-9, ordered_per_eval * evals / best * 1e-9);
}
int main() {
    const int M = 2048; const long nq = 640L * 2048;   // one direction's worth of rows: 640 pairs x 2048 rows x 2048 columns
    std::vector<float> hq(nq * 3), hc(3 * M);
    srand(1); for (auto& v : hq) v = rand() / (float)RAND_MAX; for (auto& v : hc) v = rand() / (float)RAND_MAX;
    float *q, *c, *out; unsigned long long* col;
    CK(cudaMalloc(&q, hq.size() * 4)); CK(cudaMalloc(&c, hc.size() * 4)); CK(cudaMalloc(&out, nq * 4)); CK(cudaMalloc(&col, M * 8)); CK(cudaMemset(col, 0xff, M * 8));
    CK(cudaMemcpy(q, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(c, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
#define S(R, G, T, MODE, OPE) bench("sym R" #R " G" #G " T" #T " colmode=" #MODE, sym_k<R, G, T, MODE>, R, T, OPE, q, c, out, col, M, nq)
    S(4, 32, 128, 0, 1.0);   // baseline: exact rows only (one direction), as nn_kernel<EXACT>
    S(4, 32, 128, 1, 2.0); S(8, 32, 128, 1, 2.0); S(8, 32, 64, 1, 2.0); S(4, 32, 256, 1, 2.0);
    S(4, 32, 128, 2, 2.0); S(8, 32, 128, 2, 2.0);
    return 0;
}
