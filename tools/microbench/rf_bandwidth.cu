// Register-file operand bandwidth vs FMA-pipe rate on B200: how many distinct register words can an
// FFMA / FFMA2 stream read per cycle before it falls off the 1-per-cycle (FFMA) / 1-per-2-cycles (FFMA2) rate?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int N = 8;

// MODE 0: acc_i = fma2(A,   s_i, acc_i)   shared pair A (reusable), distinct scalars      reads/inst: 1 + 2
// MODE 1: acc_i = fma2(A_i, s,   acc_i)   distinct pairs, shared scalar                   reads/inst: 2 + 2
// MODE 2: acc_i = fma2(A_i, s_i, acc_i)   everything distinct                             reads/inst: 2 + 1 + 2
// MODE 3: acc_i = fma2(A_i, B_i, acc_i)   distinct pair x distinct pair                   reads/inst: 2 + 2 + 2
// MODE 4: acc_i = fma2(A,   s,   acc_i)   only the accumulator is distinct                reads/inst: 2
template <int MODE>
__global__ void __launch_bounds__(256) k2(float* out, const float* __restrict__ in, int iters) {
    float2 A[N], B[N], acc[N]; float s[N];
    const float* p = in + threadIdx.x * 64;
#pragma unroll
    for (int i = 0; i < N; i++) { A[i] = make_float2(p[i], p[i + 8]); B[i] = make_float2(p[i + 16], p[i + 24]); s[i] = p[i + 32]; acc[i] = make_float2(p[i + 40], p[i + 48]); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (MODE == 0) acc[i] = __ffma2_rn(A[0], make_float2(s[i], s[i]), acc[i]);
                if (MODE == 1) acc[i] = __ffma2_rn(A[i], make_float2(s[0], s[0]), acc[i]);
                if (MODE == 2) acc[i] = __ffma2_rn(A[i], make_float2(s[i], s[i]), acc[i]);
                if (MODE == 3) acc[i] = __ffma2_rn(A[i], B[i], acc[i]);
                if (MODE == 4) acc[i] = __ffma2_rn(A[0], make_float2(s[0], s[0]), acc[i]);
            }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < N; i++) r += acc[i].x + acc[i].y;
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
// scalar FFMA: MODE 0: acc_i = fma(a, b_i, acc_i) (2 distinct) ; 1: fma(a_i, b_i, acc_i) (3 distinct) ; 2: fma(a, b, acc_i) (1 distinct)
template <int MODE>
__global__ void __launch_bounds__(256) k1(float* out, const float* __restrict__ in, int iters) {
    float a[N], b[N], acc[N];
    const float* p = in + threadIdx.x * 64;
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = p[i]; b[i] = p[i + 8]; acc[i] = p[i + 16]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (MODE == 0) acc[i] = fmaf(a[0], b[i], acc[i]);
                if (MODE == 1) acc[i] = fmaf(a[i], b[i], acc[i]);
                if (MODE == 2) acc[i] = fmaf(a[0], b[0], acc[i]);
            }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < N; i++) r += acc[i];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
// FADD2 / FMUL2 forms of the exact kernel: d = c + q (pair + scalar), m = d*d, t = fma2(d,d,t)
template <int MODE>
__global__ void __launch_bounds__(256) k3(float* out, const float* __restrict__ in, int iters) {
    float2 A[N], acc[N]; float s[N];
    const float* p = in + threadIdx.x * 64;
#pragma unroll
    for (int i = 0; i < N; i++) { A[i] = make_float2(p[i], p[i + 8]); s[i] = p[i + 32]; acc[i] = make_float2(p[i + 40], p[i + 48]); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (MODE == 0) acc[i] = __fadd2_rn(acc[i], make_float2(s[i], s[i]));   // pair + distinct scalar
                if (MODE == 1) acc[i] = __fmul2_rn(acc[i], acc[i]);                     // square
                if (MODE == 2) acc[i] = __ffma2_rn(A[i], A[i], acc[i]);                 // fma(d,d,t)
            }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < N; i++) r += acc[i].x + acc[i].y;
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
template <typename F> void run(const char* name, F launch, double inst_per_thread, int grid) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    double warp_inst = inst_per_thread * grid * 256 / 32.0;
    printf("%-52s %.3f ms  %.3f cycles/inst/SMSP\n", name, best, (best * 1e-3 * 1965e6) * 148 * 4 / warp_inst);
}
int main() {
    float *out, *in; CK(cudaMalloc(&out, 148 * 8 * 256 * 4)); CK(cudaMalloc(&in, 256 * 64 * 4));
    float h[256 * 64]; for (int i = 0; i < 256 * 64; i++) h[i] = 1.0f + (i % 97) * 1e-3f; CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    const int iters = 20000, grid = 148 * 8; const double ipt = 4.0 * N * iters;
    run("FFMA2 d=fma2(A, s_i, acc_i)  [1+2 words]", [&] { k2<0><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FFMA2 d=fma2(A_i, s, acc_i)  [2+2 words]", [&] { k2<1><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FFMA2 d=fma2(A_i, s_i, acc_i) [2+1+2 words]", [&] { k2<2><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FFMA2 d=fma2(A_i, B_i, acc_i) [2+2+2 words]", [&] { k2<3><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FFMA2 d=fma2(A, s, acc_i)    [2 words]", [&] { k2<4><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FFMA  d=fma(a, b_i, acc_i)   [2 words]", [&] { k1<0><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FFMA  d=fma(a_i, b_i, acc_i) [3 words]", [&] { k1<1><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FFMA  d=fma(a, b, acc_i)     [1 word]", [&] { k1<2><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FADD2 d=acc_i + s_i          [2+1 words]", [&] { k3<0><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FMUL2 d=acc_i * acc_i        [2 words]", [&] { k3<1><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    run("FFMA2 d=fma2(A_i, A_i, acc_i) [2+2 words]", [&] { k3<2><<<grid, 256>>>(out, in, iters); }, ipt, grid);
    return 0;
}
