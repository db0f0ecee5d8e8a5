// How fast are shared-memory integer atomics on sm_100?  (The DCD epilogue and the gather backward build per-pair
// histograms of the argmin indices with them.)  One CTA per "pair": 4096 random bins into a 4096-entry shared histogram,
// as the kernels do; variants: fire-and-forget add, add with the old value used, and plain (racy) stores as the
// no-atomics yardstick.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int MODE, int T>
__global__ void __launch_bounds__(T) hist_kernel(const int *__restrict__ idx, int n, int reps, int *out) {
    extern __shared__ int h[];
    const int *my = idx + (size_t)blockIdx.x * n;
    for (int k = threadIdx.x; k < n; k += T) h[k] = 0;
    __syncthreads();
    int acc = 0;
    for (int r = 0; r < reps; r++) {
#pragma unroll 4
        for (int k = threadIdx.x; k < n; k += T) {
            const int b = my[k];
            if (MODE == 0) atomicAdd(&h[b], 1);
            else if (MODE == 1) acc += atomicAdd(&h[b], 1);
            else h[b] = k;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = h[my[0]] + acc;
}

template <int MODE, int T>
void run(const char *name, const int *idx, int pairs, int n, int reps, int *out) {
    const size_t smem = (size_t)n * 4;
    CK(cudaFuncSetAttribute(hist_kernel<MODE, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    hist_kernel<MODE, T><<<pairs, T, smem>>>(idx, n, reps, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 5; i++) {
        CK(cudaEventRecord(e0)); hist_kernel<MODE, T><<<pairs, T, smem>>>(idx, n, reps, out); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hist_kernel<MODE, T>, T, smem));
    const double ops = (double)pairs * n * reps;
    printf("%-34s T=%4d occ %2d  %8.3f us  %7.2f Gop/s  %.3f ops/clk/SM (148 SMs, 1965 MHz)\n", name, T, occ, best * 1e3, ops / best * 1e-6,
           ops / (best * 1e-3) / 148.0 / 1.965e9);
}

int main() {
    const int pairs = 640, n = 4096, reps = 8;
    std::vector<int> h((size_t)pairs * n);
    srand(3);
    for (auto &v : h) v = rand() % n;
    int *idx, *out; CK(cudaMalloc(&idx, h.size() * 4)); CK(cudaMalloc(&out, pairs * 4));
    CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    run<0, 256>("atomicAdd, result unused", idx, pairs, n, reps, out);
    run<1, 256>("atomicAdd, result used", idx, pairs, n, reps, out);
    run<2, 256>("plain store (no atomics)", idx, pairs, n, reps, out);
    run<0, 512>("atomicAdd, result unused", idx, pairs, n, reps, out);
    run<1, 512>("atomicAdd, result used", idx, pairs, n, reps, out);
    run<2, 512>("plain store (no atomics)", idx, pairs, n, reps, out);
    // identity bins: every lane its own bank, no two lanes of a warp in one bank
    for (size_t i = 0; i < h.size(); i++) h[i] = (int)(i % n);
    CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    run<0, 256>("atomicAdd unused, conflict-free", idx, pairs, n, reps, out);
    run<1, 256>("atomicAdd used, conflict-free", idx, pairs, n, reps, out);
    return 0;
}
