// TMEM read microbenchmark: how fast can the warps of one CTA pull a [128 x N] fp32 accumulator out of tensor memory?
// One CTA allocates all 512 columns; W warps (4 or 8; warp w reads TMEM lanes 32 (w % 4) ..) run `iters` rounds of
// tcgen05.ld.32x32b.xC + tcgen05.wait::ld over 256 columns, with 1 or 2 loads in flight, and fold the values into a running
// minimum (so nothing is optimised away).  Prints cycles per 128 x 256 tile (128 KB) for each configuration.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_read tmem_read.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(v, a)                                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, " \
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                 \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),       \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),            \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),           \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                         \
                 : "r"(a)                                                                                                            \
                 : "memory")
#define LD16(v, a)                                                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),       \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                          \
                 : "r"(a)                                                                                                            \
                 : "memory")
#define LD16x256(v, a)                                                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, " \
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                 \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),       \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),            \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),           \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                         \
                 : "r"(a)                                                                                                            \
                 : "memory")
#define WAITLD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

__device__ __forceinline__ float fold32(const uint32_t (&v)[32], float m) {
#pragma unroll
    for (int i = 0; i < 30; i += 3) m = fminf(m, fminf(fminf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), __uint_as_float(v[i + 2])));
    return fminf(m, fminf(__uint_as_float(v[30]), __uint_as_float(v[31])));
}
__device__ __forceinline__ float fold16(const uint32_t (&v)[16], float m) {
#pragma unroll
    for (int i = 0; i < 15; i += 3) m = fminf(m, fminf(fminf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), __uint_as_float(v[i + 2])));
    return fminf(m, __uint_as_float(v[15]));
}

// MODE 0: x32, one load in flight (load, wait, fold)      MODE 1: x32, two in flight (the production loop)
// MODE 2: x32, four in flight (4 x 32 registers)           MODE 3: x16, two in flight
// MODE 4: x32 two in flight, NO fold (pure read)
template <int MODE>
__global__ void __launch_bounds__(256) read_kernel(int warps, int iters, long long *cycles, float *sink) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    float m = 3e38f;
    long long t0 = 0, t1 = 0;
    if (warp < warps) {
        // 8 warps: warps w and w + 4 share a lane quadrant and split the 256 columns
        const int quad = warp & 3, half = warp >> 2, ncol = warps > 4 ? 128 : 256;
        const uint32_t base = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 128);
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
            if (MODE == 0) {
                for (int c = 0; c < ncol; c += 32) { uint32_t v[32]; LD32(v, base + c); WAITLD(); m = fold32(v, m); }
            } else if (MODE == 1 || MODE == 4) {
                uint32_t va[32], vb[32];
                LD32(va, base); WAITLD();
#pragma unroll
                for (int c = 0; c < 256; c += 64) {
                    if (c < ncol) {
                        LD32(vb, base + c + 32);
                        if (MODE == 1) m = fold32(va, m); else m = fminf(m, __uint_as_float(va[it & 31]));
                        WAITLD();
                        if (c + 64 < ncol) LD32(va, base + c + 64);
                        if (MODE == 1) m = fold32(vb, m); else m = fminf(m, __uint_as_float(vb[it & 31]));
                        if (c + 64 < ncol) WAITLD();
                    }
                }
            } else if (MODE == 2) {
                for (int c = 0; c < ncol; c += 128) {
                    uint32_t v0[32], v1[32], v2[32], v3[32];
                    LD32(v0, base + c); LD32(v1, base + c + 32); LD32(v2, base + c + 64); LD32(v3, base + c + 96);
                    WAITLD();
                    m = fold32(v0, m); m = fold32(v1, m); m = fold32(v2, m); m = fold32(v3, m);
                }
            } else if (MODE == 5) {
                // 16 lanes x 256 bits x 8: 64 columns of 16 lanes per instruction; the quadrant's two lane halves in turn
                uint32_t va[32], vb[32];
                for (int c = 0; c < ncol; c += 64) {
                    LD16x256(va, base + c);
                    LD16x256(vb, base + (16u << 16) + c);
                    WAITLD();
                    m = fold32(va, m); m = fold32(vb, m);
                }
            } else if (MODE == 3) {
                uint32_t va[16], vb[16];
                LD16(va, base); WAITLD();
#pragma unroll
                for (int c = 0; c < 256; c += 32) {
                    if (c < ncol) {
                        LD16(vb, base + c + 16);
                        m = fold16(va, m);
                        WAITLD();
                        if (c + 32 < ncol) LD16(va, base + c + 32);
                        m = fold16(vb, m);
                        if (c + 32 < ncol) WAITLD();
                    }
                }
            }
        }
        t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0 && warp < warps) cycles[warp] = t1 - t0;
    sink[threadIdx.x] = m;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE>
static void run(const char *name, int warps) {
    long long *dc, hc[8]; float *ds;
    cudaMalloc(&dc, 64); cudaMalloc(&ds, 1024);
    const int iters = 2000;
    read_kernel<MODE><<<1, 256>>>(warps, 10, dc, ds);
    read_kernel<MODE><<<1, 256>>>(warps, iters, dc, ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(hc, dc, 64, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; w++) mx = hc[w] > mx ? hc[w] : mx;
    const double per_tile = (double)mx / iters;
    printf("%-44s warps=%d: %8.1f cycles per [128 x 256] tile  (%6.1f B/cycle/SM)\n", name, warps, per_tile, 131072.0 / per_tile);
    cudaFree(dc); cudaFree(ds);
}

int main() {
    for (int warps = 4; warps <= 8; warps += 4) {
        run<0>("x32, one load in flight, min fold", warps);
        run<1>("x32, two loads in flight, min fold", warps);
        run<2>("x32, four loads per wait, min fold", warps);
        run<3>("x16, two loads in flight, min fold", warps);
        run<5>("16x256b.x8, two per wait, min fold", warps);
    }
    return 0;
}
