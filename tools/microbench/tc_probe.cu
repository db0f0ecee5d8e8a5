// tcgen05 probe: can the screening scores s = W_c - 2 q.c of the NN search be produced by the tensor cores, and how exact are they?
//
// One CTA multiplies a [128 x 32] bf16 A tile (queries) by a [256 x 32] bf16 B tile (candidates) into a [128 x 256] fp32
// accumulator in TMEM (two tcgen05.mma kind::f16, M=128 N=256 K=16) and copies the accumulator out with tcgen05.ld.
//   test 0: random bf16 operands against a float64 product                      -> are the descriptors / layouts right?
//   test 1: the bf16x3 split of fp32 points (27 products per pair, K padded to 32) against float64 W - 2 q.c, unit-ball clouds
//   test 2: the same with clouds shifted by +1000 (large W, heavy cancellation)  -> how does the accumulator round?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run: ./tc_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 256, K = 32;
constexpr uint32_t LBO = 128, SBO = 512;   // bytes: next 8-element K chunk of the same 8 rows; next group of 8 rows

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((LBO >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
    return d;                 // base offset 0, layout type 0 = no swizzle
}

// canonical K-major no-swizzle layout: element (r, k) of a tile
__host__ __device__ inline uint32_t tile_off(int r, int k) { return (uint32_t)(r >> 3) * SBO + (uint32_t)(k >> 3) * LBO + (uint32_t)(r & 7) * 16 + (uint32_t)(k & 7) * 2; }

__global__ void __launch_bounds__(128) probe_kernel(const uint16_t *A, const uint16_t *B, float *out, int *status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem, *sB = smem + M * K * 2;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < M * K; e += 128) { const int r = e / K, k = e % K; *reinterpret_cast<uint16_t *>(sA + tile_off(r, k)) = A[e]; }
    for (int e = tid; e < N * K; e += 128) { const int r = e / K, k = e % K; *reinterpret_cast<uint16_t *>(sB + tile_off(r, k)) = B[e]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's async proxy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        // kind::f16 instruction descriptor: D = f32 (1 << 4), A = B = bf16 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int k = 0; k < K / 16; k++) {
            const uint64_t da = make_desc(smem_u32(sA) + k * 2 * LBO), db = make_desc(smem_u32(sB) + k * 2 * LBO);
            const uint32_t acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait for the MMAs
    bool ok = false;
    const long long t0 = clock64();
    while (!ok) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        ok = done != 0;
        if (!ok && clock64() - t0 > 2000000000ll) break;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && tid == 0) *status = 1;
    if (ok) {
        for (int c = 0; c < N / 32; c++) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
                         "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                           "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                           "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; i++) out[(size_t)(warp * 32 + lane) * N + c * 32 + i] = __uint_as_float(v[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

// ---- host ---------------------------------------------------------------------------------------------------------------
static uint16_t bf16_rn(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    const uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(r >> 16);
}
static float bf16_f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static void split3(float x, uint16_t p[3]) {
    float r = x;
    for (int i = 0; i < 3; i++) { p[i] = bf16_rn(r); r -= bf16_f(p[i]); }   // the residuals are exact in fp32
}
static double urand() { return rand() / (RAND_MAX + 1.0); }

static int run(const std::vector<uint16_t> &A, const std::vector<uint16_t> &B, std::vector<float> &out) {
    uint16_t *dA, *dB; float *dO; int *dS, st = 0;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, (size_t)M * N * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0xff, (size_t)M * N * 4); cudaMemset(dS, 0, 4);
    probe_kernel<<<1, 128, (M + N) * K * 2 + 1024>>>(dA, dB, dO, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
    out.resize((size_t)M * N);
    cudaMemcpy(out.data(), dO, (size_t)M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dS);
    return st;
}

int main() {
    srand(1);
    // ---- test 0: plain bf16 product
    {
        std::vector<uint16_t> A(M * K), B(N * K);
        for (auto &a : A) a = bf16_rn((float)(urand() * 2 - 1));
        for (auto &b : B) b = bf16_rn((float)(urand() * 2 - 1));
        std::vector<float> out;
        const int st = run(A, B, out);
        if (st) { printf("test0: status %d (MMA never completed or CUDA error)\n", st); return 1; }
        double maxerr = 0; int bad = 0;
        for (int i = 0; i < M; i++)
            for (int j = 0; j < N; j++) {
                double ref = 0;
                for (int k = 0; k < K; k++) ref += (double)bf16_f(A[i * K + k]) * (double)bf16_f(B[j * K + k]);
                const double err = fabs(out[(size_t)i * N + j] - ref);
                if (err > maxerr) maxerr = err;
                if (err > 1e-4) bad++;
            }
        printf("test0 plain bf16 [128x32]x[256x32]^T: max |err| = %.3e, entries off by > 1e-4: %d of %d, out[0][0..3] = %g %g %g %g\n", maxerr, bad,
               M * N, out[0], out[1], out[2], out[3]);
    }
    // ---- the split scheme on several input distributions; prints the worst error in units of S^2 (S = |q| + max |c|)
    struct Dist { const char *name; double shift, sq, sc; };
    const Dist dists[] = {{"unit cube", 0, 1, 1}, {"shift +1000", 1000, 1, 1}, {"shift -37.5", -37.5, 1, 1}, {"queries x 1e3", 0, 1e3, 1},
                          {"candidates x 1e3", 0, 1, 1e3}, {"both x 1e-3", 0, 1e-3, 1e-3}, {"both x 1e6", 0, 1e6, 1e6}, {"shift 3, x 1e-2", 3, 1e-2, 1e-2}};
    double worst = 0;
    for (const Dist &ds : dists) {
        double maxrel = 0, maxerr = 0;
        for (int seed = 0; seed < 4; seed++) {
            std::vector<float> q(M * 3), c(N * 3), W(N);
            for (auto &v : q) v = (float)((urand() * 2 - 1) * ds.sq + ds.shift);
            for (auto &v : c) v = (float)((urand() * 2 - 1) * ds.sc + ds.shift);
            double maxc = 0;
            for (int j = 0; j < N; j++) {
                W[j] = fmaf(c[j * 3 + 2], c[j * 3 + 2], fmaf(c[j * 3 + 1], c[j * 3 + 1], c[j * 3] * c[j * 3]));
                maxc = fmax(maxc, sqrt((double)W[j]));
            }
            std::vector<uint16_t> A(M * K, 0), B(N * K, 0);
            // per coordinate 8 products a_i b_j with i + j <= 5 (pieces numbered from 1): (1,1) (1,2) (2,1) (1,3) (2,2) (3,1) (2,3) (3,2)
            const int pa[8] = {0, 0, 1, 0, 1, 2, 1, 2}, pb[8] = {0, 1, 0, 2, 1, 0, 2, 1};
            for (int i = 0; i < M; i++) {
                for (int d = 0; d < 3; d++) {
                    uint16_t p[3]; split3(-2.0f * q[i * 3 + d], p);
                    for (int t = 0; t < 8; t++) A[i * K + d * 8 + t] = p[pa[t]];
                }
                for (int t = 0; t < 3; t++) A[i * K + 24 + t] = bf16_rn(1.0f);
            }
            for (int j = 0; j < N; j++) {
                for (int d = 0; d < 3; d++) {
                    uint16_t p[3]; split3(c[j * 3 + d], p);
                    for (int t = 0; t < 8; t++) B[j * K + d * 8 + t] = p[pb[t]];
                }
                uint16_t p[3]; split3(W[j], p);
                for (int t = 0; t < 3; t++) B[j * K + 24 + t] = p[t];
            }
            std::vector<float> out;
            const int st = run(A, B, out);
            if (st) { printf("split scheme: status %d\n", st); return 1; }
            for (int i = 0; i < M; i++) {
                const double qn = sqrt((double)q[i * 3] * q[i * 3] + (double)q[i * 3 + 1] * q[i * 3 + 1] + (double)q[i * 3 + 2] * q[i * 3 + 2]);
                const double S = qn + maxc;
                for (int j = 0; j < N; j++) {
                    const double ref = (double)W[j] - 2.0 * ((double)q[i * 3] * c[j * 3] + (double)q[i * 3 + 1] * c[j * 3 + 1] + (double)q[i * 3 + 2] * c[j * 3 + 2]);
                    const double err = fabs(out[(size_t)i * N + j] - ref);
                    maxerr = fmax(maxerr, err);
                    maxrel = fmax(maxrel, err / (S * S));
                }
            }
        }
        printf("split scheme, %-18s: max |s_tc - s_f64| = %.3e, max err / S^2 = 2^%.2f\n", ds.name, maxerr, log2(maxrel));
        worst = fmax(worst, maxrel);
    }
    printf("WORST err / S^2 = 2^%.2f  (allowed by the screening bound: 6 u = 2^-21.42; eps = 32 u S^2)\n", log2(worst));
    return 0;
}
