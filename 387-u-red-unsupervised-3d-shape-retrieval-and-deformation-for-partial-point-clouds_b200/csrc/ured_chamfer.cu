// ured_chamfer.cu -- B200 (sm_100a) kernels and C ABI for the Chamfer / DCD hot path.
//
// Replaces, for this one path, the reference's native op and the torch-op epilogue on top of it
// (paths relative to the reference tree, DCD/ = Density_aware_Chamfer_Distance/):
//   NmDistanceKernel / chamfer_cuda_forward        DCD/utils_v2/metrics/CD/chamfer3D/chamfer3D.cu:12-154
//   NmDistanceGradKernel / chamfer_cuda_backward   DCD/utils_v2/metrics/CD/chamfer3D/chamfer3D.cu:155-195
//   calc_cd / calc_dcd torch-op body               DCD/utils_v2/model_utils.py:13-70
//   torch.topk(cd_m, k, largest=False)             dataset/dataset_utils.py:1043-1051
// The declarations and the contract of every entry point live in include/ured_chamfer.h.
//
// Kernel plan (DESIGN.md has the numbers):
//   pack_kernel   xyz[count,n,3] -> padded SoA image X|Y|Z|W (+ max W per cloud)     HBM-bound, tiny
//   nn_kernel     both directions of the nearest-neighbour search in ONE launch.  Each CTA owns
//                 T*R query points of one (pair, direction), streams the opposing cloud through
//                 shared memory with TMA bulk copies (cp.async.bulk + mbarrier, 2 stages) and
//                 evaluates packed FP32 math (FADD2/FMUL2/FFMA2) with a register-resident chunk
//                 minimum (FMNMX3).  Two variants:
//                   EXACT   difference form d = fma(dz,dz,fma(dx,dx,dy*dy)) on every pair
//                           (the reference arithmetic, 6 FP32-pipe ops per pair);
//                   SCREEN  3-FFMA expansion form s = |c|^2 - 2 q.c as a filter, then the exact
//                           difference form only on the winning 32-candidate chunk; queries whose
//                           runner-up chunk is within a rigorous rounding bound of the winner are
//                           re-scanned exactly by their warp.  Output bits are identical.
//   dcd_fwd_kernel   per pair: shared-memory histograms of idx1/idx2, weights, loss/cd_p/cd_t
//   grad_*_kernel    gather + atomic scatter of the Chamfer backward, fused with the DCD chain rule
//   topk_kernel      k smallest (score, index) per row
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include <atomic>

#include "ured_chamfer.h"

namespace {

// statistics only (bench.py reports it as gpu_launches); no call's behaviour depends on it
std::atomic<unsigned long long> g_launches{0};
#define URED_COUNT_LAUNCH() g_launches.fetch_add(1, std::memory_order_relaxed)

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
thread_local char g_err[256] = "no error";

int fail_arg(int code, const char *what) {
    snprintf(g_err, sizeof(g_err), "ured_chamfer: %s", what);
    return code;
}
int check_cuda(cudaError_t e, const char *where) {
    if (e == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "ured_chamfer: %s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
#define URED_CUDA(call, where)                       \
    do {                                             \
        int rc_ = check_cuda((call), (where));       \
        if (rc_) return rc_;                         \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int pad32(int n) { return (n + 31) / 32 * 32; }

// ------------------------------------------------------------------------------------------
// packed cloud image
// ------------------------------------------------------------------------------------------
constexpr int kPackThreads = 256;

constexpr int kPackTail = 32;  // floats after X|Y|Z|W of each cloud: [0] = max W, rest unused (keeps blocks 128-byte aligned)
inline size_t cloud_stride(int np) { return (size_t)np * 4 + kPackTail; }

// one CTA per cloud; writes the cloud's block X|Y|Z|W (each np floats) | tail (max W)
// len (optional): number of valid points of each cloud (ragged batches); everything past it replicates the last valid point.
// One launch can pack two cloud sets (the two sides of a Chamfer call): blocks [0, count_a) take set a, the rest set b.
struct PackSet { const float *xyz; float *soa; const int *len; int n_max, np; };

__global__ void __launch_bounds__(kPackThreads) pack_kernel(const PackSet a, const PackSet bset, int count_a) {
    const bool second = (int)blockIdx.x >= count_a;
    const PackSet &ps = second ? bset : a;
    const size_t cloud = second ? blockIdx.x - count_a : blockIdx.x;
    const int n_max = ps.n_max, np = ps.np;
    const int *__restrict__ len = ps.len;
    const int n = len ? max(1, min(len[cloud], n_max)) : n_max;  // (an empty cloud is never read as candidates)
    const float *src = ps.xyz + cloud * (size_t)n_max * 3;
    float *X = ps.soa + cloud * ((size_t)np * 4 + kPackTail);
    float *Y = X + np, *Z = Y + np, *W = Z + np;
    float *wmax = W + np;
    float m = 0.0f;
    for (int k = threadIdx.x; k < np; k += kPackThreads) {
        int ks = k < n ? k : n - 1;  // padding replicates the last point: it can tie with it, never beat it
        float x = src[ks * 3 + 0], y = src[ks * 3 + 1], z = src[ks * 3 + 2];
        float w = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
        X[k] = x; Y[k] = y; Z[k] = z; W[k] = w;
        m = fmaxf(m, w);          // fmaxf drops NaN; non-finite inputs are outside the screening contract anyway
        if (!(w <= 3.0e38f)) m = __int_as_float(0x7f800000);  // inf/NaN norm: force the exact path
    }
    __shared__ float red[kPackThreads / 32];
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kPackThreads / 32; i++) m = fmaxf(m, red[i]);
        wmax[0] = m;
    }
}

struct PackedView {
    const float *soa;  // count blocks of cloud_stride(np) floats; any sub-range of clouds is itself a packed image
    int np;
};
PackedView view_packed(const void *packed, int n) {
    PackedView v;
    v.np = pad32(n);
    v.soa = (const float *)packed;
    return v;
}

// ------------------------------------------------------------------------------------------
// TMA / mbarrier helpers (PTX; SASS shows UBLKCP + SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 256-bit read-only global load (sm_100: LDG.E.256): one full 32-byte sector per lane and request
struct __align__(32) float8 { float v[8]; };
__device__ __forceinline__ float8 ldg256(const float *p) {
    float8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------------------
// nearest-neighbour kernel
// ------------------------------------------------------------------------------------------
constexpr int kNNThreads = 128;  // 4 warps, one per SM sub-partition; 5-6 CTAs resident per SM
constexpr int kTile = 1024;      // candidates per shared-memory stage (768 was tried: 7 CTAs/SM but 3% slower on cfg2)
constexpr int kStages = 2;
constexpr int kChunk = 32;       // candidates per running-minimum chunk (= the padding granule of the packed image)
constexpr float kInf = __builtin_huge_valf();

struct NNParams {
    const float *xyz[2];   // raw clouds (queries are read from here)
    const float *soa[2];   // packed images (candidates are streamed from here)
    float *dist[2];
    int *idx[2];
    int n[2];
    int np[2];
    int qtiles[2];  // query tiles per pair for direction 0 / 1
    int rep1, mod2;
    // candidate splitting: with nsplit > 1 every (pair, direction, query tile) is cut into nsplit CTAs, each
    // scanning a contiguous candidate range and writing its exact partial (d, idx) to part_*[split][B][n]
    int nsplit;
    int B;
    float *part_dist[2];
    int *part_idx[2];
    const int *len[2];  // optional valid point counts per cloud-1 / cloud-2 entry (ragged batches)
};

// the reference's pair arithmetic (chamfer3D.cu:32-35 as compiled by nvcc 12.9 for sm_100a):
// differences are candidate - query, d = fma(dz,dz, fma(dx,dx, dy*dy))
__device__ __forceinline__ float exact_d(float cx, float cy, float cz, float qx, float qy, float qz) {
    float dx = __fsub_rn(cx, qx), dy = __fsub_rn(cy, qy), dz = __fsub_rn(cz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

template <bool SCREEN, int R>
__global__ void __launch_bounds__(kNNThreads) nn_kernel(const NNParams p) {
    constexpr int NARR = SCREEN ? 4 : 3;
    constexpr int G = kChunk;
    extern __shared__ __align__(128) float stage_mem[];  // kStages * NARR * kTile floats
    __shared__ __align__(8) uint64_t full_bar[kStages];

    const int tid = threadIdx.x;
    // ---- work item: (pair b, direction, query tile) --------------------------------------
    const int per_pair = (p.qtiles[0] + p.qtiles[1]) * p.nsplit;
    const int b = blockIdx.x / per_pair;
    const int rem0 = blockIdx.x - b * per_pair;
    const int sp = rem0 % p.nsplit;  // candidate split handled by this CTA
    const int rem = rem0 / p.nsplit;
    const int dir = rem >= p.qtiles[0] ? 1 : 0;
    const int qt = dir ? rem - p.qtiles[0] : rem;
    const int c1 = b / p.rep1, c2 = b % p.mod2;
    const int cq = dir ? c2 : c1, cc = dir ? c1 : c2;
    // (ternaries, not p.x[dir]: dynamic indexing would copy the parameter block to local memory)
    const int nq = dir ? p.n[1] : p.n[0];                                  // row stride of the outputs (max points)
    const int ncp_max = dir ? p.np[0] : p.np[1];                             // padded stride of the candidate image
    const int *len_q = dir ? p.len[1] : p.len[0], *len_c = dir ? p.len[0] : p.len[1];
    const int nq_v = len_q ? max(0, min(len_q[cq], nq)) : nq;              // valid queries of this cloud
    const int nc = len_c ? max(0, min(len_c[cc], dir ? p.n[0] : p.n[1])) : (dir ? p.n[0] : p.n[1]);  // valid candidates
    const int ncp = (nc + kChunk - 1) / kChunk * kChunk;
    const float *__restrict__ qxyz = (dir ? p.xyz[1] : p.xyz[0]) + (size_t)cq * nq * 3;
    const float *__restrict__ csoa = (dir ? p.soa[0] : p.soa[1]) + (size_t)cc * ((size_t)ncp_max * 4 + kPackTail);

    float *out_d = p.nsplit == 1 ? (dir ? p.dist[1] : p.dist[0]) + (size_t)b * nq
                                 : (dir ? p.part_dist[1] : p.part_dist[0]) + ((size_t)sp * p.B + b) * nq;
    int *out_i = p.nsplit == 1 ? (dir ? p.idx[1] : p.idx[0]) + (size_t)b * nq
                               : (dir ? p.part_idx[1] : p.part_idx[0]) + ((size_t)sp * p.B + b) * nq;
    if (qt * R * kNNThreads >= nq_v || nc == 0) {
        // nothing to search: the reference kernel leaves its zero-filled outputs untouched in this case
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int j = (qt * R + r) * kNNThreads + tid;
            if (j < nq) { out_d[j] = 0.0f; out_i[j] = 0; }
        }
        return;
    }

    // candidate range of this split, in whole 32-candidate chunks
    const int chunks_per_split = (ncp / kChunk + p.nsplit - 1) / p.nsplit;
    const int k_lo = min(ncp, sp * chunks_per_split * kChunk);
    const int k_hi = min(ncp, k_lo + chunks_per_split * kChunk);

    // ---- pipeline prologue ---------------------------------------------------------------
    const int ntiles = (k_hi - k_lo + kTile - 1) / kTile;
    auto issue_tile = [&](int t) {
        const int s = t % kStages;
        const int k0 = k_lo + t * kTile;
        const int tk = min(kTile, k_hi - k0);
        const uint32_t bytes = (uint32_t)tk * sizeof(float);
        mbar_arrive_expect_tx(&full_bar[s], bytes * NARR);
#pragma unroll
        for (int a = 0; a < NARR; a++)
            tma_bulk_g2s(stage_mem + (s * NARR + a) * kTile, csoa + (size_t)a * ncp_max + k0, bytes, &full_bar[s]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        if (ntiles > 0) issue_tile(0);
        if (ntiles > 1) issue_tile(1);
    }

    // ---- queries in registers ------------------------------------------------------------
    // EXACT keeps -q (added to the candidate), SCREEN keeps -2q (multiplied); both scalings are exact and
    // are undone after the main loop, so only one copy of the query lives in registers
    float qx[R], qy[R], qz[R];
    float best[R], second[R];
    int bchunk[R];
    constexpr float kScale = SCREEN ? -2.0f : -1.0f, kUnscale = SCREEN ? -0.5f : -1.0f;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int j = (qt * R + r) * kNNThreads + tid;
        j = j < nq_v ? j : nq_v - 1;
        qx[r] = kScale * qxyz[j * 3 + 0]; qy[r] = kScale * qxyz[j * 3 + 1]; qz[r] = kScale * qxyz[j * 3 + 2];
        best[r] = kInf; second[r] = kInf; bchunk[r] = k_lo;
    }

    // ---- main loop over candidate tiles ----------------------------------------------------
    for (int t = 0; t < ntiles; t++) {
        const int s = t % kStages;
        const int k0 = k_lo + t * kTile;
        const int tk = min(kTile, k_hi - k0);
        mbar_wait(&full_bar[s], (uint32_t)(t / kStages) & 1u);
        const float *X = stage_mem + (s * NARR + 0) * kTile;
        const float *Y = stage_mem + (s * NARR + 1) * kTile;
        const float *Z = stage_mem + (s * NARR + 2) * kTile;
        const float *W = stage_mem + (s * NARR + (SCREEN ? 3 : 2)) * kTile;

        for (int c0 = 0; c0 < tk; c0 += G) {
            float cm[R];
#pragma unroll
            for (int r = 0; r < R; r++) cm[r] = kInf;
#pragma unroll
            for (int k = 0; k < G; k += 4) {
                const float4 x4 = *reinterpret_cast<const float4 *>(X + c0 + k);
                const float4 y4 = *reinterpret_cast<const float4 *>(Y + c0 + k);
                const float4 z4 = *reinterpret_cast<const float4 *>(Z + c0 + k);
                if (SCREEN) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(W + c0 + k);
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const float2 bx = make_float2(qx[r], qx[r]), by = make_float2(qy[r], qy[r]), bz = make_float2(qz[r], qz[r]);
                        float2 t0 = __ffma2_rn(make_float2(z4.x, z4.y), bz, make_float2(w4.x, w4.y));
                        float2 t1 = __ffma2_rn(make_float2(z4.z, z4.w), bz, make_float2(w4.z, w4.w));
                        t0 = __ffma2_rn(make_float2(y4.x, y4.y), by, t0);
                        t1 = __ffma2_rn(make_float2(y4.z, y4.w), by, t1);
                        t0 = __ffma2_rn(make_float2(x4.x, x4.y), bx, t0);
                        t1 = __ffma2_rn(make_float2(x4.z, x4.w), bx, t1);
                        cm[r] = fminf(fminf(cm[r], t0.x), t0.y);
                        cm[r] = fminf(fminf(cm[r], t1.x), t1.y);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const float2 bx = make_float2(qx[r], qx[r]), by = make_float2(qy[r], qy[r]), bz = make_float2(qz[r], qz[r]);
                        const float2 dx0 = __fadd2_rn(make_float2(x4.x, x4.y), bx), dx1 = __fadd2_rn(make_float2(x4.z, x4.w), bx);
                        const float2 dy0 = __fadd2_rn(make_float2(y4.x, y4.y), by), dy1 = __fadd2_rn(make_float2(y4.z, y4.w), by);
                        const float2 dz0 = __fadd2_rn(make_float2(z4.x, z4.y), bz), dz1 = __fadd2_rn(make_float2(z4.z, z4.w), bz);
                        float2 t0 = __fmul2_rn(dy0, dy0), t1 = __fmul2_rn(dy1, dy1);
                        t0 = __ffma2_rn(dx0, dx0, t0);
                        t1 = __ffma2_rn(dx1, dx1, t1);
                        t0 = __ffma2_rn(dz0, dz0, t0);
                        t1 = __ffma2_rn(dz1, dz1, t1);
                        cm[r] = fminf(fminf(cm[r], t0.x), t0.y);
                        cm[r] = fminf(fminf(cm[r], t1.x), t1.y);
                    }
                }
            }
            // chunk bookkeeping: strict '<' in ascending chunk order keeps the FIRST chunk holding the minimum
            const int chunk_id = k0 + c0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const bool better = cm[r] < best[r];
                if (SCREEN) second[r] = fminf(second[r], fmaxf(cm[r], best[r]));
                best[r] = fminf(best[r], cm[r]);
                bchunk[r] = better ? chunk_id : bchunk[r];
            }
        }
        if (t + kStages < ntiles) {  // the stage is reused: wait until every thread is done reading it
            __syncthreads();
            if (tid == 0) issue_tile(t + kStages);
        }
    }

    // ---- exact resolution of the winning chunk ---------------------------------------------
    const float *__restrict__ gX = csoa, *__restrict__ gY = csoa + ncp_max, *__restrict__ gZ = csoa + 2 * (size_t)ncp_max;
    float cn = 0.0f;
    if (SCREEN) cn = __fsqrt_ru(csoa[(size_t)ncp_max * 4]) * 1.000001f;  // max W of the candidate cloud (block tail)
    const int lane = tid & 31;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float oqx = kUnscale * qx[r], oqy = kUnscale * qy[r], oqz = kUnscale * qz[r];
        const int c = min(bchunk[r], ncp - G);  // (an empty trailing split re-reads the last chunk; its result loses every merge)
        // Exact difference form on the chunk's G candidates (packed math, same bits as exact_d).  Entries past nc
        // replicate point nc-1, so they can tie with it but sit at higher indices and never win.
        const float nq1 = SCREEN ? 0.5f : 1.0f;  // registers hold -2q (SCREEN) or -q (EXACT); both rescalings are exact
        const float2 nqx = make_float2(nq1 * qx[r], nq1 * qx[r]), nqy = make_float2(nq1 * qy[r], nq1 * qy[r]),
                     nqz = make_float2(nq1 * qz[r], nq1 * qz[r]);
        float bd = 0.0f;
        int bi = c;
#pragma unroll
        for (int k = 0; k < G; k += 8) {
            // per-lane scattered reads of the chunk: 256-bit loads fetch each 32-byte sector exactly once
            const float8 x8 = ldg256(gX + c + k), y8 = ldg256(gY + c + k), z8 = ldg256(gZ + c + k);
#pragma unroll
            for (int h = 0; h < 8; h += 2) {
                const float2 dx = __fadd2_rn(make_float2(x8.v[h], x8.v[h + 1]), nqx);
                const float2 dy = __fadd2_rn(make_float2(y8.v[h], y8.v[h + 1]), nqy);
                const float2 dz = __fadd2_rn(make_float2(z8.v[h], z8.v[h + 1]), nqz);
                const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
                if ((k == 0 && h == 0) || d.x < bd) { bd = d.x; bi = c + k + h; }  // strict '<' in ascending order: lowest index wins
                if (d.y < bd) { bd = d.y; bi = c + k + h + 1; }
            }
        }
        if (SCREEN) {
            // |s + |q|^2 - d_fp32| <= 11.02 u S^2 with S = |q| + max|c| (DESIGN.md "screening bound");
            // a runner-up chunk farther than twice that cannot hold the argmin.  eps = 32 u S^2.
            const float qn = __fsqrt_ru(__fmaf_rn(oqz, oqz, __fmaf_rn(oqy, oqy, oqx * oqx))) * 1.000001f;
            const float S = qn + cn;
            const float eps = __fmaf_rn(S * S, 1.9073486e-6f /* 2^-19 */, 1e-35f);
            const bool ambiguous = !(second[r] > best[r] + eps);
            unsigned todo = __ballot_sync(0xffffffffu, ambiguous);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const float ax = __shfl_sync(0xffffffffu, oqx, src);
                const float ay = __shfl_sync(0xffffffffu, oqy, src);
                const float az = __shfl_sync(0xffffffffu, oqz, src);
                float wd = kInf;
                int wi = 0x7fffffff;
                for (int k = k_lo + lane; k < min(k_hi, nc); k += 32) {
                    const float d = exact_d(gX[k], gY[k], gZ[k], ax, ay, az);
                    if (d < wd || wi == 0x7fffffff) { wd = d; wi = k; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, wd, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
                    if (oi != 0x7fffffff && (wi == 0x7fffffff || od < wd || (od == wd && oi < wi))) { wd = od; wi = oi; }
                }
                if (lane == src) { bd = wd; bi = wi; }
            }
        }
        const int j = (qt * R + r) * kNNThreads + tid;
        if (j < nq) {
            if (j >= nq_v) { bd = 0.0f; bi = 0; }                          // past the valid length of a ragged cloud
            else if (p.nsplit > 1 && k_lo >= k_hi) bd = kInf;              // empty split: loses every merge
            out_d[j] = bd;
            out_i[j] = bi;
        }
    }
}

// Merge of the per-split partial results: splits cover ascending candidate ranges, so a strict '<' in split
// order keeps the lowest index among equal distances -- the same rule as the reference's tile merge (chamfer3D.cu:126).
__global__ void __launch_bounds__(256) merge_splits_kernel(const float *__restrict__ pd, const int *__restrict__ pi, size_t total,
                                                           int nsplit, float *__restrict__ dist, int *__restrict__ idx) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= total) return;
    float d = pd[t];
    int i = pi[t];
    for (int s = 1; s < nsplit; s++) {
        const float ds = pd[s * total + t];
        if (ds < d) { d = ds; i = pi[s * total + t]; }
    }
    dist[t] = d;
    idx[t] = i;
}

// Launch shape: R queries per thread and nsplit candidate splits, chosen so that the grid has several waves of
// CTAs (148 SMs x ~6 resident CTAs) without cutting the candidate range below 512 points per CTA.
struct NNShape { int R; int nsplit; };
NNShape choose_nn_shape(int B, int n1, int n2) {
    auto items = [&](int R) {
        return (long long)B * ((n1 + R * kNNThreads - 1) / (R * kNNThreads) + (n2 + R * kNNThreads - 1) / (R * kNNThreads));
    };
    NNShape sh;
    sh.R = 4;
    sh.nsplit = 1;
    const long long want = 148ll * 6 * 4;
    const int max_split = (n1 < n2 ? n1 : n2) / 512;
    if (items(4) < want) {
        long long need = (want + items(4) - 1) / items(4);
        while (sh.nsplit < need && sh.nsplit * 2 <= max_split && sh.nsplit < 8) sh.nsplit *= 2;
        if (items(4) * sh.nsplit < 148ll * 4) sh.R = 2;
    }
    return sh;
}

// ------------------------------------------------------------------------------------------
// DCD / CD epilogue
// ------------------------------------------------------------------------------------------
constexpr int kDcdThreads = 512;

// torch's pow(tensor, python scalar) special cases (ATen pow_tensor_scalar): 1 -> x, 2 -> x*x,
// 0.5 -> sqrt, 0 -> 1; anything else goes through powf
__device__ __forceinline__ float pow_lambda(float c, float n_lambda) {
    if (n_lambda == 1.0f) return c;
    if (n_lambda == 0.5f) return sqrtf(c);
    if (n_lambda == 2.0f) return c * c;
    if (n_lambda == 0.0f) return 1.0f;
    return powf(c, n_lambda);
}

// one CTA per pair: shared-memory histograms of idx1 (bins = points of cloud 2) and idx2
// len1/len2 (optional): valid point counts per cloud entry; pair b uses len1[b / rep1], len2[b % mod2].  With lengths the
// means run over the valid points only and the DCD fractions are rebuilt from them (non_reg clamps them at 1).
struct DcdLens { const int *len1, *len2; int rep1, mod2, non_reg; };

__global__ void __launch_bounds__(kDcdThreads) dcd_fwd_kernel(const float *__restrict__ dist1, const float *__restrict__ dist2,
                                                              const int *__restrict__ idx1, const int *__restrict__ idx2,
                                                              int n1_max, int n2_max, float alpha, float n_lambda, float frac_12,
                                                              float frac_21, float *__restrict__ loss, float *__restrict__ cd_p,
                                                              float *__restrict__ cd_t, float *__restrict__ ew1,
                                                              float *__restrict__ ew2, const DcdLens lens) {
    extern __shared__ int hist[];  // count1[n2] | count2[n1]
    __shared__ double red[(kDcdThreads / 32) * 6];
    int *count1 = hist, *count2 = hist + n2_max;
    const size_t b = blockIdx.x;
    const float *d1 = dist1 + b * n1_max, *d2 = dist2 + b * n2_max;
    const int *i1 = idx1 + b * n1_max, *i2 = idx2 + b * n2_max;
    const int n1 = lens.len1 ? max(0, min(lens.len1[b / lens.rep1], n1_max)) : n1_max;
    const int n2 = lens.len2 ? max(0, min(lens.len2[b % lens.mod2], n2_max)) : n2_max;
    if (lens.len1 || lens.len2) {
        frac_12 = (float)((double)n2 / (double)max(n1, 1));
        frac_21 = (float)((double)n1 / (double)max(n2, 1));
        if (lens.non_reg) { frac_12 = fmaxf(frac_12, 1.0f); frac_21 = fmaxf(frac_21, 1.0f); }
    }
    if (n1 == 0 || n2 == 0) {  // an empty side: nothing to average (callers mask such pairs out)
        for (int k = threadIdx.x; k < n1_max; k += kDcdThreads) if (ew1) ew1[b * n1_max + k] = 0.0f;
        for (int k = threadIdx.x; k < n2_max; k += kDcdThreads) if (ew2) ew2[b * n2_max + k] = 0.0f;
        if (threadIdx.x == 0) {
            if (loss) loss[b] = 0.0f;
            if (cd_p) cd_p[b] = 0.0f;
            if (cd_t) cd_t[b] = 0.0f;
        }
        return;
    }
    for (int k = threadIdx.x; k < n1_max + n2_max; k += kDcdThreads) hist[k] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < n1; k += kDcdThreads) atomicAdd(&count1[i1[k]], 1);
    for (int k = threadIdx.x; k < n2; k += kDcdThreads) atomicAdd(&count2[i2[k]], 1);
    __syncthreads();

    // per-thread float64 partial sums: [side][term, d, sqrt d]; reduced once (warp shuffles, then a fixed-order
    // pass over the per-warp partials, so the result is deterministic)
    double part[6];
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int n = side ? n2 : n1;
        const int n_stride = side ? n2_max : n1_max;
        const float *d = side ? d2 : d1;
        const int *ix = side ? i2 : i1;
        const int *cnt = side ? count2 : count1;
        const float frac = side ? frac_12 : frac_21;
        float *ew = side ? ew2 : ew1;
        double a_term = 0.0, a_d = 0.0, a_sqrt = 0.0;
        if (ew) for (int k = n + threadIdx.x; k < n_stride; k += kDcdThreads) ew[b * n_stride + k] = 0.0f;
        for (int k = threadIdx.x; k < n; k += kDcdThreads) {
            const float dk = d[k];
            const float c = (float)cnt[ix[k]];
            // model_utils.py:31,35-37: exp(-d*alpha); (count**lambda + 1e-6)**(-1) * frac
            const float e = expf(__fmul_rn(-dk, alpha));
            const float w = __fmul_rn(__fdiv_rn(1.0f, __fadd_rn(pow_lambda(c, n_lambda), 1e-6f)), frac);
            const float ewk = __fmul_rn(e, w);
            if (ew) ew[b * n_stride + k] = ewk;
            a_term += (double)__fsub_rn(1.0f, ewk);
            a_d += (double)dk;
            a_sqrt += (double)sqrtf(dk);
        }
        part[side * 3 + 0] = a_term; part[side * 3 + 1] = a_d; part[side * 3 + 2] = a_sqrt;
    }
#pragma unroll
    for (int v = 0; v < 6; v++)
        for (int o = 16; o > 0; o >>= 1) part[v] += __shfl_xor_sync(0xffffffffu, part[v], o);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int v = 0; v < 6; v++) red[(threadIdx.x >> 5) * 6 + v] = part[v];
    }
    __syncthreads();
    double s_term[2], s_d[2], s_sqrt[2];
    if (threadIdx.x == 0) {
        double tot[6] = {0, 0, 0, 0, 0, 0};
        for (int w = 0; w < kDcdThreads / 32; w++)
            for (int v = 0; v < 6; v++) tot[v] += red[w * 6 + v];
        s_term[0] = tot[0]; s_d[0] = tot[1]; s_sqrt[0] = tot[2];
        s_term[1] = tot[3]; s_d[1] = tot[4]; s_sqrt[1] = tot[5];
    }
    if (threadIdx.x == 0) {
        const float loss1 = (float)(s_term[0] / n1), loss2 = (float)(s_term[1] / n2);
        const float m1 = (float)(s_d[0] / n1), m2 = (float)(s_d[1] / n2);
        const float r1 = (float)(s_sqrt[0] / n1), r2 = (float)(s_sqrt[1] / n2);
        if (loss) loss[b] = (loss1 + loss2) / 2.0f;  // model_utils.py:45
        if (cd_p) cd_p[b] = (r1 + r2) / 2.0f;        // model_utils.py:57
        if (cd_t) cd_t[b] = m1 + m2;                 // model_utils.py:58
    }
}

// ------------------------------------------------------------------------------------------
// backward: chamfer3D.cu:155-174 with the d(out)/d(dist) coefficient built in place
// ------------------------------------------------------------------------------------------
constexpr int kGradThreads = 256;

struct GradParams {
    const float *xyz[2];
    const float *dist[2];
    const int *idx[2];
    const float *ew[2];
    const float *g_dist[2];
    const float *g_loss, *g_cd_p, *g_cd_t;
    float *grad[2];
    int n[2];
    int rep1, mod2;
    float alpha;
    const int *len[2];  // optional valid point counts per cloud-1 / cloud-2 entry
};
__device__ __forceinline__ int valid_len(const int *len, size_t cloud, int n_max) {
    return len ? max(0, min(len[cloud], n_max)) : n_max;
}

__device__ __forceinline__ float point_grad_coeff(const GradParams &p, int side, size_t b, size_t pt, int n_own);

// PHASE 0: own-side terms, plain coalesced stores (every output element is written exactly once
//          when rep1 == 1 / mod2 == B; broadcast clouds fall back to atomics on a zeroed buffer)
// PHASE 1: scatter-side terms, red.global.add.f32 keyed on idx
template <int PHASE, bool SHARED1, bool SHARED2>
__global__ void __launch_bounds__(kGradThreads) grad_kernel(const GradParams p) {
    const int per_pair = p.n[0] + p.n[1];
    const size_t b = blockIdx.y;
    for (int t = blockIdx.x * kGradThreads + threadIdx.x; t < per_pair; t += gridDim.x * kGradThreads) {
        const int side = t >= p.n[0] ? 1 : 0;
        const int j = side ? t - p.n[0] : t;
        const int n_own = side ? p.n[1] : p.n[0], n_oth = side ? p.n[0] : p.n[1];
        const size_t c1 = b / p.rep1, c2 = b % p.mod2;
        const size_t c_own = side ? c2 : c1, c_oth = side ? c1 : c2;
        const size_t pt = b * n_own + j;
        const int v_own = valid_len(side ? p.len[1] : p.len[0], c_own, n_own);
        const int v_oth = valid_len(side ? p.len[0] : p.len[1], c_oth, n_oth);
        if (j >= v_own || v_oth == 0) {  // past the valid length (or nothing to match): zero gradient
            if (PHASE == 0 && !(side ? SHARED2 : SHARED1)) {
                float *dst = (side ? p.grad[1] : p.grad[0]) + (c_own * n_own + j) * 3;
                dst[0] = 0.0f; dst[1] = 0.0f; dst[2] = 0.0f;
            }
            continue;
        }
        // upstream gradient w.r.t. this point's squared NN distance
        const float gd = point_grad_coeff(p, side, b, pt, v_own);
        const int j2 = (side ? p.idx[1] : p.idx[0])[pt];
        const float *a = (side ? p.xyz[1] : p.xyz[0]) + (c_own * n_own + j) * 3;
        const float *o = (side ? p.xyz[0] : p.xyz[1]) + (c_oth * n_oth + j2) * 3;
        const float g = __fmul_rn(gd, 2.0f);  // chamfer3D.cu:166
        const float gx = __fmul_rn(g, __fsub_rn(a[0], o[0]));
        const float gy = __fmul_rn(g, __fsub_rn(a[1], o[1]));
        const float gz = __fmul_rn(g, __fsub_rn(a[2], o[2]));
        if (PHASE == 0) {
            float *dst = (side ? p.grad[1] : p.grad[0]) + (c_own * n_own + j) * 3;
            const bool shared_own = side ? SHARED2 : SHARED1;
            if (shared_own) {
                atomicAdd(dst + 0, gx); atomicAdd(dst + 1, gy); atomicAdd(dst + 2, gz);
            } else {
                dst[0] = gx; dst[1] = gy; dst[2] = gz;
            }
        } else {
            float *dst = (side ? p.grad[0] : p.grad[1]) + (c_oth * n_oth + j2) * 3;
            atomicAdd(dst + 0, -gx); atomicAdd(dst + 1, -gy); atomicAdd(dst + 2, -gz);
        }
    }
}

// One CTA per pair, both clouds' gradients accumulated in shared memory (own-side and scatter-side
// terms alike), then written out once, coalesced: no global atomics, no zero-fill, one launch.
// Used when the pair's (n1 + n2) * 12 bytes fit in shared memory and neither cloud is broadcast.
constexpr int kGradSmemThreads = 512;

__device__ __forceinline__ float point_grad_coeff(const GradParams &p, int side, size_t b, size_t pt, int n_own) {
    const float *g_dist = side ? p.g_dist[1] : p.g_dist[0];
    float gd = g_dist ? g_dist[pt] : 0.0f;
    if (p.g_cd_t) gd += p.g_cd_t[b] / (float)n_own;
    if (p.g_cd_p) gd += p.g_cd_p[b] * 0.5f / (float)n_own * (0.5f / sqrtf((side ? p.dist[1] : p.dist[0])[pt]));
    if (p.g_loss) gd += p.g_loss[b] * 0.5f / (float)n_own * (p.alpha * (side ? p.ew[1] : p.ew[0])[pt]);
    return gd;
}

__global__ void __launch_bounds__(kGradSmemThreads) grad_smem_kernel(const GradParams p) {
    extern __shared__ float acc[];  // grad of cloud 1 [n1*3] | grad of cloud 2 [n2*3]
    const size_t b = blockIdx.x;
    const int n1 = p.n[0], n2 = p.n[1];
    const float *xyz1 = p.xyz[0] + b * n1 * 3, *xyz2 = p.xyz[1] + b * n2 * 3;
    const int v1 = valid_len(p.len[0], b, n1), v2 = valid_len(p.len[1], b, n2);
    // Shared-memory float atomicAdd is a CAS loop on sm_100 (ATOMS.CAST.SPIN), so only the scatter side uses it:
    // phase 0 STORES every point's own-side term (which also initialises the accumulators), phase 1 adds the
    // scatter-side terms atomically.  The per-point term is recomputed in phase 1 (its loads hit L1).
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        for (int t = threadIdx.x; t < n1 + n2; t += kGradSmemThreads) {
            const int side = t >= n1 ? 1 : 0;
            const int j = side ? t - n1 : t;
            const int n_own = side ? n2 : n1;
            float *own = acc + (side ? n1 * 3 : 0) + j * 3;
            if (j >= (side ? v2 : v1) || (side ? v1 : v2) == 0) {
                if (phase == 0) { own[0] = 0.0f; own[1] = 0.0f; own[2] = 0.0f; }
                continue;
            }
            const size_t pt = b * n_own + j;
            const float gd = point_grad_coeff(p, side, b, pt, side ? v2 : v1);
            const int j2 = (side ? p.idx[1] : p.idx[0])[pt];
            const float *a = (side ? xyz2 : xyz1) + j * 3;
            const float *o = (side ? xyz1 : xyz2) + j2 * 3;
            const float g = __fmul_rn(gd, 2.0f);  // chamfer3D.cu:166
            const float gx = __fmul_rn(g, __fsub_rn(a[0], o[0]));
            const float gy = __fmul_rn(g, __fsub_rn(a[1], o[1]));
            const float gz = __fmul_rn(g, __fsub_rn(a[2], o[2]));
            if (phase == 0) {
                own[0] = gx; own[1] = gy; own[2] = gz;
            } else {
                float *oth = acc + (side ? 0 : n1 * 3) + j2 * 3;
                atomicAdd(oth + 0, -gx); atomicAdd(oth + 1, -gy); atomicAdd(oth + 2, -gz);
            }
        }
        if (phase == 0) __syncthreads();
    }
    __syncthreads();
    float *g1 = p.grad[0] + b * n1 * 3, *g2 = p.grad[1] + b * n2 * 3;
    for (int k = threadIdx.x; k < n1 * 3; k += kGradSmemThreads) g1[k] = acc[k];
    for (int k = threadIdx.x; k < n2 * 3; k += kGradSmemThreads) g2[k] = acc[n1 * 3 + k];
}

// ------------------------------------------------------------------------------------------
// k smallest (score, index) per row
// ------------------------------------------------------------------------------------------
constexpr int kTopkThreads = 256;

__device__ __forceinline__ unsigned long long score_key(float s, int i) {
    unsigned u = __float_as_uint(s);
    u = (s != s) ? 0xffffffffu : ((u & 0x80000000u) ? ~u : (u | 0x80000000u));  // total order, NaN last
    return ((unsigned long long)u << 32) | (unsigned)i;
}

// ids == nullptr: rank columns by (score, column) and report column + idx_offset;
// ids != nullptr: rank by (score, ids[column]) and report the id (merge of per-shard lists); ids < 0 are padding.
__global__ void __launch_bounds__(kTopkThreads) topk_kernel(const float *__restrict__ scores, const int *__restrict__ ids, int cols,
                                                            int k, int idx_offset, float *__restrict__ out_scores,
                                                            int *__restrict__ out_idx) {
    __shared__ unsigned long long red[kTopkThreads / 32];
    __shared__ unsigned long long chosen;
    const float *row = scores + (size_t)blockIdx.x * cols;
    const int *row_ids = ids ? ids + (size_t)blockIdx.x * cols : nullptr;
    unsigned long long last = 0ull;
    bool have_last = false;
    for (int it = 0; it < k; it++) {
        unsigned long long m = ~0ull;
        int mc = -1;
        for (int c = threadIdx.x; c < cols; c += kTopkThreads) {
            const int id = row_ids ? row_ids[c] : c;
            if (id < 0) continue;  // padding entry of a short shard
            const unsigned long long key = score_key(row[c], id);
            if ((!have_last || key > last) && key < m) { m = key; mc = c; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
            const int oc = __shfl_xor_sync(0xffffffffu, mc, o);
            if (other < m) { m = other; mc = oc; }
        }
        __shared__ int red_col[kTopkThreads / 32];
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = m; red_col[threadIdx.x >> 5] = mc; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < kTopkThreads / 32; i++)
                if (red[i] < m) { m = red[i]; mc = red_col[i]; }
            chosen = m;
            const size_t o = (size_t)blockIdx.x * k + it;
            if (mc < 0) {  // fewer than k real entries: pad
                out_scores[o] = kInf;
                out_idx[o] = -1;
            } else {
                out_scores[o] = row[mc];
                out_idx[o] = (row_ids ? row_ids[mc] : mc) + idx_offset;
            }
        }
        __syncthreads();
        last = chosen;
        have_last = true;
    }
}

// ------------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------------
int check_pairs(int B, int n1, int n2, int rep1, int mod2) {
    if (B < 0 || n1 < 0 || n2 < 0) return fail_arg(URED_E_SHAPE, "negative size");
    if (rep1 < 1 || mod2 < 1) return fail_arg(URED_E_SHAPE, "rep1 and mod2 must be >= 1");
    if ((long long)B * (long long)(n1 > n2 ? n1 : n2) >= (1ll << 40)) return fail_arg(URED_E_SHAPE, "problem too large");
    return 0;
}
inline int count1_of(int B, int rep1) { return (B + rep1 - 1) / rep1; }
inline int count2_of(int B, int mod2) { return B < mod2 ? B : mod2; }

template <bool SCREEN, int R>
int launch_nn(const NNParams &p, int B, cudaStream_t st) {
    constexpr int NARR = SCREEN ? 4 : 3;
    const size_t smem = (size_t)kStages * NARR * kTile * sizeof(float);
    const long long grid = (long long)B * (p.qtiles[0] + p.qtiles[1]) * p.nsplit;
    if (grid > 0x7fffffffll) return fail_arg(URED_E_SHAPE, "too many work items for one launch");
    nn_kernel<SCREEN, R><<<(unsigned)grid, kNNThreads, smem, st>>>(p);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "nn_kernel launch");
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int ured_abi_version(void) { return URED_ABI_VERSION; }
const char *ured_last_error_string(void) { return g_err; }
unsigned long long ured_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

size_t ured_packed_bytes(int count, int n) {
    if (count <= 0 || n <= 0) return 256;
    return align_up((size_t)count * cloud_stride(pad32(n)) * sizeof(float), 256);
}

int ured_pack_clouds(const float *xyz, int count, int n, const int *len, void *packed, void *stream) {
    if (count < 0 || n < 0) return fail_arg(URED_E_SHAPE, "negative size");
    if (count == 0 || n == 0) return 0;
    if (!xyz || !packed) return fail_arg(URED_E_NULL, "ured_pack_clouds: NULL pointer");
    if ((uintptr_t)packed % 256) return fail_arg(URED_E_WORKSPACE, "packed image must be 256-byte aligned");
    PackedView v = view_packed(packed, n);
    PackSet a;
    a.xyz = xyz; a.soa = (float *)v.soa; a.len = len; a.n_max = n; a.np = v.np;
    pack_kernel<<<count, kPackThreads, 0, (cudaStream_t)stream>>>(a, a, count);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "pack_kernel launch");
}

size_t ured_nn_scratch_bytes(int B, int n1, int n2) {
    if (B <= 0 || n1 <= 0 || n2 <= 0) return 0;
    const NNShape sh = choose_nn_shape(B, n1, n2);
    return sh.nsplit > 1 ? align_up((size_t)sh.nsplit * B * ((size_t)n1 + n2) * 8, 256) : 0;
}

int ured_nn_packed(const float *xyz1, const void *packed1, int n1, const float *xyz2, const void *packed2, int n2, int B,
                   int rep1, int mod2, const int *len1, const int *len2, float *dist1, float *dist2, int *idx1, int *idx2,
                   void *scratch, size_t scratch_bytes, unsigned flags, void *stream) {
    int rc = check_pairs(B, n1, n2, rep1, mod2);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0 || (n1 == 0 && n2 == 0)) return 0;
    const bool skip2 = (flags & URED_FLAG_ONE_DIRECTION) != 0;
    if ((n1 && (!dist1 || !idx1)) || (n2 && !skip2 && (!dist2 || !idx2))) return fail_arg(URED_E_NULL, "ured_nn_packed: NULL output");
    if (n1 == 0 || n2 == 0) {
        // the reference kernel never writes when the opposing cloud is empty: zeros stay zeros
        if (n1) { URED_CUDA(cudaMemsetAsync(dist1, 0, (size_t)B * n1 * 4, st), "memset"); URED_CUDA(cudaMemsetAsync(idx1, 0, (size_t)B * n1 * 4, st), "memset"); }
        if (n2 && !skip2) { URED_CUDA(cudaMemsetAsync(dist2, 0, (size_t)B * n2 * 4, st), "memset"); URED_CUDA(cudaMemsetAsync(idx2, 0, (size_t)B * n2 * 4, st), "memset"); }
        return 0;
    }
    if (!xyz1 || !xyz2 || (!packed1 && !skip2) || !packed2) return fail_arg(URED_E_NULL, "ured_nn_packed: NULL input");
    PackedView v1 = view_packed(packed1, n1), v2 = view_packed(packed2, n2);
    NNParams p;
    p.xyz[0] = xyz1; p.xyz[1] = xyz2;
    p.soa[0] = v1.soa; p.soa[1] = v2.soa;
    p.dist[0] = dist1; p.dist[1] = dist2;
    p.idx[0] = idx1; p.idx[1] = idx2;
    p.n[0] = n1; p.n[1] = n2;
    p.np[0] = v1.np; p.np[1] = v2.np;
    p.rep1 = rep1; p.mod2 = mod2;
    p.len[0] = len1; p.len[1] = len2;
    const bool exact = (flags & URED_FLAG_EXACT_ONLY) != 0;
    const NNShape sh = choose_nn_shape(B, n1, n2);
    const int R = sh.R;
    const bool one_dir = (flags & URED_FLAG_ONE_DIRECTION) != 0;  // only cloud-1 points search cloud 2 (K=1 kNN)
    p.qtiles[0] = (n1 + R * kNNThreads - 1) / (R * kNNThreads);
    p.qtiles[1] = one_dir ? 0 : (n2 + R * kNNThreads - 1) / (R * kNNThreads);
    p.nsplit = sh.nsplit;
    p.B = B;
    const size_t tot1 = (size_t)B * n1, tot2 = (size_t)B * n2;
    if (sh.nsplit > 1) {
        if (!scratch) return fail_arg(URED_E_NULL, "ured_nn_packed: scratch buffer required for this shape (ured_nn_scratch_bytes)");
        if ((uintptr_t)scratch % 256 || scratch_bytes < ured_nn_scratch_bytes(B, n1, n2))
            return fail_arg(URED_E_WORKSPACE, "ured_nn_packed: scratch too small or misaligned");
        // [dist1 parts | dist2 parts | idx1 parts | idx2 parts]
        p.part_dist[0] = (float *)scratch;
        p.part_dist[1] = p.part_dist[0] + sh.nsplit * tot1;
        p.part_idx[0] = (int *)(p.part_dist[1] + sh.nsplit * tot2);
        p.part_idx[1] = p.part_idx[0] + sh.nsplit * tot1;
    } else {
        p.part_dist[0] = p.part_dist[1] = nullptr;
        p.part_idx[0] = p.part_idx[1] = nullptr;
    }
    if (exact) rc = R == 4 ? launch_nn<false, 4>(p, B, st) : launch_nn<false, 2>(p, B, st);
    else rc = R == 4 ? launch_nn<true, 4>(p, B, st) : launch_nn<true, 2>(p, B, st);
    if (rc || sh.nsplit == 1) return rc;
    merge_splits_kernel<<<(unsigned)((tot1 + 255) / 256), 256, 0, st>>>(p.part_dist[0], p.part_idx[0], tot1, sh.nsplit, dist1, idx1);
    URED_COUNT_LAUNCH();
    if (!one_dir) {
        merge_splits_kernel<<<(unsigned)((tot2 + 255) / 256), 256, 0, st>>>(p.part_dist[1], p.part_idx[1], tot2, sh.nsplit, dist2, idx2);
        URED_COUNT_LAUNCH();
    }
    return check_cuda(cudaGetLastError(), "merge_splits_kernel launch");
}

size_t ured_chamfer_workspace_bytes(int B, int n1, int n2) {
    if (B < 0) B = 0;
    return ured_packed_bytes(B, n1) + ured_packed_bytes(B, n2) + ured_nn_scratch_bytes(B, n1, n2);
}

int ured_chamfer_forward(const float *xyz1, const float *xyz2, int B, int n1, int n2, const int *len1, const int *len2,
                         float *dist1, float *dist2, int *idx1, int *idx2, void *workspace, size_t workspace_bytes,
                         unsigned flags, void *stream) {
    int rc = check_pairs(B, n1, n2, 1, B > 0 ? B : 1);
    if (rc) return rc;
    if (B == 0) return 0;
    if (n1 > 0 && n2 > 0) {
        if (!workspace) return fail_arg(URED_E_NULL, "ured_chamfer_forward: NULL workspace");
        if ((uintptr_t)workspace % 256) return fail_arg(URED_E_WORKSPACE, "workspace must be 256-byte aligned");
        if (workspace_bytes < ured_chamfer_workspace_bytes(B, n1, n2)) return fail_arg(URED_E_WORKSPACE, "workspace too small");
        if (!xyz1 || !xyz2) return fail_arg(URED_E_NULL, "ured_chamfer_forward: NULL input");
    }
    void *pk1 = workspace;
    void *pk2 = (char *)workspace + ured_packed_bytes(B, n1);
    if (n1 > 0 && n2 > 0) {  // both sides in one launch
        PackSet a, b2;
        a.xyz = xyz1; a.soa = (float *)pk1; a.len = len1; a.n_max = n1; a.np = pad32(n1);
        b2.xyz = xyz2; b2.soa = (float *)pk2; b2.len = len2; b2.n_max = n2; b2.np = pad32(n2);
        pack_kernel<<<2 * B, kPackThreads, 0, (cudaStream_t)stream>>>(a, b2, B);
        URED_COUNT_LAUNCH();
        rc = check_cuda(cudaGetLastError(), "pack_kernel launch");
        if (rc) return rc;
    }
    void *scratch = (char *)pk2 + ured_packed_bytes(B, n2);
    return ured_nn_packed(xyz1, pk1, n1, xyz2, pk2, n2, B, 1, B, len1, len2, dist1, dist2, idx1, idx2, scratch,
                          ured_nn_scratch_bytes(B, n1, n2), flags, stream);
}

int ured_dcd_forward(const float *dist1, const float *dist2, const int *idx1, const int *idx2, int B, int n1, int n2,
                     int rep1, int mod2, const int *len1, const int *len2, float alpha, float n_lambda, float frac_12,
                     float frac_21, unsigned flags, float *loss, float *cd_p, float *cd_t, float *ew1, float *ew2, void *stream) {
    if (B < 0 || n1 < 0 || n2 < 0) return fail_arg(URED_E_SHAPE, "negative size");
    if (rep1 < 1 || mod2 < 1) return fail_arg(URED_E_SHAPE, "rep1 and mod2 must be >= 1");
    if (B == 0) return 0;
    if (n1 == 0 || n2 == 0) return fail_arg(URED_E_SHAPE, "ured_dcd_forward: empty cloud");
    if (!dist1 || !dist2 || !idx1 || !idx2) return fail_arg(URED_E_NULL, "ured_dcd_forward: NULL input");
    const size_t smem = (size_t)(n1 + n2) * sizeof(int);
    if (smem > 200 * 1024) return fail_arg(URED_E_RANGE, "ured_dcd_forward: n1 + n2 > 51200 points per pair not supported");
    static thread_local bool attr_done = false;
    if (!attr_done) {
        URED_CUDA(cudaFuncSetAttribute(dcd_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "dcd smem attribute");
        attr_done = true;
    }
    DcdLens lens;
    lens.len1 = len1; lens.len2 = len2; lens.rep1 = rep1; lens.mod2 = mod2; lens.non_reg = (flags & URED_FLAG_NON_REG) ? 1 : 0;
    dcd_fwd_kernel<<<B, kDcdThreads, smem, (cudaStream_t)stream>>>(dist1, dist2, idx1, idx2, n1, n2, alpha, n_lambda, frac_12,
                                                                   frac_21, loss, cd_p, cd_t, ew1, ew2, lens);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "dcd_fwd_kernel launch");
}

int ured_dcd_backward(const float *xyz1, const float *xyz2, int B, int n1, int n2, int rep1, int mod2, const int *len1,
                      const int *len2, const float *dist1,
                      const float *dist2, const int *idx1, const int *idx2, const float *ew1, const float *ew2, float alpha,
                      const float *g_loss, const float *g_cd_p, const float *g_cd_t, const float *g_dist1,
                      const float *g_dist2, float *gradxyz1, float *gradxyz2, void *stream) {
    int rc = check_pairs(B, n1, n2, rep1, mod2);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int cnt1 = count1_of(B, rep1), cnt2 = count2_of(B, mod2);
    if (B == 0) return 0;
    if (n1 && !gradxyz1) return fail_arg(URED_E_NULL, "ured_dcd_backward: NULL gradxyz1");
    if (n2 && !gradxyz2) return fail_arg(URED_E_NULL, "ured_dcd_backward: NULL gradxyz2");
    const bool shared1 = rep1 != 1, shared2 = mod2 < B;
    if (n1 == 0 || n2 == 0 || shared1) { if (n1) URED_CUDA(cudaMemsetAsync(gradxyz1, 0, (size_t)cnt1 * n1 * 12, st), "memset"); }
    if (n1 == 0 || n2 == 0 || shared2) { if (n2) URED_CUDA(cudaMemsetAsync(gradxyz2, 0, (size_t)cnt2 * n2 * 12, st), "memset"); }
    if (n1 == 0 || n2 == 0) return 0;
    if (!xyz1 || !xyz2 || !idx1 || !idx2) return fail_arg(URED_E_NULL, "ured_dcd_backward: NULL input");
    if (g_loss && (!ew1 || !ew2)) return fail_arg(URED_E_NULL, "ured_dcd_backward: g_loss needs ew1/ew2");
    if (g_cd_p && (!dist1 || !dist2)) return fail_arg(URED_E_NULL, "ured_dcd_backward: g_cd_p needs dist1/dist2");
    GradParams p;
    p.xyz[0] = xyz1; p.xyz[1] = xyz2;
    p.dist[0] = dist1; p.dist[1] = dist2;
    p.idx[0] = idx1; p.idx[1] = idx2;
    p.ew[0] = ew1; p.ew[1] = ew2;
    p.g_dist[0] = g_dist1; p.g_dist[1] = g_dist2;
    p.g_loss = g_loss; p.g_cd_p = g_cd_p; p.g_cd_t = g_cd_t;
    p.grad[0] = gradxyz1; p.grad[1] = gradxyz2;
    p.n[0] = n1; p.n[1] = n2;
    p.rep1 = rep1; p.mod2 = mod2;
    p.alpha = alpha;
    p.len[0] = len1; p.len[1] = len2;
    const size_t smem_need = (size_t)(n1 + n2) * 3 * sizeof(float);
    if (!shared1 && !shared2 && smem_need <= 96 * 1024) {
        static thread_local bool attr_done = false;
        if (!attr_done) {
            URED_CUDA(cudaFuncSetAttribute(grad_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024), "grad smem attribute");
            attr_done = true;
        }
        grad_smem_kernel<<<B, kGradSmemThreads, smem_need, st>>>(p);
        URED_COUNT_LAUNCH();
        return check_cuda(cudaGetLastError(), "grad_smem_kernel launch");
    }
    int gx = (n1 + n2 + kGradThreads - 1) / kGradThreads;
    if (gx > 64) gx = 64;
    if (B > 65535) return fail_arg(URED_E_SHAPE, "ured_dcd_backward: B > 65535 pairs per call");
    dim3 grid(gx, B);
    if (!shared1 && !shared2) grad_kernel<0, false, false><<<grid, kGradThreads, 0, st>>>(p);
    else if (shared1 && !shared2) grad_kernel<0, true, false><<<grid, kGradThreads, 0, st>>>(p);
    else if (!shared1 && shared2) grad_kernel<0, false, true><<<grid, kGradThreads, 0, st>>>(p);
    else grad_kernel<0, true, true><<<grid, kGradThreads, 0, st>>>(p);
    URED_COUNT_LAUNCH();
    URED_CUDA(cudaGetLastError(), "grad_kernel<own> launch");
    grad_kernel<1, false, false><<<grid, kGradThreads, 0, st>>>(p);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "grad_kernel<scatter> launch");
}

int ured_chamfer_backward(const float *xyz1, const float *xyz2, int B, int n1, int n2, int rep1, int mod2, const int *len1,
                          const int *len2, const float *graddist1, const float *graddist2, const int *idx1, const int *idx2,
                          float *gradxyz1, float *gradxyz2, void *stream) {
    return ured_dcd_backward(xyz1, xyz2, B, n1, n2, rep1, mod2, len1, len2, nullptr, nullptr, idx1, idx2, nullptr, nullptr, 0.0f, nullptr,
                             nullptr, nullptr, graddist1, graddist2, gradxyz1, gradxyz2, stream);
}

int ured_topk_smallest(const float *scores, int rows, int cols, int k, int idx_offset, float *out_scores, int *out_idx,
                       void *stream) {
    if (rows < 0 || cols < 0 || k < 0) return fail_arg(URED_E_SHAPE, "negative size");
    if (k > cols || k > 1024) return fail_arg(URED_E_RANGE, "ured_topk_smallest: need k <= cols and k <= 1024");
    if (rows == 0 || k == 0) return 0;
    if (!scores || !out_scores || !out_idx) return fail_arg(URED_E_NULL, "ured_topk_smallest: NULL pointer");
    topk_kernel<<<rows, kTopkThreads, 0, (cudaStream_t)stream>>>(scores, nullptr, cols, k, idx_offset, out_scores, out_idx);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "topk_kernel launch");
}

int ured_merge_topk(const float *scores, const int *ids, int rows, int cols, int k, float *out_scores, int *out_ids, void *stream) {
    if (rows < 0 || cols < 0 || k < 0) return fail_arg(URED_E_SHAPE, "negative size");
    if (k > 1024) return fail_arg(URED_E_RANGE, "ured_merge_topk: k <= 1024");
    if (rows == 0 || k == 0) return 0;
    if (!scores || !ids || !out_scores || !out_ids) return fail_arg(URED_E_NULL, "ured_merge_topk: NULL pointer");
    topk_kernel<<<rows, kTopkThreads, 0, (cudaStream_t)stream>>>(scores, ids, cols, k, 0, out_scores, out_ids);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "topk_kernel(merge) launch");
}

}  // extern "C"
