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
//   nn_tc_kernel  (default) both directions of the nearest-neighbour search in ONE launch, the screening scores
//                 |c|^2 - 2 q.c computed by the TENSOR CORES: exact bf16x3 operand splits (27 products per pair,
//                 K = 32), tcgen05.mma kind::f16 into fp32 accumulators in TMEM; persistent warp-specialised CTAs
//                 (TMEM readers keeping 32-candidate chunk minima, MMA issuers, resolver warps doing the exact
//                 difference-form re-check of the winning chunk).  Output bits identical to the kernels below.
//   nn_kernel     the same search on the FP32 pipes (URED_FLAG_FP32_SCREEN / URED_FLAG_EXACT_ONLY).  Each CTA owns
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
#include <stdlib.h>
#include <math.h>

#include <atomic>

#include <cooperative_groups.h>

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

// dynamic shared memory above this needs the opt-in attribute: the 48 KB default limit covers static + dynamic shared
// memory together, and the kernels here declare up to ~1 KB statically
constexpr size_t kSmemOptIn = 40 * 1024;

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
    // candidate splitting: work items (pair, direction, query tile) [0, full_items) are scanned by one CTA each; the
    // split_items items after them are cut into nsplit CTAs, each scanning a contiguous candidate range and writing its
    // exact partial (d, idx) to part_*[split][split item][query of the tile].  Splitting only the LAST items of a launch
    // (longest jobs first) fills the final, partly empty wave with short jobs.
    int nsplit;
    int full_items, split_items;
    float *part_dist;
    int *part_idx;
    const int *len[2];  // optional valid point counts per cloud-1 / cloud-2 entry (ragged batches)
};

// the reference's pair arithmetic (chamfer3D.cu:32-35 as compiled by nvcc 12.9 for sm_100a):
// differences are candidate - query, d = fma(dz,dz, fma(dx,dx, dy*dy))
__device__ __forceinline__ float exact_d(float cx, float cy, float cz, float qx, float qy, float qz) {
    float dx = __fsub_rn(cx, qx), dy = __fsub_rn(cy, qy), dz = __fsub_rn(cz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Launch variants: T threads per CTA, R queries per thread in registers, MINB resident CTAs per SM (register cap).
// Few fat warps win on this loop: the FFMA2 stream needs no latency hiding beyond its own ILP, a larger R divides the
// LDS and chunk-bookkeeping instructions per pair, and a looser register cap lets ptxas keep the half-major order.
template <bool SCREEN, int R, int T, int MINB, bool PF = false>
__global__ void __launch_bounds__(T, MINB) nn_kernel(const NNParams p) {
    constexpr int NARR = SCREEN ? 4 : 3;
    constexpr int G = kChunk;
    constexpr int QT = R * T;  // queries per CTA
    extern __shared__ __align__(128) float stage_mem[];  // kStages * NARR * kTile floats
    __shared__ __align__(8) uint64_t full_bar[kStages];

    const int tid = threadIdx.x;
    // ---- work item: (pair b, direction, query tile) --------------------------------------
    int item = blockIdx.x, sp = 0, nsplit = 1;   // sp: candidate split handled by this CTA
    if ((int)blockIdx.x >= p.full_items) {
        const int r0 = blockIdx.x - p.full_items;
        item = p.full_items + r0 / p.nsplit;
        sp = r0 % p.nsplit;
        nsplit = p.nsplit;
    }
    const int per_pair = p.qtiles[0] + p.qtiles[1];
    const int b = item / per_pair;
    const int rem = item - b * per_pair;
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
    // queries come from the query cloud's own packed image when there is one (three coalesced loads; the raw cloud need
    // not exist any more), otherwise from the raw [n, 3] rows (one-direction calls pack only the candidate side)
    const int nqp_max = dir ? p.np[1] : p.np[0];
    const float *__restrict__ qsoa = (dir ? p.soa[1] : p.soa[0]);
    if (qsoa) qsoa += (size_t)cq * ((size_t)nqp_max * 4 + kPackTail);
    const float *__restrict__ csoa = (dir ? p.soa[0] : p.soa[1]) + (size_t)cc * ((size_t)ncp_max * 4 + kPackTail);

    // unsplit items write the final arrays at the query's index j; split items write partials at the query's slot in the tile
    const bool split = nsplit > 1;
    const size_t part_base = ((size_t)sp * p.split_items + (item - p.full_items)) * QT;
    float *out_d = split ? p.part_dist + part_base : (dir ? p.dist[1] : p.dist[0]) + (size_t)b * nq;
    int *out_i = split ? p.part_idx + part_base : (dir ? p.idx[1] : p.idx[0]) + (size_t)b * nq;
    const int out_shift = split ? qt * QT : 0;   // out index = j - out_shift
    if (qt * QT >= nq_v || nc == 0) {
        // nothing to search: the reference kernel leaves its zero-filled outputs untouched in this case
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int j = (qt * R + r) * T + tid;
            if (j < nq) { out_d[j - out_shift] = 0.0f; out_i[j - out_shift] = 0; }
        }
        return;
    }

    // candidate range of this split, in whole 32-candidate chunks
    const int chunks_per_split = (ncp / kChunk + nsplit - 1) / nsplit;
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
        int j = (qt * R + r) * T + tid;
        j = j < nq_v ? j : nq_v - 1;
        if (qsoa) { qx[r] = kScale * qsoa[j]; qy[r] = kScale * qsoa[nqp_max + j]; qz[r] = kScale * qsoa[2 * nqp_max + j]; }
        else { qx[r] = kScale * qxyz[j * 3 + 0]; qy[r] = kScale * qxyz[j * 3 + 1]; qz[r] = kScale * qxyz[j * 3 + 2]; }
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

        // PF: the next group of four candidates is loaded (shared -> registers) while the current one is evaluated, across
        // chunk boundaries too -- for launches that leave an SM with one or two CTAs, where no other warp hides the LDS latency
        float4 px = make_float4(0.f, 0.f, 0.f, 0.f), py = px, pz = px, pw = px;
        if (PF) {
            px = *reinterpret_cast<const float4 *>(X); py = *reinterpret_cast<const float4 *>(Y);
            pz = *reinterpret_cast<const float4 *>(Z); pw = *reinterpret_cast<const float4 *>(W);
        }
        for (int c0 = 0; c0 < tk; c0 += G) {
            float cm[R];
#pragma unroll
            for (int r = 0; r < R; r++) cm[r] = kInf;
#pragma unroll
            for (int k = 0; k < G; k += 4) {
                float4 x4, y4, z4, w4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (PF) {
                    x4 = px; y4 = py; z4 = pz; w4 = pw;
                    int nk = c0 + k + 4;
                    nk = nk < tk ? nk : 0;   // (past the tile: a harmless re-read of its first group)
                    px = *reinterpret_cast<const float4 *>(X + nk); py = *reinterpret_cast<const float4 *>(Y + nk);
                    pz = *reinterpret_cast<const float4 *>(Z + nk); pw = *reinterpret_cast<const float4 *>(W + nk);
                } else {
                    x4 = *reinterpret_cast<const float4 *>(X + c0 + k);
                    y4 = *reinterpret_cast<const float4 *>(Y + c0 + k);
                    z4 = *reinterpret_cast<const float4 *>(Z + c0 + k);
                    if (SCREEN) w4 = *reinterpret_cast<const float4 *>(W + c0 + k);
                }
                // half-major order: one candidate pair against all R queries, layer by layer, so that the 64-bit
                // candidate operand stays in the operand-reuse slot and only q (1 word) + t (2 words) are read per FFMA2
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float2 xp = h ? make_float2(x4.z, x4.w) : make_float2(x4.x, x4.y);
                    const float2 yp = h ? make_float2(y4.z, y4.w) : make_float2(y4.x, y4.y);
                    const float2 zp = h ? make_float2(z4.z, z4.w) : make_float2(z4.x, z4.y);
                    float2 tt[R];
                    if (SCREEN) {
                        const float2 wp = h ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
#pragma unroll
                        for (int r = 0; r < R; r++) tt[r] = __ffma2_rn(zp, make_float2(qz[r], qz[r]), wp);
#pragma unroll
                        for (int r = 0; r < R; r++) tt[r] = __ffma2_rn(yp, make_float2(qy[r], qy[r]), tt[r]);
#pragma unroll
                        for (int r = 0; r < R; r++) tt[r] = __ffma2_rn(xp, make_float2(qx[r], qx[r]), tt[r]);
                    } else {
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const float2 dy = __fadd2_rn(yp, make_float2(qy[r], qy[r]));
                            tt[r] = __fmul2_rn(dy, dy);
                        }
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const float2 dx = __fadd2_rn(xp, make_float2(qx[r], qx[r]));
                            tt[r] = __ffma2_rn(dx, dx, tt[r]);
                        }
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const float2 dz = __fadd2_rn(zp, make_float2(qz[r], qz[r]));
                            tt[r] = __ffma2_rn(dz, dz, tt[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; r++) cm[r] = fminf(fminf(cm[r], tt[r].x), tt[r].y);
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
    // Tiles t >= ntiles - kStages are still resident in shared memory (nothing was issued over them), i.e. candidates
    // from k_res on; for clouds of up to kStages * kTile = 2048 candidates per split that is the whole range.
    const int k_res = k_lo + max(0, ntiles - kStages) * kTile;
    const float *__restrict__ gX = csoa, *__restrict__ gY = csoa + ncp_max, *__restrict__ gZ = csoa + 2 * (size_t)ncp_max;
    float cn = 0.0f;
    if (SCREEN) cn = __fsqrt_ru(csoa[(size_t)ncp_max * 4]) * 1.000001f;  // max W of the candidate cloud (block tail)
    const int lane = tid & 31;
    const int rot = lane & 7;  // 16-byte piece this lane starts its chunk at (see below)
    // U queries per iteration (unrolled: their dependent compare chains interleave), always elements 0..U-1 of the register
    // arrays, which are then shifted down by U: for R > U the loop stays rolled (the re-check is ~400 instructions per
    // query; unrolled 16 times it would not fit the instruction cache) without ever indexing a register array dynamically.
    constexpr int U = R < 4 ? R : 4;
#pragma unroll 1
    for (int it = 0; it < R; it += U) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const float oqx = kUnscale * qx[u], oqy = kUnscale * qy[u], oqz = kUnscale * qz[u];
        const int c = min(bchunk[u], ncp - G);  // (an empty trailing split re-reads the last chunk; its result loses every merge)
        // Exact difference form on the chunk's G candidates (packed math, same bits as exact_d).  Entries past nc
        // replicate point nc-1, so they can tie with it but sit at higher indices and never win.
        const float nq1 = SCREEN ? 0.5f : 1.0f;  // registers hold -2q (SCREEN) or -q (EXACT); both rescalings are exact
        const float2 nqx = make_float2(nq1 * qx[u], nq1 * qx[u]), nqy = make_float2(nq1 * qy[u], nq1 * qy[u]),
                     nqz = make_float2(nq1 * qz[u], nq1 * qz[u]);
        // d >= +0 for every finite pair, so the running minimum is kept on the float's bit pattern (same order)
        unsigned bdu = 0u;
        int bi = c;
        if (c >= k_res) {
            // From the resident tile.  Every lane reads its own chunk, 16 bytes (one bank group) at a time.  Chunks start on
            // 128-byte boundaries, so in natural order the 8 lanes of a quarter-warp phase would all hit the same bank
            // group; lane l therefore starts at piece (l & 7) and wraps around, which makes every phase conflict-free.
            // Lowest index on ties in that rotated order: the pieces after the wrap hold LOWER indices than anything seen
            // before it, so at the wrap the running minimum is bumped by one ulp -- a tie from there on wins, and from
            // then on strict '<' keeps the first (lowest) of the remaining ties.  If nothing after the wrap won, the
            // bump is undone.
            const int tt_i = (c - k_lo) / kTile;
            const float *sX = stage_mem + ((tt_i % kStages) * NARR + 0) * kTile + (c - k_lo - tt_i * kTile);
            const float *sY = sX + kTile, *sZ = sX + 2 * kTile;
            unsigned hdu = 0u;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int piece = (k + rot) & 7;
                if (k > 0 && piece == 0) { hdu = bdu; bdu += 1u; }
                const float4 x4 = *reinterpret_cast<const float4 *>(sX + piece * 4);
                const float4 y4 = *reinterpret_cast<const float4 *>(sY + piece * 4);
                const float4 z4 = *reinterpret_cast<const float4 *>(sZ + piece * 4);
                const float2 dx0 = __fadd2_rn(make_float2(x4.x, x4.y), nqx), dx1 = __fadd2_rn(make_float2(x4.z, x4.w), nqx);
                const float2 dy0 = __fadd2_rn(make_float2(y4.x, y4.y), nqy), dy1 = __fadd2_rn(make_float2(y4.z, y4.w), nqy);
                const float2 dz0 = __fadd2_rn(make_float2(z4.x, z4.y), nqz), dz1 = __fadd2_rn(make_float2(z4.z, z4.w), nqz);
                const float2 d0 = __ffma2_rn(dz0, dz0, __ffma2_rn(dx0, dx0, __fmul2_rn(dy0, dy0)));
                const float2 d1 = __ffma2_rn(dz1, dz1, __ffma2_rn(dx1, dx1, __fmul2_rn(dy1, dy1)));
                const int i0 = c + piece * 4;
                const unsigned u0 = __float_as_uint(d0.x), u1 = __float_as_uint(d0.y), u2 = __float_as_uint(d1.x), u3 = __float_as_uint(d1.y);
                if (k == 0 || u0 < bdu) { bdu = u0; bi = i0; }
                if (u1 < bdu) { bdu = u1; bi = i0 + 1; }
                if (u2 < bdu) { bdu = u2; bi = i0 + 2; }
                if (u3 < bdu) { bdu = u3; bi = i0 + 3; }
            }
            if (rot != 0 && bi >= c + rot * 4) bdu = hdu;  // the winner predates the wrap: undo the bump
        } else {
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
                    const unsigned u0 = __float_as_uint(d.x), u1 = __float_as_uint(d.y);
                    if ((k == 0 && h == 0) || u0 < bdu) { bdu = u0; bi = c + k + h; }  // strict '<' in ascending order: lowest index wins
                    if (u1 < bdu) { bdu = u1; bi = c + k + h + 1; }
                }
            }
        }
        float bd = __uint_as_float(bdu);
        if (SCREEN) {
            // |s + |q|^2 - d_fp32| <= 11.02 u S^2 with S = |q| + max|c| (DESIGN.md "screening bound");
            // a runner-up chunk farther than twice that cannot hold the argmin.  eps = 32 u S^2.
            const float qn = __fsqrt_ru(__fmaf_rn(oqz, oqz, __fmaf_rn(oqy, oqy, oqx * oqx))) * 1.000001f;
            const float S = qn + cn;
            const float eps = __fmaf_rn(S * S, 1.9073486e-6f /* 2^-19 */, 1e-35f);
            const bool ambiguous = !(second[u] > best[u] + eps);
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
        const int j = (qt * R + it + u) * T + tid;
        if (j < nq) {
            if (j >= nq_v) { bd = 0.0f; bi = 0; }                          // past the valid length of a ragged cloud
            else if (split && k_lo >= k_hi) bd = kInf;                     // empty split: loses every merge
            out_d[j - out_shift] = bd;
            out_i[j - out_shift] = bi;
        }
      }
      if (R > U) {
#pragma unroll
        for (int r = 0; r + U < R; r++) {
            qx[r] = qx[r + U]; qy[r] = qy[r + U]; qz[r] = qz[r + U];
            best[r] = best[r + U]; second[r] = second[r + U]; bchunk[r] = bchunk[r + U];
        }
      }
    }
}

// Merge of the per-split partial results: splits cover ascending candidate ranges, so a strict '<' in split
// order keeps the lowest index among equal distances -- the same rule as the reference's tile merge (chamfer3D.cu:126).
// One thread per (split item, query slot of its tile); both directions in one launch.
struct MergeParams {
    const float *pd;
    const int *pi;
    float *dist[2];
    int *idx[2];
    int n[2], qtiles[2];
    int nsplit, full_items, split_items, QT;
};
__global__ void __launch_bounds__(256) merge_splits_kernel(const MergeParams m) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t per_split = (size_t)m.split_items * m.QT;
    if (t >= per_split) return;
    const int titem = (int)(t / m.QT), l = (int)(t - (size_t)titem * m.QT);
    const int item = m.full_items + titem;
    const int per_pair = m.qtiles[0] + m.qtiles[1];
    const int b = item / per_pair, rem = item - b * per_pair;
    const int dir = rem >= m.qtiles[0] ? 1 : 0;
    const int qt = dir ? rem - m.qtiles[0] : rem;
    const int nq = dir ? m.n[1] : m.n[0];
    const int j = qt * m.QT + l;
    if (j >= nq) return;
    float d = m.pd[t];
    int i = m.pi[t];
    for (int s = 1; s < m.nsplit; s++) {
        const float ds = m.pd[s * per_split + t];
        if (ds < d) { d = ds; i = m.pi[s * per_split + t]; }
    }
    (dir ? m.dist[1] : m.dist[0])[(size_t)b * nq + j] = d;
    (dir ? m.idx[1] : m.idx[0])[(size_t)b * nq + j] = i;
}

// ------------------------------------------------------------------------------------------
// nearest-neighbour kernel, screening pass on the tensor cores (tcgen05 + TMEM)
// ------------------------------------------------------------------------------------------
// The screening scores s = W_c - 2 q.c of nn_kernel<SCREEN> are a [queries x K] . [candidates x K]^T product.  With K = 3 the
// tensor cores are no use -- unless the operands are made exact in their input type: every fp32 coordinate is split into three
// bf16 pieces (x = x1 + x2 + x3 exactly: 3 x 8 significant bits), and the product (-2 q_d) c_d is expanded into the eight piece
// products a_i b_j with i + j <= 5 (the ninth, a_3 b_3, is below 2^-32 of the product); W_c enters as its three pieces times
// 1.0.  That is 27 exact bf16 x bf16 products per pair, K padded to 32, accumulated in fp32 by tcgen05.mma kind::f16
// (M = 128 queries, N = 256 candidates, two K = 16 instructions per tile).  Measured on the B200 (tools/microbench/tc_probe.cu,
// profiles/r02_tc_probe.txt): |s_tc - s_exact| <= 2^-22.6 S^2 over eight input distributions (unit cube, offsets up to 1000,
// scales 1e-3 .. 1e6, mixed scales).  eps = 32 u S^2 tolerates 7.99 u S^2 = 2^-21.0 S^2 there (W's rounding 3 u and the
// reference distance's 5.01 u are the rest of the budget, DESIGN.md "screening bound"), so the SAME eps decides ambiguity and
// everything downstream -- the exact difference-form re-check of the winning 32-candidate chunk, the warp-cooperative exact
// rescan of ambiguous queries, the lowest-index tie rule -- is nn_kernel's.  Output bits are identical.
//
// One persistent CTA per SM.  Work item = (pair, direction, group of <= 16 query tiles); per item the candidate cloud's operand
// image (64 B per point, canonical K-major no-swizzle core-matrix layout) is built ONCE in shared memory and stays resident
// with the fp32 X|Y|Z arrays (TMA bulk copy) for the re-check; clouds of more than 2048 candidates are scanned range by range.
// Roles (default configuration, 448 threads):
//   warps 0-3, 4-7  two sets of TMEM readers.  A thread owns one query of the tile (TMEM lane = query = accumulator row); the
//                   sets take the [128 x 256] accumulator tiles in turn, read them 32 columns (= one candidate chunk) at a
//                   time with tcgen05.ld and keep the chunk minimum / best / runner-up exactly like nn_kernel;
//   warps 8, 9      one lane each issues the MMAs of its set's tiles into the set's accumulator (2 x 256 TMEM columns in all);
//   warps 10-13     resolvers: merge the two sets' partial results, resolve the winning chunk exactly, write the outputs and
//                   build the operand rows of the query tile two ahead.
// Every hand-over is an mbarrier (tcgen05.commit -> full, reader arrivals -> empty, operand rows, partial results).
constexpr int kTcM = 128, kTcK = 32;
constexpr int kTcDefaultSets = 2;                                  // query-warp sets of the default configuration
constexpr int kTcLaunchThreads = 32 * (5 * kTcDefaultSets + 4);    // query warps + one MMA warp per set + 4 resolver warps
constexpr int kTcMaxC = 2048;                       // resident candidates per item
constexpr uint32_t kTcLBO = 128, kTcSBO = 512;      // bytes between the K chunks of 8 rows / between groups of 8 rows
constexpr size_t kTcSmemB = (size_t)(kTcMaxC + 128) * kTcK * 2;    // 128 KB (+ tile padding) operand image of the candidates
constexpr size_t kTcSmemC = (size_t)3 * kTcMaxC * sizeof(float);   // 24 KB X|Y|Z
constexpr size_t kTcSmemA = (size_t)2 * kTcM * kTcK * 2;           // 16 KB, two query tiles
constexpr int kTcMaxGroup = 16;                     // query tiles per work item at most (their running results live in shared memory)
constexpr size_t kTcSmemR = (size_t)kTcMaxGroup * kTcM * 8;        // 16 KB running (distance, index) per query of the group
constexpr size_t kTcSmem = kTcSmemB + kTcSmemC + kTcSmemA + kTcSmemR;

struct TCParams {
    NNParams nn;      // clouds, images, outputs, sizes, lengths (the split fields are unused)
    int groups[2];    // query groups per pair for direction 0 / 1
    int gtiles;       // query tiles (128 queries) per group
    int items;        // B * (groups[0] + groups[1])
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// a wait that cannot hang the GPU: a protocol bug traps (the launch fails) after ten seconds instead of spinning forever
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (unsigned spins = 0; !done; spins++) {
        if (SLEEP_NS > 0 && spins > 0) __nanosleep(SLEEP_NS);   // leave the issue slots to the TMEM readers
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && (spins & 0xfffu) == 0xfffu) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    // start address, leading (K chunk) and stride (8-row group) byte offsets in 16-byte units; version 1 (sm_100); no swizzle
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)(kTcLBO >> 4) << 16) | ((uint64_t)(kTcSBO >> 4) << 32) | ((uint64_t)1 << 46);
}
// x = p1 + p2 + p3 exactly (three bf16 values, returned as their 16-bit patterns)
__device__ __forceinline__ void bf16_split3(float x, uint32_t &p1, uint32_t &p2, uint32_t &p3) {
    uint32_t h;
    asm("{\n\t.reg .b16 t;\n\tcvt.rn.bf16.f32 t, %1;\n\tmov.b32 %0, {t, t};\n\t}" : "=r"(h) : "f"(x));
    p1 = h & 0xffffu;
    const float r1 = __fsub_rn(x, __uint_as_float(p1 << 16));
    asm("{\n\t.reg .b16 t;\n\tcvt.rn.bf16.f32 t, %1;\n\tmov.b32 %0, {t, t};\n\t}" : "=r"(h) : "f"(r1));
    p2 = h & 0xffffu;
    const float r2 = __fsub_rn(r1, __uint_as_float(p2 << 16));
    asm("{\n\t.reg .b16 t;\n\tcvt.rn.bf16.f32 t, %1;\n\tmov.b32 %0, {t, t};\n\t}" : "=r"(h) : "f"(r2));
    p3 = h & 0xffffu;
}
// one operand row (32 bf16 = four 16-byte K chunks) of the canonical layout: row r of a tile starting at `base`
__device__ __forceinline__ unsigned char *tc_row(unsigned char *base, int r) { return base + (size_t)(r >> 3) * kTcSBO + (size_t)(r & 7) * 16; }
// candidate row: per coordinate [b1 b2 b1 b3 b2 b1 b3 b2], then [w1 w2 w3 0 0 0 0 0]
__device__ __forceinline__ void tc_write_candidate(unsigned char *row, float x, float y, float z, float w) {
    const float c[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        uint32_t b1, b2, b3;
        bf16_split3(c[d], b1, b2, b3);
        *reinterpret_cast<uint4 *>(row + d * kTcLBO) = make_uint4(b1 | (b2 << 16), b1 | (b3 << 16), b2 | (b1 << 16), b3 | (b2 << 16));
    }
    uint32_t w1, w2, w3;
    bf16_split3(w, w1, w2, w3);
    *reinterpret_cast<uint4 *>(row + 3 * kTcLBO) = make_uint4(w1 | (w2 << 16), w3, 0u, 0u);
}
// query row: per coordinate the pieces of -2 q_d as [a1 a1 a2 a1 a2 a3 a2 a3], then [1 1 1 0 0 0 0 0]
__device__ __forceinline__ void tc_write_query(unsigned char *row, float x, float y, float z) {
    const float q[3] = {-2.0f * x, -2.0f * y, -2.0f * z};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        uint32_t a1, a2, a3;
        bf16_split3(q[d], a1, a2, a3);
        *reinterpret_cast<uint4 *>(row + d * kTcLBO) = make_uint4(a1 | (a1 << 16), a2 | (a1 << 16), a2 | (a3 << 16), a2 | (a3 << 16));
    }
    *reinterpret_cast<uint4 *>(row + 3 * kTcLBO) = make_uint4(0x3f803f80u, 0x3f80u, 0u, 0u);
}
__device__ __forceinline__ float min32(const uint32_t (&v)[32]) {
    float t[11];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = fminf(fminf(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1])), __uint_as_float(v[3 * i + 2]));
    t[10] = fminf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float u0 = fminf(fminf(t[0], t[1]), t[2]), u1 = fminf(fminf(t[3], t[4]), t[5]), u2 = fminf(fminf(t[6], t[7]), t[8]);
    return fminf(fminf(fminf(u0, u1), u2), fminf(t[9], t[10]));
}
#define URED_TMEM_LD32(v, taddr)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, " \
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                 \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),       \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),            \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),           \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                         \
                 : "r"(taddr)                                                                                                        \
                 : "memory")

#ifdef URED_TC_PROFILE
__device__ long long g_tc_trace[4][24];
__device__ long long g_tc_prof[16];   // [0] items [1] build [2] wait full [3] tmem read + reduce [4] merge + re-check [5] total  (thread 0 of CTA 0)
#define TC_T(x) const long long x = clock64()
#define TC_ADD(i, a, b) do { if (tid == 0) pacc[i] += (b) - (a); } while (0)            // (registers: a global update here would stall the thread)
#define TC_ADDM(i, a, b) do { if (is_mma_thread && mma_id == 0) pacc[i] += (b) - (a); } while (0)
#define TC_PROF_DECL long long pacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; int trace_n = 0, fine_n = 0
// timeline of the SECOND work item of CTA 0: role r (0/1 reader sets, 2 MMA warp 0, 3 resolvers) stamps the end of every query tile
#define TC_TRACE_RESET trace_n = 0; fine_n = 0
#define TC_FINE(v) do { if (blockIdx.x == 0 && tid == 0 && item == (int)(blockIdx.x + gridDim.x) && qt == qt_lo + 3 && fine_n < 24) g_tc_trace[1][fine_n++] = (v); } while (0)
#define TC_TRACE(role) do { if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0 && item == (int)(blockIdx.x + gridDim.x) && trace_n < 24) g_tc_trace[role][trace_n++] = clock64(); } while (0)
#define TC_PROF_FLUSH do { if (blockIdx.x == 0 && (tid == 0 || (is_mma_thread && mma_id == 0))) for (int i_ = 0; i_ < 10; i_++) atomicAdd((unsigned long long *)&g_tc_prof[i_], (unsigned long long)pacc[i_]); } while (0)
#else
#define TC_PROF_DECL
#define TC_FINE(v)
#define TC_TRACE_RESET
#define TC_TRACE(role)
#define TC_PROF_FLUSH
#define TC_ADDM(i, a, b)
#define TC_T(x)
#define TC_ADD(i, a, b)
#endif
// NSETS sets of four query warps take the accumulator tiles in turn (tile counter % NSETS); NACC accumulators of N columns each
// form the ring the MMA warp fills (NACC * N <= 512 TMEM columns).
// SPLIT: every set reads every tile, set s the columns [s N / NSETS, (s + 1) N / NSETS) -- the tensor core then refills one
// accumulator while ALL query warps drain the other, instead of each set waiting out the refill of its own.
// One MMA-issuing warp per set (non-SPLIT): issuing a tile's two MMAs and the commit costs the issuing thread ~300 cycles
// whatever the tile width (measured), so one warp feeding every accumulator in turn paces the whole CTA.
template <int NSETS, int NACC, int N, bool SPLIT>
__global__ void __launch_bounds__(32 * (4 * NSETS + (SPLIT ? 1 : NSETS) + 4), 1) nn_tc_kernel(const TCParams tp) {
    constexpr int kMmaWarps = SPLIT ? 1 : NSETS;
    constexpr int kChunksPerSet = SPLIT ? N / 32 / NSETS : N / 32;
    static_assert(!SPLIT || (N / 32) % NSETS == 0, "SPLIT: the tile's chunks must divide among the sets");
    constexpr int kTcN = N, kTcAcc = NACC, kTcThreads = 32 * (4 * NSETS + kMmaWarps + 4);
    constexpr int kQueryWarps = 4 * NSETS;
    static_assert(NACC * N <= 512 && N % 32 == 0 && N <= 256, "accumulator ring must fit the 512 TMEM columns");
    const NNParams &p = tp.nn;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *Bimg = tc_smem;
    float *Cx = reinterpret_cast<float *>(tc_smem + kTcSmemB), *Cy = Cx + kTcMaxC, *Cz = Cy + kTcMaxC;
    unsigned char *Abuf = tc_smem + kTcSmemB + kTcSmemC;
    float *run_d = reinterpret_cast<float *>(tc_smem + kTcSmemB + kTcSmemC + kTcSmemA);   // clouds of more than kTcMaxC candidates:
    int *run_i = reinterpret_cast<int *>(run_d + kTcMaxGroup * kTcM);                     // best so far over the candidate ranges
    __shared__ __align__(8) uint64_t full_bar[kTcAcc], empty_bar[kTcAcc], a_bar[2], r_bar[2], done_bar[2], c_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float x_best[2][NSETS][kTcM], x_second[2][NSETS][kTcM];   // [query-tile parity][set][query]: the sets' partial results
    __shared__ int x_chunk[2][NSETS][kTcM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // warps [0, 4 NSETS): query warps, set = warp / 4 (warp w reads TMEM lanes 32 (w % 4) ..); then the MMA warp; then 4 resolver warps
    const bool is_query_thread = warp < kQueryWarps, is_mma_warp = warp >= kQueryWarps && warp < kQueryWarps + kMmaWarps;
    const bool is_mma_thread = is_mma_warp && lane == 0, is_resolver = warp >= kQueryWarps + kMmaWarps;
    const int set = warp >> 2;
    const int mma_id = warp - kQueryWarps;   // MMA warp m issues the tiles of set m
    const int t = is_resolver ? tid - 32 * (kQueryWarps + kMmaWarps) : (tid & (kTcM - 1));   // the query of the tile this thread looks after

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kTcAcc; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], SPLIT ? NSETS * kTcM : kTcM); }
#pragma unroll
        for (int i = 0; i < 2; i++) { mbar_init(&a_bar[i], kTcM); mbar_init(&r_bar[i], NSETS * kTcM); mbar_init(&done_bar[i], kTcM); }
        mbar_init(&c_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // kind::f16 instruction descriptor: fp32 accumulator, bf16 x bf16, both operands K-major, N / 8 at bit 17, M / 16 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

    TC_PROF_DECL;
    uint32_t tile_ctr = 0, qt_ctr = 0, item_ctr = 0;   // accumulator tiles / query tiles / (item, range) passes this CTA has gone through (all threads agree)
    const int per_pair = tp.groups[0] + tp.groups[1];
    for (int item = blockIdx.x; item < tp.items; item += gridDim.x) {
        const int b = item / per_pair;
        const int rem = item - b * per_pair;
        const int dir = rem >= tp.groups[0] ? 1 : 0;
        const int grp = dir ? rem - tp.groups[0] : rem;
        const int c1 = b / p.rep1, c2 = b % p.mod2;
        const int cq = dir ? c2 : c1, cc = dir ? c1 : c2;
        const int nq = dir ? p.n[1] : p.n[0];
        const int ncp_max = dir ? p.np[0] : p.np[1];
        const int *len_q = dir ? p.len[1] : p.len[0], *len_c = dir ? p.len[0] : p.len[1];
        const int nq_v = len_q ? max(0, min(len_q[cq], nq)) : nq;
        const int nc_all = len_c ? max(0, min(len_c[cc], dir ? p.n[0] : p.n[1])) : (dir ? p.n[0] : p.n[1]);
        const float *__restrict__ qxyz = (dir ? p.xyz[1] : p.xyz[0]) + (size_t)cq * nq * 3;
        const int nqp_max = dir ? p.np[1] : p.np[0];
        const float *__restrict__ qsoa = (dir ? p.soa[1] : p.soa[0]);
        if (qsoa) qsoa += (size_t)cq * ((size_t)nqp_max * 4 + kPackTail);
        const float *__restrict__ csoa = (dir ? p.soa[0] : p.soa[1]) + (size_t)cc * ((size_t)ncp_max * 4 + kPackTail);
        float *out_d = (dir ? p.dist[1] : p.dist[0]) + (size_t)b * nq;
        int *out_i = (dir ? p.idx[1] : p.idx[0]) + (size_t)b * nq;
        const int qt_lo = grp * tp.gtiles;
        const int qt_all = (nq + kTcM - 1) / kTcM;                       // tiles that have outputs to write
        const int qt_hi = min(qt_lo + tp.gtiles, qt_all);
        const int qt_search = nc_all > 0 ? min(qt_hi, (nq_v + kTcM - 1) / kTcM) : qt_lo;   // tiles [qt_lo, qt_search) hold valid queries
        // Clouds of more than kTcMaxC candidates are scanned range by range (ascending); every range is resolved exactly and
        // a strict '<' against the running result keeps the lowest index among equal distances (chamfer3D.cu:126's rule).
        const int nranges = qt_search > qt_lo ? (nc_all + kTcMaxC - 1) / kTcMaxC : 1;
        auto load_query = [&](int qt, float &x, float &y, float &z) {
            int j = qt * kTcM + t;
            j = j < nq_v ? j : nq_v - 1;
            if (qsoa) { x = qsoa[j]; y = qsoa[nqp_max + j]; z = qsoa[2 * nqp_max + j]; }
            else { x = qxyz[j * 3 + 0]; y = qxyz[j * 3 + 1]; z = qxyz[j * 3 + 2]; }
        };
      for (int rng = 0; rng < nranges; rng++) {
        const int k_lo = rng * kTcMaxC;
        const int nc = min(nc_all - k_lo, kTcMaxC);
        const int ncp = (nc + kChunk - 1) / kChunk * kChunk;
        const int ntiles = (ncp + kTcN - 1) / kTcN;
        const bool first_range = rng == 0, last_range = rng == nranges - 1;

        __syncthreads();   // the previous pass is finished everywhere: its operand image and X|Y|Z may be overwritten
        TC_T(t_item);
        if (qt_search > qt_lo) {
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)ncp * sizeof(float);
                mbar_arrive_expect_tx(&c_bar, 3 * bytes);
                tma_bulk_g2s(Cx, csoa + k_lo, bytes, &c_bar);
                tma_bulk_g2s(Cy, csoa + ncp_max + k_lo, bytes, &c_bar);
                tma_bulk_g2s(Cz, csoa + 2 * (size_t)ncp_max + k_lo, bytes, &c_bar);
            }
            // operand image of the candidates; rows past the padded cloud can never win (W = 3e38, coordinates 0)
            // (all of a thread's rows are fetched before the first is converted: the build is one memory round trip, not five)
            constexpr int kRowsPerThread = (kTcMaxC + kTcThreads - 1) / kTcThreads;
            float bx[kRowsPerThread], by[kRowsPerThread], bz[kRowsPerThread], bw[kRowsPerThread];
#pragma unroll
            for (int i = 0; i < kRowsPerThread; i++) {
                const int r = tid + i * kTcThreads;
                const bool real = r < ncp;
                bx[i] = real ? csoa[k_lo + r] : 0.0f;
                by[i] = real ? csoa[ncp_max + k_lo + r] : 0.0f;
                bz[i] = real ? csoa[2 * (size_t)ncp_max + k_lo + r] : 0.0f;
                bw[i] = real ? csoa[3 * (size_t)ncp_max + k_lo + r] : 3.0e38f;
            }
#pragma unroll
            for (int i = 0; i < kRowsPerThread; i++) {
                const int r = tid + i * kTcThreads;
                if (r < ntiles * kTcN) tc_write_candidate(tc_row(Bimg, r), bx[i], by[i], bz[i], bw[i]);
            }
            if (is_resolver) {   // operand rows of the first two query tiles (the later ones follow during the pass)
#pragma unroll
                for (int i = 0; i < 2; i++)
                    if (qt_lo + i < qt_search) {
                        float x, y, z;
                        load_query(qt_lo + i, x, y, z);
                        tc_write_query(tc_row(Abuf + ((qt_ctr + i) & 1) * (kTcSmemA / 2), t), x, y, z);
                    }
            }
            fence_proxy_async();   // generic-proxy writes of this thread -> visible to the tensor core's (async proxy) reads
            if (is_resolver) {
                mbar_arrive(&a_bar[qt_ctr & 1]);
                if (qt_lo + 1 < qt_search) mbar_arrive(&a_bar[(qt_ctr + 1) & 1]);
            }
        }
        __syncthreads();       // operand image complete
        TC_T(t_built);
        TC_ADD(1, t_item, t_built);
        TC_ADD(0, 0, 1);
        TC_TRACE_RESET;
        if (is_query_thread) { if (set == 0) TC_TRACE(0); } else if (is_mma_warp) { if (mma_id == 0) TC_TRACE(2); } else { TC_TRACE(3); }

        if (is_mma_warp) {   // the whole warp walks the loop (it stays converged for the block barriers); lane 0 issues
            uint32_t tc = tile_ctr, qc = qt_ctr;
            for (int qt = qt_lo; qt < qt_search; qt++, qc++) {
                TC_T(t_a0);
                mbar_wait_bounded(&a_bar[qc & 1], (qc >> 1) & 1);
                tc_fence_after();
                TC_T(t_a1);
                TC_ADDM(9, t_a0, t_a1);
                const uint32_t a_addr = smem_u32(Abuf + (qc & 1) * (kTcSmemA / 2));
                for (int j = 0; j < ntiles; j++, tc++) {
                    if (kMmaWarps > 1 && (int)(tc % kMmaWarps) != mma_id) continue;
                    const uint32_t acc = tc % kTcAcc, use = tc / kTcAcc;
                    TC_T(t_m0);
                    if (use > 0) mbar_wait_bounded(&empty_bar[acc], (use - 1) & 1);
                    TC_T(t_m1);
                    if (use > 0) tc_fence_after();
                    TC_T(t_m2);
                    TC_ADDM(6, t_m0, t_m1);
                    TC_ADDM(7, t_m1, t_m2);
                    if (is_mma_thread) {
                        const uint32_t b_addr = smem_u32(Bimg) + (uint32_t)j * (kTcN * kTcK * 2);
#pragma unroll
                        for (int k = 0; k < kTcK / 16; k++) {
                            const uint64_t da = tc_smem_desc(a_addr + k * 2 * kTcLBO), db = tc_smem_desc(b_addr + k * 2 * kTcLBO);
                            const uint32_t accumulate = k > 0;
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(tmem + acc * kTcN), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&full_bar[acc])) : "memory");
                    }
                    __syncwarp();
                    TC_T(t_m3);
                    TC_ADDM(8, t_m2, t_m3);
                }
                if (mma_id == 0) TC_TRACE(2);
            }
        }
        if (is_query_thread) {
            // The sets take the tiles in turn, so one set's barrier waits and fences overlap another's TMEM reads.  The partial
            // (best, runner-up, chunk) triples go to the resolvers through shared memory.
            uint32_t tc = tile_ctr, qc = qt_ctr;
            for (int qt = qt_lo; qt < qt_search; qt++, qc++) {
                float best = kInf, second = kInf;
                int bchunk = 0;
                for (int j = 0; j < ntiles; j++, tc++) {
                    if (!SPLIT && (int)(tc % NSETS) != set) continue;
                    const uint32_t acc = tc % kTcAcc, use = tc / kTcAcc;
                    TC_T(t_w0);
                    TC_FINE(t_w0);
                    mbar_wait_bounded(&full_bar[acc], use & 1);
                    TC_T(t_wm);
                    TC_FINE(t_wm);
                    tc_fence_after();
                    TC_T(t_w1);
                    TC_FINE(t_w1);
                    TC_ADD(2, t_w0, t_w1);
                    const int c_lo = SPLIT ? set * kChunksPerSet : 0;
                    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + acc * kTcN + c_lo * 32;
                    uint32_t va[32], vb[32];
                    URED_TMEM_LD32(va, taddr);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < kChunksPerSet; c++) {
                        // chunk c is in registers; chunk c+1 is requested before c is reduced
                        if (c + 1 < kChunksPerSet) {
                            if (c & 1) URED_TMEM_LD32(va, taddr + (c + 1) * 32); else URED_TMEM_LD32(vb, taddr + (c + 1) * 32);
                        }
                        const float cm = (c & 1) ? min32(vb) : min32(va);
                        const int chunk_id = j * kTcN + (c_lo + c) * 32;
                        const bool better = cm < best;      // strict '<' in ascending chunk order keeps the FIRST chunk holding the minimum
                        second = fminf(second, fmaxf(cm, best));
                        best = fminf(best, cm);
                        bchunk = better ? chunk_id : bchunk;
                        if (c + 1 < kChunksPerSet) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                    tc_fence_before();
                    mbar_arrive(&empty_bar[acc]);
                    TC_T(t_w2);
                    TC_FINE(t_w2);
                    TC_ADD(3, t_w1, t_w2);
                }
                TC_T(t_r0);
                const int slot = qc & 1;
                const uint32_t suse = qc >> 1;
                if (suse > 0) mbar_wait_bounded(&done_bar[slot], (suse - 1) & 1);   // the resolvers have read the slot's previous contents
                x_best[slot][set][t] = best; x_second[slot][set][t] = second; x_chunk[slot][set][t] = bchunk;
                mbar_arrive(&r_bar[slot]);
                TC_T(t_r1);
                TC_ADD(4, t_r0, t_r1);
                if (set == 0) TC_TRACE(0);
            }
            TC_T(t_end);
            TC_ADD(5, t_item, t_end);
        }
        if (is_resolver) {
            // One thread per query of the tile: merge the two sets' triples, resolve the winning chunk exactly (nn_kernel's
            // re-check, from the resident X|Y|Z), rescan ambiguous queries, write the result -- off the TMEM readers' path.
            const float cn = __fsqrt_ru(csoa[(size_t)ncp_max * 4]) * 1.000001f;   // max |c| of the candidate cloud (block tail)
            const int rot = lane & 7;
            uint32_t qc = qt_ctr;
            bool c_ready = false;
            for (int qt = qt_lo; qt < qt_search; qt++, qc++) {
                float qx, qy, qz, fx = 0.0f, fy = 0.0f, fz = 0.0f;
                load_query(qt, qx, qy, qz);
                const bool feed = qt + 2 < qt_search;
                if (feed) load_query(qt + 2, fx, fy, fz);
                const int slot = qc & 1;
                mbar_wait_bounded<200>(&r_bar[slot], (qc >> 1) & 1);
                float best = x_best[slot][0][t], second = x_second[slot][0][t];
                int bchunk = x_chunk[slot][0][t];
#pragma unroll
                for (int o = 1; o < NSETS; o++) {
                    const float ob = x_best[slot][o][t], os = x_second[slot][o][t];
                    const int oc = x_chunk[slot][o][t];
                    second = fminf(fminf(second, os), fmaxf(best, ob));
                    if (ob < best || (ob == best && oc < bchunk)) bchunk = oc;   // the FIRST chunk holding the minimum
                    best = fminf(best, ob);
                }
                mbar_arrive(&done_bar[slot]);
                if (feed) {   // every tile of query tile qt has been reduced, so its MMAs -- the last readers of this operand buffer -- are complete
                    tc_write_query(tc_row(Abuf + (qc & 1) * (kTcSmemA / 2), t), fx, fy, fz);
                    fence_proxy_async();
                    mbar_arrive(&a_bar[qc & 1]);
                }
                if (!c_ready) { mbar_wait_bounded(&c_bar, item_ctr & 1); c_ready = true; }
                const int c = min(bchunk, ncp - kChunk);
                const float2 nqx = make_float2(-qx, -qx), nqy = make_float2(-qy, -qy), nqz = make_float2(-qz, -qz);
                unsigned bdu = 0u, hdu = 0u;
                int bi = c;
                const float *sX = Cx + c, *sY = Cy + c, *sZ = Cz + c;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int piece = (k + rot) & 7;
                    if (k > 0 && piece == 0) { hdu = bdu; bdu += 1u; }
                    const float4 x4 = *reinterpret_cast<const float4 *>(sX + piece * 4);
                    const float4 y4 = *reinterpret_cast<const float4 *>(sY + piece * 4);
                    const float4 z4 = *reinterpret_cast<const float4 *>(sZ + piece * 4);
                    const float2 dx0 = __fadd2_rn(make_float2(x4.x, x4.y), nqx), dx1 = __fadd2_rn(make_float2(x4.z, x4.w), nqx);
                    const float2 dy0 = __fadd2_rn(make_float2(y4.x, y4.y), nqy), dy1 = __fadd2_rn(make_float2(y4.z, y4.w), nqy);
                    const float2 dz0 = __fadd2_rn(make_float2(z4.x, z4.y), nqz), dz1 = __fadd2_rn(make_float2(z4.z, z4.w), nqz);
                    const float2 d0 = __ffma2_rn(dz0, dz0, __ffma2_rn(dx0, dx0, __fmul2_rn(dy0, dy0)));
                    const float2 d1 = __ffma2_rn(dz1, dz1, __ffma2_rn(dx1, dx1, __fmul2_rn(dy1, dy1)));
                    const int i0 = c + piece * 4;
                    const unsigned u0 = __float_as_uint(d0.x), u1 = __float_as_uint(d0.y), u2 = __float_as_uint(d1.x), u3 = __float_as_uint(d1.y);
                    if (k == 0 || u0 < bdu) { bdu = u0; bi = i0; }
                    if (u1 < bdu) { bdu = u1; bi = i0 + 1; }
                    if (u2 < bdu) { bdu = u2; bi = i0 + 2; }
                    if (u3 < bdu) { bdu = u3; bi = i0 + 3; }
                }
                if (rot != 0 && bi >= c + rot * 4) bdu = hdu;  // the winner predates the wrap: undo the bump
                float bd = __uint_as_float(bdu);
                {
                    const float qn = __fsqrt_ru(__fmaf_rn(qz, qz, __fmaf_rn(qy, qy, qx * qx))) * 1.000001f;
                    const float S = qn + cn;
                    const float eps = __fmaf_rn(S * S, 1.9073486e-6f /* 2^-19 */, 1e-35f);
                    const bool ambiguous = !(second > best + eps);
                    unsigned todo = __ballot_sync(0xffffffffu, ambiguous);
                    while (todo) {
                        const int src = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const float ax = __shfl_sync(0xffffffffu, qx, src), ay = __shfl_sync(0xffffffffu, qy, src), az = __shfl_sync(0xffffffffu, qz, src);
                        float wd = kInf;
                        int wi = 0x7fffffff;
                        for (int k = lane; k < nc; k += 32) {
                            const float d = exact_d(Cx[k], Cy[k], Cz[k], ax, ay, az);
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
                bi += k_lo;
                const int slot_r = (qt - qt_lo) * kTcM + t;
                if (!first_range) {
                    const float rd = run_d[slot_r];
                    if (!(bd < rd)) { bd = rd; bi = run_i[slot_r]; }
                }
                if (!last_range) { run_d[slot_r] = bd; run_i[slot_r] = bi; }
                const int j = qt * kTcM + t;
                if (last_range && j < nq) {
                    if (j >= nq_v) { bd = 0.0f; bi = 0; }   // past the valid length of a ragged cloud
                    out_d[j] = bd;
                    out_i[j] = bi;
                }
                TC_TRACE(3);
            }
            // tiles without a valid query (or an empty candidate cloud): the reference leaves its zero-filled outputs untouched
            if (last_range)
                for (int qt = qt_search; qt < qt_hi; qt++) {
                    const int j = qt * kTcM + t;
                    if (j < nq) { out_d[j] = 0.0f; out_i[j] = 0; }
                }
        }
        const uint32_t did = (uint32_t)max(qt_search - qt_lo, 0);
        tile_ctr += did * (uint32_t)ntiles;
        qt_ctr += did;
        item_ctr += did ? 1u : 0u;
      }
    }
    TC_PROF_FLUSH;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// Launch shape: a kernel variant (T threads x R queries per thread, MINB resident CTAs per SM) and nsplit candidate
// splits.  Measured on B200 (profiles/r02_sweep_nn*.jsonl: every variant x split count on the BASELINE shapes):
//   * the 4-query, 128-thread shape wins whenever there are enough query tiles to give every SM two CTAs; fatter
//     warps (8-16 queries per thread) lose to it in the full kernel although they win in the isolated main loop;
//   * clouds of more than 2048 candidates are best cut into 2048-candidate ranges: each CTA then keeps its whole range
//     resident in the two shared-memory stages (exact re-check from shared memory) and the tail of the grid is made of
//     small items (cfg4: 6.64 -> 7.36 Tpair/s with 8 splits);
//   * launches that cannot fill the machine take the 256-query shape and are split down to 512 candidates per CTA.
struct NNVariant { int R, T, minb; };
constexpr int kNumVariants = 8;
const NNVariant kVariants[kNumVariants] = {
    {4, 128, 5},   // 0: default
    {4, 128, 6},   // 1
    {8, 128, 4},   // 2
    {8, 64, 8},    // 3
    {6, 128, 4},   // 4
    {2, 128, 6},   // 5: smallest tile (256 queries) for launches that cannot fill the machine otherwise
    {4, 128, 4},   // 6: with register prefetch of the next candidate group
    {4, 64, 8},    // 7
};
constexpr int kDefaultVariant = 0, kSmallVariant = 5;
struct NNShape { int variant; int R, T; int nsplit; bool split_all; };

inline long long nn_items(int B, int n1, int n2, int qt) {
    return (long long)B * ((n1 + qt - 1) / qt + (n2 + qt - 1) / qt);
}
inline int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
NNShape choose_nn_shape(int B, int n1, int n2, bool exact) {
    const int kSMs = 148;
    // (one-direction calls are sized like two-direction ones: the scratch contract has no flags)
    const int cand = n1 < n2 ? n1 : n2;
    const int max_split = cand / 512 < 1 ? 1 : cand / 512;
    int v = nn_items(B, n1, n2, 512) * 10 >= kSMs * 17 ? kDefaultVariant : kSmallVariant;
    const int forced_v = env_int("URED_NN_VARIANT", -1), forced_s = env_int("URED_NN_NSPLIT", 0);
    if (!exact && forced_v >= 0 && forced_v < kNumVariants) v = forced_v;   // EXACT is built for shapes 0 and 5 only
    if (exact && v != kDefaultVariant && v != kSmallVariant) v = kDefaultVariant;
    const int qt = kVariants[v].R * kVariants[v].T;
    int ns = 1;
    while (ns < 8 && cand / (ns * 2) >= 2048) ns *= 2;                                        // 2048-candidate ranges
    while (ns < 8 && ns * 2 <= max_split && nn_items(B, n1, n2, qt) * ns * 10 < 17 * kSMs) ns *= 2;  // fill the machine
    if (forced_s > 0) { ns = 1; while (ns < forced_s && ns < 8 && ns * 2 <= max_split) ns *= 2; }
    NNShape sh = {v, kVariants[v].R, kVariants[v].T, ns, ns > 1};
    // Tail rule (split_plan): measured on B200 it does NOT pay -- the 260-item tail of the 125-shape shard takes 49 us either
    // way, because every extra CTA costs ~5 us of slot time (TMA prologue + exact re-check) and the merge is one more launch
    // (profiles/README.md).  It stays available for experiments (URED_NN_TAIL_SPLIT=1) and is off by default.
    if (ns == 1 && env_int("URED_NN_TAIL_SPLIT", 0)) {
        int t = 1;
        while (t < 4 && t * 2 <= max_split) t *= 2;
        sh.nsplit = t;
    }
    return sh;
}

// Which items of a launch of `items` equal work items are split.  split_all: every item (chosen above).  Otherwise, when
// the launch is a few waves long and its last wave is only partly full, the items of that last wave are cut into
// candidate ranges: the block scheduler hands out CTAs in index order, so the long (unsplit) jobs run first and the
// short ones fill the machine at the end instead of leaving most SMs with one or two CTAs (cfg3 shard of 125 shapes:
// 1000 items on 740 slots).
struct SplitPlan { int full_items, split_items, nsplit; };
SplitPlan split_plan(const NNShape &sh, long long items) {
    SplitPlan pl = {(int)items, 0, 1};
    if (sh.split_all) { pl.full_items = 0; pl.split_items = (int)items; pl.nsplit = sh.nsplit; return pl; }
    if (sh.nsplit < 2) return pl;
    const long long slots = 148ll * kVariants[sh.variant].minb;
    const long long frac = items % slots;
    if (items > slots && items < 4 * slots && frac * 16 > slots && frac * 5 < slots * 4) {
        pl.split_items = (int)frac; pl.full_items = (int)(items - frac); pl.nsplit = sh.nsplit;
    }
    return pl;
}
size_t split_scratch_bytes(const NNShape &sh, long long items) {
    const SplitPlan pl = split_plan(sh, items);
    return pl.split_items ? align_up((size_t)pl.nsplit * pl.split_items * (size_t)(sh.R * sh.T) * 8, 256) : 0;
}

// ------------------------------------------------------------------------------------------
// DCD / CD epilogue
// ------------------------------------------------------------------------------------------
constexpr int kDcdThreads = 256;  // 8 warps: 6 ordered sums + 2 F-score counts; 8 CTAs/SM so that 640 pairs are ONE wave

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

// Row sum in the order of torch's reduction kernel, so that the per-pair means carry the SAME BITS as the reference's
// `dist.mean(1)` / `(1 - e*w).mean(dim=1)` (model_utils.py:39-45,57-58) and rankings built on them cannot differ by a
// last-bit swap.  ATen reduces a contiguous fp32 row of a [rows >= 16, n] tensor with a 32-lane warp per row
// (block 32 x 16, 4-wide vector loads): lane t owns the 16-byte vectors t, t+32, ... with one accumulator per vector
// slot, an unaligned head / a tail of n % 4 elements go to slot 0 of the lanes that own them, the four slots are
// added left to right, and the lanes are combined by shuffle-down with offsets 16, 8, 4, 2, 1
// (aten/src/ATen/native/cuda/Reduce.cuh: input_vectorized_thread_reduce_impl, block_x_reduce; the schedule was
// identified against torch 2.11 on a B200 with tools/diag_torch_reduce.py -- of ten candidate orders this is the only
// one that reproduces torch's sums on every row, profiles/r02_diag_torch_reduce.txt).  `f(k)` is the k-th
// element of the row, `mis` the row's misalignment in elements ((address / 4) % 4).  Result valid in lane 0.
template <typename F>
__device__ __forceinline__ float torch_row_sum(F f, int n, int mis, int lane) {
    float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
    int end = n, off = 0;  // element e of the (re-based) row is f(e + off)
    if (mis > 0) {
        if (lane >= mis && lane < 4 && lane - mis < n) acc0 = __fadd_rn(acc0, f(lane - mis));
        off = 4 - mis;
        end = n + mis - 4;
    }
    if (end > 0) {
        for (int idx = lane; idx * 4 + 3 < end; idx += 32) {
            const int e = idx * 4 + off;
            acc0 = __fadd_rn(acc0, f(e)); acc1 = __fadd_rn(acc1, f(e + 1)); acc2 = __fadd_rn(acc2, f(e + 2)); acc3 = __fadd_rn(acc3, f(e + 3));
        }
        const int tail = end - end % 4 + lane;
        if (tail < end) acc0 = __fadd_rn(acc0, f(tail + off));
    }
    float v = __fadd_rn(__fadd_rn(__fadd_rn(acc0, acc1), acc2), acc3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// The same sum over a row staged in shared memory, SQRT: of the square roots.  Aligned rows (the usual case) take 128-bit
// loads and no per-element index arithmetic; the order of the additions is the one above.
template <bool SQRT>
__device__ __forceinline__ float torch_row_sum_staged(const float *row, int n, int mis, int lane) {
    auto f = [&](float v) { return SQRT ? sqrtf(v) : v; };
    if (mis != 0 || (reinterpret_cast<uintptr_t>(row) & 15u) != 0)
        return torch_row_sum([&](int k) { return f(row[k]); }, n, mis, lane);
    float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
    const float4 *row4 = reinterpret_cast<const float4 *>(row);
    const int nvec = n >> 2;
#pragma unroll 4
    for (int idx = lane; idx < nvec; idx += 32) {
        const float4 v = row4[idx];
        acc0 = __fadd_rn(acc0, f(v.x)); acc1 = __fadd_rn(acc1, f(v.y)); acc2 = __fadd_rn(acc2, f(v.z)); acc3 = __fadd_rn(acc3, f(v.w));
    }
    const int tail = (nvec << 2) + lane;
    if (tail < n) acc0 = __fadd_rn(acc0, f(row[tail]));
    float v = __fadd_rn(__fadd_rn(__fadd_rn(acc0, acc1), acc2), acc3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

struct DcdOut {
    float *loss, *cd_p, *cd_t, *ew1, *ew2;
    float *fscore;      // optional [3, B]: f-score, precision_1, precision_2 (metrics/CD/fscore.py:3-16)
    float f_threshold;
    int B;
};

// STAGED: the per-point terms are computed by ALL threads (independent, unrolled loads: memory-level parallelism), staged
// in shared memory next to the histograms, and the six ordered row sums then run over shared memory.  Needs 8 bytes per
// point; pairs too large for that (more than 25 600 points) compute the terms inside the ordered sums instead.
template <bool STAGED, int THREADS>
__global__ void __launch_bounds__(THREADS) dcd_fwd_kernel(const float *__restrict__ dist1, const float *__restrict__ dist2,
                                                              const int *__restrict__ idx1, const int *__restrict__ idx2,
                                                              int n1_max, int n2_max, float alpha, float n_lambda, float frac_12,
                                                              float frac_21, const DcdOut out, const DcdLens lens) {
    extern __shared__ int hist[];  // count1[n2] | count2[n1]   (STAGED: later reused for d1 | d2)   then term1[n1] | term2[n2]
    __shared__ float sums[8];
    int *count1 = hist, *count2 = hist + n2_max;
    const int nt_max = n1_max + n2_max;
    float *term = reinterpret_cast<float *>(hist + nt_max);   // STAGED only
    const size_t b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *__restrict__ d1 = dist1 + b * n1_max, *__restrict__ d2 = dist2 + b * n2_max;
    const int *__restrict__ i1 = idx1 + b * n1_max, *__restrict__ i2 = idx2 + b * n2_max;
    const bool ragged = lens.len1 || lens.len2;
    const int n1 = lens.len1 ? max(0, min(lens.len1[b / lens.rep1], n1_max)) : n1_max;
    const int n2 = lens.len2 ? max(0, min(lens.len2[b % lens.mod2], n2_max)) : n2_max;
    if (ragged) {
        frac_12 = (float)((double)n2 / (double)max(n1, 1));
        frac_21 = (float)((double)n1 / (double)max(n2, 1));
        if (lens.non_reg) { frac_12 = fmaxf(frac_12, 1.0f); frac_21 = fmaxf(frac_21, 1.0f); }
    }
    if (n1 == 0 || n2 == 0) {  // an empty side: nothing to average (callers mask such pairs out)
        for (int k = tid; k < n1_max; k += THREADS) if (out.ew1) out.ew1[b * n1_max + k] = 0.0f;
        for (int k = tid; k < n2_max; k += THREADS) if (out.ew2) out.ew2[b * n2_max + k] = 0.0f;
        if (tid == 0) {
            if (out.loss) out.loss[b] = 0.0f;
            if (out.cd_p) out.cd_p[b] = 0.0f;
            if (out.cd_t) out.cd_t[b] = 0.0f;
            if (out.fscore) { out.fscore[b] = 0.0f; out.fscore[out.B + b] = 0.0f; out.fscore[2 * (size_t)out.B + b] = 0.0f; }
        }
        return;
    }
    const bool want_loss = out.loss || out.ew1 || out.ew2;
    if (want_loss) {
        for (int k = tid; k < nt_max; k += THREADS) hist[k] = 0;
        __syncthreads();
        // (rows are 16-byte aligned whenever the row stride is a multiple of four points: 128-bit loads, four points per thread and step)
        const bool vec1 = (n1_max & 3) == 0 && (reinterpret_cast<uintptr_t>(i1) & 15u) == 0;
        const bool vec2 = (n2_max & 3) == 0 && (reinterpret_cast<uintptr_t>(i2) & 15u) == 0;
        auto hist_side = [&](const int *__restrict__ ix, int n, bool vec, int *cnt) {
            int k0 = 0;
            if (vec) {
                const int4 *ix4 = reinterpret_cast<const int4 *>(ix);
#pragma unroll 4
                for (int q = tid; q < (n >> 2); q += THREADS) {
                    const int4 v = ix4[q];
                    atomicAdd(&cnt[v.x], 1); atomicAdd(&cnt[v.y], 1); atomicAdd(&cnt[v.z], 1); atomicAdd(&cnt[v.w], 1);
                }
                k0 = n & ~3;
            }
            for (int k = k0 + tid; k < n; k += THREADS) atomicAdd(&cnt[ix[k]], 1);
        };
        hist_side(i1, n1, vec1, count1);
        hist_side(i2, n2, vec2, count2);
        __syncthreads();
    }
    // model_utils.py:31,35-37: exp(-d*alpha); (count**lambda + 1e-6)**(-1) * frac; the term is 1 - e*w
    auto weight_of = [&](int c, float frac) {
        return __fmul_rn(__fdiv_rn(1.0f, __fadd_rn(pow_lambda((float)c, n_lambda), 1e-6f)), frac);
    };
    if (STAGED) {
        const bool vd1 = (n1_max & 3) == 0 && (reinterpret_cast<uintptr_t>(d1) & 15u) == 0 && (reinterpret_cast<uintptr_t>(i1) & 15u) == 0;
        const bool vd2 = (n2_max & 3) == 0 && (reinterpret_cast<uintptr_t>(d2) & 15u) == 0 && (reinterpret_cast<uintptr_t>(i2) & 15u) == 0;
        if (want_loss) {
            // the weight depends on the BIN only: turn the counts into weights in place (one division per bin instead of one
            // per point, no data-dependent work left in the per-point loop), then stream the points of each side
            float *w1 = reinterpret_cast<float *>(count1), *w2 = reinterpret_cast<float *>(count2);
            for (int j = tid; j < n2_max; j += THREADS) w1[j] = weight_of(count1[j], frac_21);
            for (int j = tid; j < n1_max; j += THREADS) w2[j] = weight_of(count2[j], frac_12);
            __syncthreads();
            const float neg_alpha = -alpha;   // (-d)*alpha == d*(-alpha) bit for bit
            auto terms_side = [&](const float *__restrict__ d, const int *__restrict__ ix, const float *w, float *__restrict__ e, float *t_out,
                                  int n, int n_max, bool vec) {
                int k0 = 0;
                if (vec && (e == nullptr || (reinterpret_cast<uintptr_t>(e) & 15u) == 0) && (reinterpret_cast<uintptr_t>(t_out) & 15u) == 0) {
                    const float4 *d4 = reinterpret_cast<const float4 *>(d);
                    const int4 *ix4 = reinterpret_cast<const int4 *>(ix);
#pragma unroll 2
                    for (int q = tid; q < (n >> 2); q += THREADS) {
                        const float4 dv = d4[q];
                        const int4 iv = ix4[q];
                        float4 ev;
                        ev.x = __fmul_rn(expf(__fmul_rn(dv.x, neg_alpha)), w[iv.x]);
                        ev.y = __fmul_rn(expf(__fmul_rn(dv.y, neg_alpha)), w[iv.y]);
                        ev.z = __fmul_rn(expf(__fmul_rn(dv.z, neg_alpha)), w[iv.z]);
                        ev.w = __fmul_rn(expf(__fmul_rn(dv.w, neg_alpha)), w[iv.w]);
                        if (e) reinterpret_cast<float4 *>(e)[q] = ev;
                        reinterpret_cast<float4 *>(t_out)[q] = make_float4(__fsub_rn(1.0f, ev.x), __fsub_rn(1.0f, ev.y), __fsub_rn(1.0f, ev.z),
                                                                           __fsub_rn(1.0f, ev.w));
                    }
                    k0 = n & ~3;
                }
                for (int k = k0 + tid; k < n; k += THREADS) {
                    const float ewk = __fmul_rn(expf(__fmul_rn(d[k], neg_alpha)), w[ix[k]]);
                    if (e) e[k] = ewk;
                    t_out[k] = __fsub_rn(1.0f, ewk);
                }
                if (e) for (int k = n + tid; k < n_max; k += THREADS) e[k] = 0.0f;   // past the valid length of a ragged cloud
            };
            terms_side(d1, i1, w1, out.ew1 ? out.ew1 + b * n1_max : nullptr, term, n1, n1_max, vd1);
            terms_side(d2, i2, w2, out.ew2 ? out.ew2 + b * n2_max : nullptr, term + n1_max, n2, n2_max, vd2);
            __syncthreads();   // every weight has been read: the histogram space now takes the distances
        }
        float *sd = reinterpret_cast<float *>(hist);
        auto copy_side = [&](const float *__restrict__ d, float *dst, int n_max, bool vec) {
            int k0 = 0;
            if (vec && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
#pragma unroll 2
                for (int q = tid; q < (n_max >> 2); q += THREADS) reinterpret_cast<float4 *>(dst)[q] = reinterpret_cast<const float4 *>(d)[q];
                k0 = n_max & ~3;
            }
            for (int k = k0 + tid; k < n_max; k += THREADS) dst[k] = d[k];
        };
        copy_side(d1, sd, n1_max, vd1);
        copy_side(d2, sd + n1_max, n2_max, vd2);
        __syncthreads();
    }
    auto ew_of = [&](float dk, int c, float frac) { return __fmul_rn(expf(__fmul_rn(-dk, alpha)), weight_of(c, frac)); };

    // warps 0-2: side 1 (term, d, sqrt d); warps 3-5: side 2; warps 6, 7: F-score counts of side 1 / 2
    const int side = warp < 6 ? warp / 3 : (warp - 6) & 1;
    const int what = warp < 6 ? warp % 3 : (warp < 8 ? 3 : 4);   // (warps past the eighth of a wide CTA have nothing to sum)
    const int n = side ? n2 : n1;
    const int n_stride = side ? n2_max : n1_max;
    const float *d = side ? d2 : d1;
    const int *ix = side ? i2 : i1;
    const int *cnt = side ? count2 : count1;
    const float frac = side ? frac_12 : frac_21;
    float *ew = side ? out.ew2 : out.ew1;
    // the row's misalignment decides which elements torch's kernel treats as the unaligned head (a ragged pair stands for
    // a per-sample call: aligned)
    const int mis = ragged ? 0 : (int)((reinterpret_cast<uintptr_t>(d) >> 2) & 3u);
    const float *sdist = reinterpret_cast<const float *>(hist) + (side ? n1_max : 0);   // STAGED: distances in shared memory
    const float *sterm = term + (side ? n1_max : 0);
    float r = 0.0f;
    if (what == 0) {
        if (want_loss) {
            if (STAGED) {
                r = torch_row_sum_staged<false>(sterm, n, mis, lane);
            } else {
                r = torch_row_sum([&](int k) {
                        const float ewk = ew_of(d[k], cnt[ix[k]], frac);
                        if (ew) ew[b * n_stride + k] = ewk;
                        return __fsub_rn(1.0f, ewk);
                    }, n, mis, lane);
                if (ew) for (int k = n + lane; k < n_stride; k += 32) ew[b * n_stride + k] = 0.0f;
            }
        }
    } else if (what == 1) {
        r = STAGED ? torch_row_sum_staged<false>(sdist, n, mis, lane) : torch_row_sum([&](int k) { return d[k]; }, n, mis, lane);
    } else if (what == 2) {
        r = STAGED ? torch_row_sum_staged<true>(sdist, n, mis, lane) : torch_row_sum([&](int k) { return sqrtf(d[k]); }, n, mis, lane);
    } else if (what == 3 && out.fscore) {
        int c = 0;
        for (int k = lane; k < n; k += 32) c += (STAGED ? sdist[k] : d[k]) < out.f_threshold ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        r = (float)c;  // a sum of 0/1 floats is exact in any order
    }
    if (lane == 0 && warp < 8) sums[warp] = r;
    __syncthreads();
    if (tid == 0) {
        // torch's mean: sum * factor with factor = float(outputs) / float(numel) (ReduceMomentKernel.cu mean_kernel_impl);
        // a ragged pair is the reference's per-sample call (one output row)
        const float rows = ragged ? 1.0f : (float)out.B;
        const float f1 = __fdiv_rn(rows, ragged ? (float)n1 : (float)((long long)out.B * n1));
        const float f2 = __fdiv_rn(rows, ragged ? (float)n2 : (float)((long long)out.B * n2));
        if (out.loss) out.loss[b] = __fdiv_rn(__fadd_rn(__fmul_rn(sums[0], f1), __fmul_rn(sums[3], f2)), 2.0f);  // model_utils.py:45
        if (out.cd_p) out.cd_p[b] = __fdiv_rn(__fadd_rn(__fmul_rn(sums[2], f1), __fmul_rn(sums[5], f2)), 2.0f);  // model_utils.py:57
        if (out.cd_t) out.cd_t[b] = __fadd_rn(__fmul_rn(sums[1], f1), __fmul_rn(sums[4], f2));                   // model_utils.py:58
        if (out.fscore) {
            const float p1 = __fmul_rn(sums[6], f1), p2 = __fmul_rn(sums[7], f2);
            float f = __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, p1), p2), __fadd_rn(p1, p2));  // fscore.py:14
            if (f != f) f = 0.0f;                                                          // fscore.py:15
            out.fscore[b] = f; out.fscore[out.B + b] = p1; out.fscore[2 * (size_t)out.B + b] = p2;
        }
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
    int one_dir;        // only cloud-1 points searched cloud 2 (K=1 kNN): cloud-2 points have no term of their own
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
    const size_t b = blockIdx.x;   // pairs on the x dimension of the grid (2^31 - 1 of them), point slabs on y
    for (int t = blockIdx.y * kGradThreads + threadIdx.x; t < per_pair; t += gridDim.y * kGradThreads) {
        const int side = t >= p.n[0] ? 1 : 0;
        const int j = side ? t - p.n[0] : t;
        const int n_own = side ? p.n[1] : p.n[0], n_oth = side ? p.n[0] : p.n[1];
        const size_t c1 = b / p.rep1, c2 = b % p.mod2;
        const size_t c_own = side ? c2 : c1, c_oth = side ? c1 : c2;
        const size_t pt = b * n_own + j;
        const int v_own = valid_len(side ? p.len[1] : p.len[0], c_own, n_own);
        const int v_oth = valid_len(side ? p.len[0] : p.len[1], c_oth, n_oth);
        if (j >= v_own || v_oth == 0 || (side && p.one_dir)) {  // past the valid length, nothing to match, or no search from this side: zero gradient
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

// One CTA per pair, atomic-free and bit-reproducible.  The scatter side of the reference's backward
// (grad_xyz2[idx1[i]] -= g_i (p1_i - p2_idx), chamfer3D.cu:168-172, six float atomics per point) is turned into a
// gather: a counting sort of the argmin indices gives, for every point j, the list of points of the other cloud that
// chose j; each list is put in ascending order and summed by ONE thread, so the result does not depend on any
// scheduling (the reference's atomics and a shared-memory CAS loop both do).  Everything lives in shared memory:
//   vec[(n1+n2)*3]  every point's own term g (p_own - p_nn)
//   seg[n2 | n1]    segment ends after the counting sort (bins = points of the cloud being pointed AT)
//   lst[n1 | n2]    the pointing points (16-bit indices), grouped by the point they chose
// Used when the pair fits (18 bytes per point) and neither cloud is broadcast over several pairs.
constexpr int kGradSmemThreads = 512;
constexpr int kGradSortMax = 32;  // longer lists (adversarial clouds: many points choosing one) are rebuilt by a linear scan

__device__ __forceinline__ float point_grad_coeff(const GradParams &p, int side, size_t b, size_t pt, int n_own) {
    const float *g_dist = side ? p.g_dist[1] : p.g_dist[0];
    float gd = g_dist ? g_dist[pt] : 0.0f;
    if (p.g_cd_t) gd += p.g_cd_t[b] / (float)n_own;
    if (p.g_cd_p) gd += p.g_cd_p[b] * 0.5f / (float)n_own * (0.5f / sqrtf((side ? p.dist[1] : p.dist[0])[pt]));
    if (p.g_loss) gd += p.g_loss[b] * 0.5f / (float)n_own * (p.alpha * (side ? p.ew[1] : p.ew[0])[pt]);
    return gd;
}

// in-place exclusive prefix sum of a[0..n) by the whole CTA (THREADS threads); scratch: 33 ints
template <int THREADS>
__device__ __forceinline__ void block_exclusive_scan(int *a, int n, int *scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + THREADS - 1) / THREADS;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    int sum = 0;
    for (int k = lo; k < hi; k++) sum += a[k];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < THREADS / 32 ? scratch[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += y;
        }
        scratch[lane] = winc - w;  // exclusive offset of each warp
    }
    __syncthreads();
    int run = scratch[warp] + inc - sum;
    for (int k = lo; k < hi; k++) { const int t = a[k]; a[k] = run; run += t; }
    __syncthreads();
}

// one side of the pair: own terms of its `v_own` valid points -> vec_own, and every point filed under the point it chose
// (pos = seg_bins[choice]++ leaves seg_bins[bin] = END of the bin's segment).  Loads are issued four points at a time.
struct GradSide {
    const float *xyz_own, *xyz_oth;   // this pair's clouds
    const int *idx;                   // argmin of every own point in the other cloud
    const float *g_dist, *dist, *ew;  // per-point upstream inputs of this side (any may be NULL), already offset to the pair
    float k_t, k_p, k_l, alpha;       // per-pair coefficients of cd_t, cd_p and the DCD loss (0 when not requested)
    int v_own;
};
typedef unsigned short grad_idx_t;   // a pair handled in shared memory has far fewer than 65 536 points per cloud
__device__ __forceinline__ void grad_own_terms(const GradSide &sd, float *__restrict__ vec_own, int n_own, int *seg_bins, grad_idx_t *lst_own,
                                               bool file) {
    constexpr int U = 4;
    const int tid = threadIdx.x;
    for (int base = 0; base < n_own; base += kGradSmemThreads * U) {
        int j2[U];
        float gd[U], ax[U], ay[U], az[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = base + u * kGradSmemThreads + tid;
            j2[u] = -1; gd[u] = 0.0f; ax[u] = ay[u] = az[u] = 0.0f;
            if (j < sd.v_own) {
                j2[u] = sd.idx[j];
                ax[u] = sd.xyz_own[j * 3 + 0]; ay[u] = sd.xyz_own[j * 3 + 1]; az[u] = sd.xyz_own[j * 3 + 2];
                float g = sd.g_dist ? sd.g_dist[j] : 0.0f;
                if (sd.k_t != 0.0f) g += sd.k_t;
                if (sd.dist) g += sd.k_p * (0.5f / sqrtf(sd.dist[j]));
                if (sd.ew) g += sd.k_l * (sd.alpha * sd.ew[j]);
                gd[u] = g;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = base + u * kGradSmemThreads + tid;
            if (j >= n_own) continue;
            float gx = 0.0f, gy = 0.0f, gz = 0.0f;
            if (j2[u] >= 0) {
                const float *o = sd.xyz_oth + j2[u] * 3;
                const float g = __fmul_rn(gd[u], 2.0f);  // chamfer3D.cu:166
                gx = __fmul_rn(g, __fsub_rn(ax[u], o[0]));
                gy = __fmul_rn(g, __fsub_rn(ay[u], o[1]));
                gz = __fmul_rn(g, __fsub_rn(az[u], o[2]));
                if (file) lst_own[atomicAdd(&seg_bins[j2[u]], 1)] = (grad_idx_t)j;
            }
            vec_own[j * 3 + 0] = gx; vec_own[j * 3 + 1] = gy; vec_own[j * 3 + 2] = gz;
        }
    }
}

// point j of one cloud: its own term minus, in ascending index order, the terms of the other cloud's points that chose it.
// Lists are short (a Poisson-like spread around one), but their lengths differ from lane to lane: up to four entries are
// ordered by a branch-free sorting network and subtracted under predicates, so a warp does not serialise over its lanes'
// different lengths; only longer lists take the data-dependent paths.
__device__ __forceinline__ void grad_gather_side(const float *__restrict__ vec_own, const float *__restrict__ vec_oth, int n_own,
                                                 const int *seg_bins, grad_idx_t *lst_oth, const int *idx_oth, int v_oth,
                                                 float *__restrict__ grad_out) {
    constexpr int kNone = 0x7fffffff;
    for (int j = threadIdx.x; j < n_own; j += kGradSmemThreads) {
        const int start = j == 0 ? 0 : seg_bins[j - 1], end = seg_bins[j];
        const int len = end - start;
        float ax = vec_own[j * 3 + 0], ay = vec_own[j * 3 + 1], az = vec_own[j * 3 + 2];
        if (len <= 4) {
            int a0 = len > 0 ? (int)lst_oth[start] : kNone, a1 = len > 1 ? (int)lst_oth[start + 1] : kNone;
            int a2 = len > 2 ? (int)lst_oth[start + 2] : kNone, a3 = len > 3 ? (int)lst_oth[start + 3] : kNone;
#define URED_CSWAP(x, y) { const int lo_ = min(x, y), hi_ = max(x, y); x = lo_; y = hi_; }
            URED_CSWAP(a0, a1) URED_CSWAP(a2, a3) URED_CSWAP(a0, a2) URED_CSWAP(a1, a3) URED_CSWAP(a1, a2)
#undef URED_CSWAP
            if (a0 != kNone) { ax = __fsub_rn(ax, vec_oth[a0 * 3 + 0]); ay = __fsub_rn(ay, vec_oth[a0 * 3 + 1]); az = __fsub_rn(az, vec_oth[a0 * 3 + 2]); }
            if (a1 != kNone) { ax = __fsub_rn(ax, vec_oth[a1 * 3 + 0]); ay = __fsub_rn(ay, vec_oth[a1 * 3 + 1]); az = __fsub_rn(az, vec_oth[a1 * 3 + 2]); }
            if (a2 != kNone) { ax = __fsub_rn(ax, vec_oth[a2 * 3 + 0]); ay = __fsub_rn(ay, vec_oth[a2 * 3 + 1]); az = __fsub_rn(az, vec_oth[a2 * 3 + 2]); }
            if (a3 != kNone) { ax = __fsub_rn(ax, vec_oth[a3 * 3 + 0]); ay = __fsub_rn(ay, vec_oth[a3 * 3 + 1]); az = __fsub_rn(az, vec_oth[a3 * 3 + 2]); }
        } else if (len <= kGradSortMax) {
            for (int u = start + 1; u < end; u++) {              // insertion sort of a short, thread-private segment
                const grad_idx_t key = lst_oth[u];
                int w = u - 1;
                while (w >= start && lst_oth[w] > key) { lst_oth[w + 1] = lst_oth[w]; w--; }
                lst_oth[w + 1] = key;
            }
            for (int u = start; u < end; u++) {
                const int i = lst_oth[u];
                ax = __fsub_rn(ax, vec_oth[i * 3 + 0]); ay = __fsub_rn(ay, vec_oth[i * 3 + 1]); az = __fsub_rn(az, vec_oth[i * 3 + 2]);
            }
        } else {
            for (int i = 0; i < v_oth; i++)                      // ascending scan of the other cloud's argmins
                if (idx_oth[i] == j) { ax = __fsub_rn(ax, vec_oth[i * 3 + 0]); ay = __fsub_rn(ay, vec_oth[i * 3 + 1]); az = __fsub_rn(az, vec_oth[i * 3 + 2]); }
        }
        // straight to global memory (consecutive threads write consecutive 12-byte triples)
        grad_out[j * 3 + 0] = ax; grad_out[j * 3 + 1] = ay; grad_out[j * 3 + 2] = az;
    }
}

__global__ void __launch_bounds__(kGradSmemThreads, 3) grad_gather_kernel(const GradParams p) {
    extern __shared__ float gsm[];
    __shared__ int scan_scratch[33];
    const size_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const int n1 = p.n[0], n2 = p.n[1], nt = n1 + n2;
    float *vec = gsm;                                   // [nt * 3]
    int *seg = reinterpret_cast<int *>(gsm + nt * 3);   // [n2 | n1]: bins over the points of cloud 2 (chosen by cloud 1), then over cloud 1
    grad_idx_t *lst = reinterpret_cast<grad_idx_t *>(seg + nt);   // [n1 | n2]: cloud-1 points grouped by their choice, then cloud-2 points
    const float *xyz1 = p.xyz[0] + b * n1 * 3, *xyz2 = p.xyz[1] + b * n2 * 3;
    const int *idx1 = p.idx[0] + b * n1, *idx2 = p.idx[1] ? p.idx[1] + b * n2 : nullptr;
    int v1 = valid_len(p.len[0], b, n1), v2 = valid_len(p.len[1], b, n2);
    if (v1 == 0 || v2 == 0) { v1 = 0; v2 = 0; }          // an empty side: no matches, zero gradients
    const int q2 = p.one_dir ? 0 : v2;                   // cloud-2 points that searched cloud 1

#ifdef URED_TC_PROFILE
    const long long gp0 = clock64();
#define GRAD_STAMP(i, since) do { __syncthreads(); if (blockIdx.x == 7 && tid == 0) g_tc_prof[i] = clock64() - (since); } while (0)
#else
#define GRAD_STAMP(i, since)
#endif
    // ---- 1. histogram of the argmins (bins = the points being chosen) ----------------------------------------------
    for (int k = tid; k < nt; k += kGradSmemThreads) seg[k] = 0;
    __syncthreads();
#pragma unroll 4
    for (int i = tid; i < v1; i += kGradSmemThreads) atomicAdd(&seg[idx1[i]], 1);          // integer shared atomics are native
#pragma unroll 4
    for (int i = tid; i < q2; i += kGradSmemThreads) atomicAdd(&seg[n2 + idx2[i]], 1);
    __syncthreads();
    // ---- 2. segment starts ------------------------------------------------------------------------------------------
    GRAD_STAMP(10, gp0);
    block_exclusive_scan<kGradSmemThreads>(seg, n2, scan_scratch);
    block_exclusive_scan<kGradSmemThreads>(seg + n2, n1, scan_scratch);
    GRAD_STAMP(11, gp0);
    // ---- 3. own terms + grouping ------------------------------------------------------------------------------------
    GradSide s1, s2;
    s1.xyz_own = xyz1; s1.xyz_oth = xyz2; s1.idx = idx1; s1.v_own = v1;
    s1.g_dist = p.g_dist[0] ? p.g_dist[0] + b * n1 : nullptr;
    s1.dist = p.g_cd_p ? p.dist[0] + b * n1 : nullptr;
    s1.ew = p.g_loss ? p.ew[0] + b * n1 : nullptr;
    s1.k_t = p.g_cd_t ? p.g_cd_t[b] / (float)max(v1, 1) : 0.0f;
    s1.k_p = p.g_cd_p ? p.g_cd_p[b] * 0.5f / (float)max(v1, 1) : 0.0f;
    s1.k_l = p.g_loss ? p.g_loss[b] * 0.5f / (float)max(v1, 1) : 0.0f;
    s1.alpha = s2.alpha = p.alpha;
    s2.xyz_own = xyz2; s2.xyz_oth = xyz1; s2.idx = idx2; s2.v_own = q2;
    s2.g_dist = p.g_dist[1] ? p.g_dist[1] + b * n2 : nullptr;
    s2.dist = p.g_cd_p ? p.dist[1] + b * n2 : nullptr;
    s2.ew = p.g_loss ? p.ew[1] + b * n2 : nullptr;
    s2.k_t = p.g_cd_t ? p.g_cd_t[b] / (float)max(v2, 1) : 0.0f;
    s2.k_p = p.g_cd_p ? p.g_cd_p[b] * 0.5f / (float)max(v2, 1) : 0.0f;
    s2.k_l = p.g_loss ? p.g_loss[b] * 0.5f / (float)max(v2, 1) : 0.0f;
    grad_own_terms(s1, vec, n1, seg, lst, true);
    grad_own_terms(s2, vec + n1 * 3, n2, seg + n2, lst + n1, true);
    __syncthreads();
    GRAD_STAMP(12, gp0);
    // ---- 4. gather ----------------------------------------------------------------------------------------------------
    // cloud-1 point j was chosen by the cloud-2 points filed in bins seg[n2 + j]; cloud-2 point j by the cloud-1 points in seg[j]
    grad_gather_side(vec, vec + n1 * 3, n1, seg + n2, lst + n1, idx2, q2, p.grad[0] + b * n1 * 3);
    grad_gather_side(vec + n1 * 3, vec, n2, seg, lst, idx1, v1, p.grad[1] + b * n2 * 3);
    GRAD_STAMP(13, gp0);
}

// ---- one CTA per (pair, cloud) --------------------------------------------------------------------------------------
// The CTA that writes the gradient of cloud `s` needs, in shared memory, only the OTHER cloud's own terms (they are fetched
// through the lists) plus the bins over its own points; the own term of the point a thread is finishing is evaluated in
// registers from global memory.  That is 14 bytes per point of the other cloud + 4 per own point -- 36 KB for 2048 + 2048,
// six CTAs per SM instead of three, and twice as many, half as long CTAs (a 32-pair training batch fills 64 SMs, not 32).
// Every own term is evaluated twice (once by each CTA of the pair) by the same code, so both see the same bits; the
// subtraction order (own term, then the choosers in ascending index) is the one of grad_gather_kernel.
struct OwnTerm { float x, y, z; int j2; };
template <int U, int THREADS>
__device__ __forceinline__ void grad_term_batch(const GradSide &sd, int base, OwnTerm (&t)[U]) {
    int j2[U];
    float gd[U], ax[U], ay[U], az[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int j = base + u * THREADS;
        const bool ok = j < sd.v_own;
        j2[u] = ok ? sd.idx[j] : -1;
        ax[u] = ok ? sd.xyz_own[j * 3 + 0] : 0.0f; ay[u] = ok ? sd.xyz_own[j * 3 + 1] : 0.0f; az[u] = ok ? sd.xyz_own[j * 3 + 2] : 0.0f;
        float g = (ok && sd.g_dist) ? sd.g_dist[j] : 0.0f;
        if (sd.k_t != 0.0f) g += sd.k_t;
        if (sd.dist) g += sd.k_p * (0.5f / sqrtf(ok ? sd.dist[j] : 1.0f));
        if (sd.ew) g += sd.k_l * (sd.alpha * (ok ? sd.ew[j] : 0.0f));
        gd[u] = g;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int k = max(j2[u], 0);
        const float ox = sd.xyz_oth[k * 3 + 0], oy = sd.xyz_oth[k * 3 + 1], oz = sd.xyz_oth[k * 3 + 2];
        const float g = j2[u] >= 0 ? __fmul_rn(gd[u], 2.0f) : 0.0f;  // chamfer3D.cu:166
        t[u].x = __fmul_rn(g, __fsub_rn(ax[u], ox));
        t[u].y = __fmul_rn(g, __fsub_rn(ay[u], oy));
        t[u].z = __fmul_rn(g, __fsub_rn(az[u], oz));
        t[u].j2 = j2[u];
    }
}

__device__ __forceinline__ void grad_side_setup(const GradParams &p, size_t b, int s, int v_search, int v_norm, GradSide &sd) {
    const int n_own = p.n[s], n_oth = p.n[1 - s];
    sd.xyz_own = p.xyz[s] + b * n_own * 3; sd.xyz_oth = p.xyz[1 - s] + b * n_oth * 3;
    sd.idx = p.idx[s] ? p.idx[s] + b * n_own : nullptr;
    sd.v_own = v_search;
    sd.g_dist = p.g_dist[s] ? p.g_dist[s] + b * n_own : nullptr;
    sd.dist = p.g_cd_p ? p.dist[s] + b * n_own : nullptr;
    sd.ew = p.g_loss ? p.ew[s] + b * n_own : nullptr;
    sd.k_t = p.g_cd_t ? p.g_cd_t[b] / (float)max(v_norm, 1) : 0.0f;
    sd.k_p = p.g_cd_p ? p.g_cd_p[b] * 0.5f / (float)max(v_norm, 1) : 0.0f;
    sd.k_l = p.g_loss ? p.g_loss[b] * 0.5f / (float)max(v_norm, 1) : 0.0f;
    sd.alpha = p.alpha;
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) grad_side_kernel(const GradParams p) {
    extern __shared__ float gsm[];
    __shared__ int scan_scratch[33];
    const size_t b = blockIdx.x >> 1;
    const int s = blockIdx.x & 1;                        // the cloud whose gradient this CTA writes
    const int tid = threadIdx.x;
    const int n_own = p.n[s], n_oth = p.n[1 - s];
    float *vec = gsm;                                            // [n_oth * 3] own terms of the other cloud's points
    int *seg = reinterpret_cast<int *>(gsm + (size_t)n_oth * 3);  // [n_own]     bins over this cloud's points
    grad_idx_t *lst = reinterpret_cast<grad_idx_t *>(seg + n_own);   // [n_oth] the other cloud's points grouped by their choice
    int v1 = valid_len(p.len[0], b, p.n[0]), v2 = valid_len(p.len[1], b, p.n[1]);
    if (v1 == 0 || v2 == 0) { v1 = 0; v2 = 0; }          // an empty side: no matches, zero gradients
    const int q2 = p.one_dir ? 0 : v2;                   // cloud-2 points that searched cloud 1
    const int s_own = s ? q2 : v1, s_oth = s ? v1 : q2;  // searching points of this cloud / of the other one
    GradSide own, oth;
    grad_side_setup(p, b, s, s_own, s ? v2 : v1, own);
    grad_side_setup(p, b, 1 - s, s_oth, s ? v1 : v2, oth);

    // ---- 1. histogram of the other cloud's argmins (bins = this cloud's points) -----------------------------------------
    for (int k = tid; k < n_own; k += THREADS) seg[k] = 0;
    __syncthreads();
#pragma unroll 4
    for (int i = tid; i < s_oth; i += THREADS) atomicAdd(&seg[oth.idx[i]], 1);
    __syncthreads();
    // ---- 2. segment starts ------------------------------------------------------------------------------------------------
    block_exclusive_scan<THREADS>(seg, n_own, scan_scratch);
    // ---- 3. the other cloud's own terms, filed under the point each one chose (seg[bin] ends as the END of the bin) -------
    {
        constexpr int U = 4;
        for (int base = tid; base < s_oth; base += THREADS * U) {
            OwnTerm t[U];
            grad_term_batch<U, THREADS>(oth, base, t);
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = base + u * THREADS;
                if (t[u].j2 >= 0) {
                    vec[i * 3 + 0] = t[u].x; vec[i * 3 + 1] = t[u].y; vec[i * 3 + 2] = t[u].z;
                    lst[atomicAdd(&seg[t[u].j2], 1)] = (grad_idx_t)i;
                }
            }
        }
    }
    __syncthreads();
    // ---- 4. this cloud's points: own term minus, in ascending index order, the terms of the points that chose it ---------
    constexpr int kNone = 0x7fffffff;
    constexpr int UG = 2;
    float *grad_out = p.grad[s] + b * n_own * 3;
    for (int base = tid; base < n_own; base += THREADS * UG) {
        OwnTerm t[UG];
        grad_term_batch<UG, THREADS>(own, base, t);
#pragma unroll
        for (int u = 0; u < UG; u++) {
            const int j = base + u * THREADS;
            if (j >= n_own) continue;
            const int start = j == 0 ? 0 : seg[j - 1], end = seg[j];
            const int len = end - start;
            float ax = t[u].x, ay = t[u].y, az = t[u].z;
            if (len <= 4) {
                int a0 = len > 0 ? (int)lst[start] : kNone, a1 = len > 1 ? (int)lst[start + 1] : kNone;
                int a2 = len > 2 ? (int)lst[start + 2] : kNone, a3 = len > 3 ? (int)lst[start + 3] : kNone;
#define URED_CSWAP(x, y) { const int lo_ = min(x, y), hi_ = max(x, y); x = lo_; y = hi_; }
                URED_CSWAP(a0, a1) URED_CSWAP(a2, a3) URED_CSWAP(a0, a2) URED_CSWAP(a1, a3) URED_CSWAP(a1, a2)
#undef URED_CSWAP
                if (a0 != kNone) { ax = __fsub_rn(ax, vec[a0 * 3 + 0]); ay = __fsub_rn(ay, vec[a0 * 3 + 1]); az = __fsub_rn(az, vec[a0 * 3 + 2]); }
                if (a1 != kNone) { ax = __fsub_rn(ax, vec[a1 * 3 + 0]); ay = __fsub_rn(ay, vec[a1 * 3 + 1]); az = __fsub_rn(az, vec[a1 * 3 + 2]); }
                if (a2 != kNone) { ax = __fsub_rn(ax, vec[a2 * 3 + 0]); ay = __fsub_rn(ay, vec[a2 * 3 + 1]); az = __fsub_rn(az, vec[a2 * 3 + 2]); }
                if (a3 != kNone) { ax = __fsub_rn(ax, vec[a3 * 3 + 0]); ay = __fsub_rn(ay, vec[a3 * 3 + 1]); az = __fsub_rn(az, vec[a3 * 3 + 2]); }
            } else if (len <= kGradSortMax) {
                for (int w0 = start + 1; w0 < end; w0++) {          // insertion sort of a short, thread-private segment
                    const grad_idx_t key = lst[w0];
                    int w = w0 - 1;
                    while (w >= start && lst[w] > key) { lst[w + 1] = lst[w]; w--; }
                    lst[w + 1] = key;
                }
                for (int w0 = start; w0 < end; w0++) {
                    const int i = lst[w0];
                    ax = __fsub_rn(ax, vec[i * 3 + 0]); ay = __fsub_rn(ay, vec[i * 3 + 1]); az = __fsub_rn(az, vec[i * 3 + 2]);
                }
            } else {
                for (int i = 0; i < s_oth; i++)                  // ascending scan of the other cloud's argmins
                    if (oth.idx[i] == j) { ax = __fsub_rn(ax, vec[i * 3 + 0]); ay = __fsub_rn(ay, vec[i * 3 + 1]); az = __fsub_rn(az, vec[i * 3 + 2]); }
            }
            grad_out[j * 3 + 0] = ax; grad_out[j * 3 + 1] = ay; grad_out[j * 3 + 2] = az;
        }
    }
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
// fused local top-k + peer exchange + merge (sharded retrieval)
// ------------------------------------------------------------------------------------------
// Every rank owns one exchange buffer that all its peers have mapped (CUDA IPC or any other peer mapping the caller
// provides).  One kernel per query batch replaces "top-k kernel -> NCCL all_gather -> merge kernel":
//   CTA q selects the k smallest (score, id) keys of row q of the local shard, STORES them straight into slot
//   [parity][my rank][q] of every peer's buffer over NVLink, publishes a release flag per peer, then spins (acquire)
//   on the world flags of row q in its OWN buffer, merges the world*k keys it has received and writes the final
//   [k] (score, id) list -- identical on every rank, because the merge key (score, id) is a total order.
// Two parity slots make the buffers reusable without any further synchronisation: a peer can run at most one exchange
// ahead of this rank (it needs this rank's flag of exchange e+1 before it can start e+2), and e+1 uses the other slot.
// The epoch lives in the buffer header and is advanced by the last CTA, so the launch is replayable from a CUDA graph.
constexpr int kXchgThreads = 256;
constexpr int kXchgMaxWorld = 16;
constexpr int kXchgMaxK = 64;
constexpr int kXchgMaxRows = 512;   // all CTAs of a launch must be co-resident while they wait for their peers
constexpr int kXchgHeaderWords = 32;  // [0] epoch, [1] CTAs done, [2] status (0 ok, 1 = a peer never arrived)

struct XchgLayout {
    int world, rows, k;
    __host__ __device__ size_t flags_off(int slot, int src, int row) const {
        return (size_t)kXchgHeaderWords * 4 + (((size_t)slot * world + src) * rows + row) * 4;
    }
    __host__ __device__ size_t data_base() const {
        return ((size_t)kXchgHeaderWords * 4 + (size_t)2 * world * rows * 4 + 255) / 256 * 256;
    }
    __host__ __device__ size_t data_off(int slot, int src, int row) const {
        return data_base() + ((((size_t)slot * world + src) * rows + row) * k) * 8;
    }
    __host__ __device__ size_t total() const { return (data_base() + (size_t)2 * world * rows * k * 8 + 255) / 256 * 256; }
};

struct XchgParams {
    unsigned char *buf[kXchgMaxWorld];  // buf[r]: rank r's exchange buffer as mapped in THIS process (buf[rank] = own)
    XchgLayout lay;
    int rank;
    const float *scores;
    int cols, idx_offset;
    float *out_scores;
    int *out_ids;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ float key_score(unsigned long long key) {
    const unsigned u = (unsigned)(key >> 32);
    return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned long long score_key(float s, int i);

__global__ void __launch_bounds__(kXchgThreads) topk_exchange_kernel(const XchgParams p) {
    __shared__ unsigned long long red[kXchgThreads / 32];
    __shared__ unsigned long long mine[kXchgMaxK];
    __shared__ unsigned long long all_keys[kXchgMaxWorld * kXchgMaxK];
    __shared__ unsigned timed_out;
    const int row = blockIdx.x, tid = threadIdx.x;
    const int world = p.lay.world, k = p.lay.k;
    unsigned char *own = p.buf[p.rank];
    unsigned *hdr = reinterpret_cast<unsigned *>(own);
    const unsigned epoch = *reinterpret_cast<volatile unsigned *>(hdr);  // advanced only after every CTA of this launch is done
    const int slot = (int)(epoch & 1u);
    if (tid == 0) timed_out = 0u;

    // ---- 1. local top-k of this row ----------------------------------------------------------------------------------
    const float *srow = p.scores + (size_t)row * p.cols;
    if (p.cols <= kXchgMaxWorld * kXchgMaxK) {
        // short rows (a small shard): every key's rank by counting the keys below it -- one pass, no k-fold reduction
        unsigned long long *keys = all_keys;   // (free until the merge)
        for (int c = tid; c < p.cols; c += kXchgThreads) keys[c] = score_key(srow[c], c + p.idx_offset);
        for (int j = p.cols + tid; j < k; j += kXchgThreads) mine[j] = ~0ull;   // fewer than k shapes: padding, sorts last
        __syncthreads();
        for (int c = tid; c < p.cols; c += kXchgThreads) {
            const unsigned long long key = keys[c];
            int below = 0;
            for (int u = 0; u < p.cols; u++) below += keys[u] < key ? 1 : 0;   // ids are unique -> keys are unique
            if (below < k) mine[below] = key;
        }
        __syncthreads();
    } else {
        // long rows: k passes, each the block-wide minimum of the keys above the previous pick
        unsigned long long last = 0ull;
        for (int it = 0; it < k; it++) {
            unsigned long long m = ~0ull;
            for (int c = tid; c < p.cols; c += kXchgThreads) {
                const unsigned long long key = score_key(srow[c], c + p.idx_offset);
                if ((it == 0 || key > last) && key < m) m = key;
            }
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
                if (other < m) m = other;
            }
            if ((tid & 31) == 0) red[tid >> 5] = m;
            __syncthreads();
            if (tid == 0) {
                for (int i = 1; i < kXchgThreads / 32; i++) if (red[i] < m) m = red[i];
                mine[it] = m;
            }
            __syncthreads();
            last = mine[it];
        }
    }

    // ---- 2. push my k keys into every rank's buffer (own included), then publish one flag per rank ------------------
    for (int t = tid; t < world * k; t += kXchgThreads) {
        const int dst = t / k, i = t - dst * k;
        unsigned long long *d = reinterpret_cast<unsigned long long *>(p.buf[dst] + p.lay.data_off(slot, p.rank, row));
        d[i] = mine[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world)
        st_release_sys(reinterpret_cast<unsigned *>(p.buf[tid] + p.lay.flags_off(slot, p.rank, row)), epoch + 1u);

    // ---- 3. wait for the world's keys of this row (bounded: a missing peer sets status instead of hanging the GPU) ----
    if (tid < world) {
        const unsigned *f = reinterpret_cast<const unsigned *>(own + p.lay.flags_off(slot, tid, row));
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(f) != epoch + 1u) {
            if ((++spins & 1023u) == 0u && globaltimer_ns() - t0 > p.timeout_ns) { atomicExch(&timed_out, 1u); break; }
        }
    }
    __syncthreads();
    if (timed_out) {
        if (tid == 0) atomicExch(hdr + 2, 1u);
        for (int i = tid; i < k; i += kXchgThreads) {
            p.out_scores[(size_t)row * k + i] = __int_as_float(0x7fc00000);
            p.out_ids[(size_t)row * k + i] = -2;
        }
    } else {
        // ---- 4. merge: rank of every received key among the world*k keys (ids are unique, padding keys are all ~0) ---
        const int total = world * k;
        for (int t = tid; t < total; t += kXchgThreads) {
            const int src = t / k, i = t - src * k;
            const unsigned long long *d = reinterpret_cast<const unsigned long long *>(own + p.lay.data_off(slot, src, row));
            all_keys[t] = __ldcg(d + i);  // written by a peer over NVLink: read at L2, never from a stale L1 line
        }
        __syncthreads();
        for (int t = tid; t < total; t += kXchgThreads) {
            const unsigned long long key = all_keys[t];
            int rank_of = 0;
            for (int u = 0; u < total; u++) {
                const unsigned long long o = all_keys[u];
                rank_of += (o < key || (o == key && u < t)) ? 1 : 0;
            }
            if (rank_of < k) {
                const bool pad = key == ~0ull;
                p.out_scores[(size_t)row * k + rank_of] = pad ? kInf : key_score(key);
                p.out_ids[(size_t)row * k + rank_of] = pad ? -1 : (int)(unsigned)(key & 0xffffffffull);
            }
        }
    }

    // ---- 5. the last CTA of the launch advances the epoch ----------------------------------------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned done = atomicAdd(hdr + 1, 1u) + 1u;
        if (done == gridDim.x) {
            hdr[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned *>(hdr) = epoch + 1u;
        }
    }
}

// ------------------------------------------------------------------------------------------
// EMD by the auction algorithm (the re-rank step after the Chamfer top-k: engine/generate_pair.py:101-104)
// ------------------------------------------------------------------------------------------
// Reference: DCD/utils_v2/metrics/EMD/emd_cuda.cu:23-316 -- per iteration seven kernel launches (count / prefix / list the
// unassigned points, Bid, GetMax, Assign) over the whole batch, `iters` times from the host.  Here ONE launch runs the
// whole auction: a thread-block cluster of 8 CTAs owns one pair of clouds, keeps the auction state in a small global
// workspace (L2-resident) and separates the phases with cluster barriers (~0.2 us each) instead of kernel boundaries;
// the bidders of an iteration are spread over all 8 x 512 threads, each group of threads scanning the candidate objects
// from a shared-memory copy.  Arithmetic follows the reference instruction for instruction (value = 3.0 - sqrt(d2) -
// price evaluated in double and rounded to float; best / second best with strict '>' in ascending index order), so the
// assignment is the reference's whenever its own outcome is defined: when two bidders for one object are within its
// 1e-6 tolerance the reference lets whichever thread stores last win (emd_cuda.cu:177-190) -- here the lowest point
// index wins, one of the outcomes the reference itself can produce.
namespace cg = cooperative_groups;
constexpr int kEmdCluster = 8;
constexpr int kEmdThreads = 512;
constexpr int kEmdChunk = 4096;   // candidate objects staged in shared memory at a time (64 KB)

struct EmdParams {
    const float *xyz1, *xyz2;   // [B, n, 3]
    float *dist;                // [B, n]
    int *assignment;            // [B, n]
    // workspace, per pair: price[n] | bid_inc[n] | assign_inv[n] | bid[n] | max_inc[n] (float bits) | max_idx[n] | unass[n] | cnt[16]
    unsigned char *ws;
    size_t ws_stride;
    int n, iters;
    float eps;
};

struct EmdBest { float best, better; int idx; };
// combination of two partial scans: the larger value wins, the LOWER index on equal values; second best with multiplicity
__device__ __forceinline__ EmdBest emd_combine(const EmdBest &a, const EmdBest &b) {
    EmdBest r;
    const bool take_b = b.best > a.best || (b.best == a.best && b.idx >= 0 && (a.idx < 0 || b.idx < a.idx));
    const EmdBest &hi = take_b ? b : a, &lo = take_b ? a : b;
    r.best = hi.best; r.idx = hi.idx;
    r.better = fmaxf(fmaxf(hi.better, lo.best), lo.better);
    return r;
}

__global__ void __cluster_dims__(kEmdCluster, 1, 1) __launch_bounds__(kEmdThreads, 2) emd_auction_kernel(const EmdParams p) {
    extern __shared__ float esm[];          // xyz2 chunk [3 * kEmdChunk] | price chunk [kEmdChunk]
    __shared__ int s_scan[kEmdThreads / 32 + 1];
    __shared__ int s_total;
    __shared__ int s_cnt[kEmdCluster + 1];   // exclusive prefix of the slices' unassigned counts (this iteration)
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const size_t pair = blockIdx.x / kEmdCluster;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n;
    const float *x1 = p.xyz1 + pair * n * 3, *x2 = p.xyz2 + pair * n * 3;
    int *assignment = p.assignment + pair * n;
    unsigned char *w = p.ws + pair * p.ws_stride;
    float *price = reinterpret_cast<float *>(w);
    float *bid_inc = price + n;
    int *assign_inv = reinterpret_cast<int *>(bid_inc + n);
    int *bid = assign_inv + n;
    unsigned *max_inc = reinterpret_cast<unsigned *>(bid + n);
    int *max_idx = reinterpret_cast<int *>(max_inc + n);
    int *unass = max_idx + n;
    int *cnt = unass + n;                   // cnt[rank] = unassigned points in this CTA's slice
    const int gthreads = kEmdCluster * kEmdThreads, gtid = rank * kEmdThreads + tid;
    float *sprice = esm + 3 * min(n, kEmdChunk);   // the launch sizes shared memory for min(n, kEmdChunk) candidates

    // ---- initial state (emd_module.py:53-57: zeros, assignment = assignment_inv = -1) --------------------------------
    for (int i = gtid; i < n; i += gthreads) {
        price[i] = 0.0f; assignment[i] = -1; assign_inv[i] = -1; max_inc[i] = 0u; max_idx[i] = 0x7fffffff;
    }
    cluster.sync();

    bool x2_resident = false;
    const int slice = (n + kEmdCluster - 1) / kEmdCluster;   // this CTA lists the unassigned points of [lo, hi)
    const int lo = min(n, rank * slice), hi = min(n, lo + slice);

    for (int it = 0; it < p.iters; it++) {
        const bool last = it == p.iters - 1;
        // ---- 1. list the unassigned points, in ascending order (emd_cuda.cu:31-100 do it with a scan and atomics) ------
        int base = 0;                        // running count of this CTA's slice
        for (int i0 = lo; i0 < hi; i0 += kEmdThreads) {
            const int i = i0 + tid;
            const int un = (i < hi && __ldcg(&assignment[i]) == -1) ? 1 : 0;
            const unsigned ballot = __ballot_sync(0xffffffffu, un);
            if (lane == 0) s_scan[warp] = __popc(ballot);
            __syncthreads();
            if (warp == 0) {
                int v = lane < kEmdThreads / 32 ? s_scan[lane] : 0, inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
                if (lane < kEmdThreads / 32) s_scan[lane] = inc - v;
                if (lane == 31) s_total = inc;
            }
            __syncthreads();
            // positions are relative to the slice for now; the slice's offset is added once all counts are known
            if (un) unass[lo + base + s_scan[warp] + __popc(ballot & ((1u << lane) - 1u))] = i;
            base += s_total;
            __syncthreads();
        }
        if (tid == 0) cnt[rank] = base;
        cluster.sync();   // (arrive.release / wait.acquire at cluster scope: orders the global-memory state between the phases)
        if (tid == 0) {
            int run = 0;
            for (int r = 0; r < kEmdCluster; r++) { s_cnt[r] = run; run += __ldcg(&cnt[r]); }
            s_cnt[kEmdCluster] = run;
        }
        __syncthreads();
        const int U = s_cnt[kEmdCluster];
        if (U == 0) break;                    // everything is assigned: the remaining iterations change nothing
        // bidder g of the iteration = g-th unassigned point: slice r holds its unassigned points at unass[lo_r ...]
        auto bidder = [&](int g) {
            int r = 0;
#pragma unroll
            for (int q = 1; q < kEmdCluster; q++) r += g >= s_cnt[q] ? 1 : 0;
            return __ldcg(&unass[min(n, r * slice) + (g - s_cnt[r])]);
        };

        // ---- 2. Bid (emd_cuda.cu:103-175): tpb threads per bidder, each scanning a range of every candidate chunk --------
        int tpb = 1;
        while (tpb < 32 && (long long)U * tpb * 2 <= gthreads) tpb *= 2;
        const int groups = gthreads / tpb;    // bidders in flight at a time
        const int sub = gtid % tpb;
        for (int g0 = 0; g0 < U; g0 += groups) {
            const int g = g0 + gtid / tpb;
            const bool active = g < U;
            int j = -1;
            float qx = 0.f, qy = 0.f, qz = 0.f;
            if (active) { j = bidder(g); qx = x1[j * 3 + 0]; qy = x1[j * 3 + 1]; qz = x1[j * 3 + 2]; }
            EmdBest acc; acc.best = -1e9f; acc.better = -1e9f; acc.idx = -1;
            for (int k2 = 0; k2 < n; k2 += kEmdChunk) {
                const int end_k = min(n, k2 + kEmdChunk) - k2;
                __syncthreads();
                if (n > kEmdChunk || !x2_resident)   // a cloud that fits one chunk is staged once for the whole auction; prices change every iteration
                    for (int t = tid; t < end_k * 3; t += kEmdThreads) esm[t] = x2[(size_t)k2 * 3 + t];
                x2_resident = true;
                for (int t = tid; t < end_k; t += kEmdThreads) sprice[t] = __ldcg(&price[k2 + t]);
                __syncthreads();
                if (active) {
                    // Sub-thread `sub` takes candidates sub, sub + tpb, ...: the lanes of a warp then read consecutive candidates
                    // (conflict-free; contiguous per-thread ranges of 2^m candidates would put all lanes on ONE bank).  The
                    // interleaving is invisible in the result: emd_combine breaks equal values by the lower index.
                    // The reference evaluates "3.0 - sqrtf(d2) - price" in DOUBLE (the literal is a double) and rounds to float;
                    // FP64 adds and conversions are the slow instructions of this loop on B200.  Screen in float first: the float
                    // value differs from the reference's by < 4e-7 * max(4, |v|), so a candidate more than `margin` below the
                    // range's second-best float value can be neither the best nor the second best -- only the few candidates
                    // above that threshold are re-evaluated the reference's way (x2*x2 + y2*y2 + z2*z2 as nvcc contracts it:
                    // fma(z,z, fma(x,x, y*y)); checked against the reference op's bits).
                    auto approx = [&](int k, float &s2) {
                        const float dx = __fsub_rn(esm[k * 3 + 0], qx), dy = __fsub_rn(esm[k * 3 + 1], qy), dz = __fsub_rn(esm[k * 3 + 2], qz);
                        s2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                        return __fsub_rn(__fsub_rn(3.0f, sqrtf(s2)), sprice[k]);
                    };
                    float t1 = -kInf, t2 = -kInf, s2;
#pragma unroll 4
                    for (int k = sub; k < end_k; k += tpb) {
                        const float v = approx(k, s2);
                        t2 = fmaxf(t2, fminf(t1, v));
                        t1 = fmaxf(t1, v);
                    }
                    const float thr = t2 - 1.9073486e-6f /* 2^-19 */ * fmaxf(4.0f, fabsf(t2));   // (-inf when the range holds < 2 candidates)
#pragma unroll 4
                    for (int k = sub; k < end_k; k += tpb) {
                        const float v = approx(k, s2);
                        if (v >= thr) {
                            const float d = (float)((3.0 - (double)sqrtf(s2)) - (double)sprice[k]);
                            if (d > acc.best) { acc.better = acc.best; acc.best = d; acc.idx = k + k2; }
                            else if (d > acc.better) acc.better = d;
                        }
                    }
                }
            }
            // combine the tpb partial scans of a bidder (a power of two within one warp)
            for (int o = 1; o < tpb; o <<= 1) {
                EmdBest oth;
                oth.best = __shfl_xor_sync(0xffffffffu, acc.best, o);
                oth.better = __shfl_xor_sync(0xffffffffu, acc.better, o);
                oth.idx = __shfl_xor_sync(0xffffffffu, acc.idx, o);
                acc = emd_combine(acc, oth);
            }
            if (active && sub == 0) {
                const float inc = __fadd_rn(__fsub_rn(acc.best, acc.better), p.eps);
                bid[j] = acc.idx;
                bid_inc[j] = inc;
                atomicMax(&max_inc[acc.idx], __float_as_uint(fmaxf(inc, 0.0f)));   // increments are >= 0: bit order == value order
                max_idx[acc.idx] = 0x7fffffff;                                      // (reset for the arg-max of step 3)
            }
        }
        cluster.sync();   // (arrive.release / wait.acquire at cluster scope: orders the global-memory state between the phases)
        // ---- 3. GetMax (emd_cuda.cu:177-190): who placed the highest bid on each object (tolerance 1e-6, in double) ------
        for (int g = gtid; g < U; g += gthreads) {
            const int j = bidder(g);
            const int b_id = __ldcg(&bid[j]);
            const double inc = (double)__ldcg(&bid_inc[j]), mx = (double)__uint_as_float(__ldcg(&max_inc[b_id]));
            if (inc - 1e-6 <= mx && mx <= inc + 1e-6) atomicMin(&max_idx[b_id], j);
        }
        cluster.sync();   // (arrive.release / wait.acquire at cluster scope: orders the global-memory state between the phases)
        // ---- 4. Assign (emd_cuda.cu:192-212) -------------------------------------------------------------------------------
        for (int g = gtid; g < U; g += gthreads) {
            const int j = bidder(g);
            const int b_id = __ldcg(&bid[j]);
            if (last || __ldcg(&max_idx[b_id]) == j) {
                const int prev = __ldcg(&assign_inv[b_id]);
                if (!last && prev != -1) assignment[prev] = -1;
                assign_inv[b_id] = j;
                assignment[j] = b_id;
                price[b_id] = __fadd_rn(__ldcg(&price[b_id]), __ldcg(&bid_inc[j]));
                max_inc[b_id] = 0u;
            }
        }
        cluster.sync();   // (arrive.release / wait.acquire at cluster scope: orders the global-memory state between the phases)
    }
    cluster.sync();
    // ---- CalcDist (emd_cuda.cu:214-224) ------------------------------------------------------------------------------------
    for (int j = gtid; j < n; j += gthreads) {
        const int k = __ldcg(&assignment[j]);
        float d = 0.0f;
        if (k >= 0) {
            const float dx = __fsub_rn(x1[j * 3 + 0], x2[k * 3 + 0]), dy = __fsub_rn(x1[j * 3 + 1], x2[k * 3 + 1]), dz = __fsub_rn(x1[j * 3 + 2], x2[k * 3 + 2]);
            d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));   // same contraction (checked against the reference op's bits)
        }
        p.dist[pair * n + j] = d;
    }
}

// emd_cuda.cu:279-300: grad_xyz1 = 2 g (x1 - x2[assignment]); xyz2 gets no gradient in the reference
__global__ void __launch_bounds__(256) emd_grad_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                      const float *__restrict__ graddist, const int *__restrict__ assignment,
                                                      float *__restrict__ gradxyz1, int n, size_t total) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= total) return;
    const size_t b = t / n;
    const int k = assignment[t];
    const float g = __fmul_rn(graddist[t], 2.0f);
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (k >= 0) {
        const float *a = xyz1 + t * 3, *o = xyz2 + (b * n + k) * 3;
        gx = __fmul_rn(g, __fsub_rn(a[0], o[0])); gy = __fmul_rn(g, __fsub_rn(a[1], o[1])); gz = __fmul_rn(g, __fsub_rn(a[2], o[2]));
    }
    gradxyz1[t * 3 + 0] = gx; gradxyz1[t * 3 + 1] = gy; gradxyz1[t * 3 + 2] = gz;
}

// ------------------------------------------------------------------------------------------
// FP32 FMA peak probe: the measured denominator of the roofline (MEASURED_PEAKS.json has no FP32 figure)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_probe_kernel(float *sink, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmaf_rn(v[i], a, b);   // 8 independent chains, 64 FFMA per iteration
    }
    float sacc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) sacc += v[i];
    if (sacc == 123.456f) sink[0] = sacc;  // keeps the chains alive without a store per thread
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

inline bool flags_env_general() { return env_int("URED_GRAD_GENERAL", 0) != 0; }  // tests: force the global-atomic backward

// which kernel screens: the tensor-core one unless the caller asks for the difference form on every pair, for the FP32-pipe
// screen (flag), or the environment switches it off (URED_NN_TC=0: A/B runs)
inline bool nn_uses_tensor_cores(unsigned flags) {
    return !(flags & (URED_FLAG_EXACT_ONLY | URED_FLAG_FP32_SCREEN)) && env_int("URED_NN_TC", 1) != 0;
}
// nn_tc_kernel: query tiles per work item.  An item costs one operand-image build plus pipeline fill / drain (about 1.05 query
// tiles' worth of time, whatever the cloud size) and `g` query tiles; the persistent CTAs take ceil(items / 148) rounds.
// Measured (B200): cfg1 41 / 46 / 56 / 66 us with g = 8 / 4 / 2 / 16, cfg2 0.488 / 0.534 / 0.620 ms with g = 16 / 8 / 4 --
// the order this estimate gives.
inline int tc_group_tiles(int B, int t1, int t2) {
    const int forced = env_int("URED_TC_GTILES", 0);
    if (forced > 0) return forced < kTcMaxGroup ? forced : kTcMaxGroup;
    const int tmax = t1 > t2 ? t1 : t2;
    int best_g = 1;
    double best_cost = 0.0;
    for (int g = 1; g <= kTcMaxGroup; g <<= 1) {
        const long long items = (long long)B * ((t1 + g - 1) / g + (t2 + g - 1) / g);
        const double cost = (double)((items + 147) / 148) * (1.05 + (double)(g < tmax ? g : tmax));
        if (g == 1 || cost <= best_cost) { best_cost = cost; best_g = g; }
        if (g >= tmax) break;
    }
    return best_g;
}

template <bool SCREEN, int R, int T, int MINB, bool PF = false>
int launch_nn(const NNParams &p, int B, cudaStream_t st) {
    constexpr int NARR = SCREEN ? 4 : 3;
    const size_t smem = (size_t)kStages * NARR * kTile * sizeof(float);
    const long long grid = (long long)p.full_items + (long long)p.split_items * p.nsplit;
    (void)B;
    if (grid > 0x7fffffffll) return fail_arg(URED_E_SHAPE, "too many work items for one launch");
    nn_kernel<SCREEN, R, T, MINB, PF><<<(unsigned)grid, T, smem, st>>>(p);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "nn_kernel launch");
}

int launch_nn_variant(int variant, bool exact, const NNParams &p, int B, cudaStream_t st) {
    // the difference-form (EXACT) kernel carries more live values per query: it is built for the two 4-query shapes only
    if (exact) return (variant == 5) ? launch_nn<false, 2, 128, 6>(p, B, st) : launch_nn<false, 4, 128, 5>(p, B, st);
    switch (variant) {
        case 0: return launch_nn<true, 4, 128, 5>(p, B, st);
        case 1: return launch_nn<true, 4, 128, 6>(p, B, st);
        case 2: return launch_nn<true, 8, 128, 4>(p, B, st);
        case 3: return launch_nn<true, 8, 64, 8>(p, B, st);
        case 4: return launch_nn<true, 6, 128, 4>(p, B, st);
        case 5: return launch_nn<true, 2, 128, 6>(p, B, st);
        case 6: return launch_nn<true, 4, 128, 4, true>(p, B, st);
        case 7: return launch_nn<true, 4, 64, 8>(p, B, st);
    }
    return fail_arg(URED_E_RANGE, "unknown nn_kernel variant");
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int ured_abi_version(void) { return URED_ABI_VERSION; }
#ifdef URED_TC_PROFILE
// development build only (-DURED_TC_PROFILE): read and clear nn_tc_kernel's phase counters
int ured_debug_tc_profile(long long *out16) {
    long long zero[16] = {0};
    cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(zero));
    cudaMemcpyToSymbol(g_tc_prof, zero, sizeof(zero));
    return 0;
}
int ured_debug_tc_trace(long long *out96) {
    cudaMemcpyFromSymbol(out96, g_tc_trace, sizeof(long long) * 96);
    return 0;
}
#endif
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
    // the caller need not know which kernel flavour (screen / exact, one or two directions) will run: the largest need of the four
    size_t need = 0;
    for (int exact = 0; exact < 2; exact++) {
        const NNShape sh = choose_nn_shape(B, n1, n2, exact != 0);
        const int qt = sh.R * sh.T;
        const long long two = nn_items(B, n1, n2, qt), one = (long long)B * ((n1 + qt - 1) / qt);
        const size_t a = split_scratch_bytes(sh, two), b = split_scratch_bytes(sh, one);
        need = a > need ? a : need;
        need = b > need ? b : need;
    }
    return need;
}

int ured_nn_launch_shape(int B, int n1, int n2, unsigned flags, int *variant, int *queries_per_cta, int *threads, int *nsplit,
                         int *items, int *split_items) {
    if (B <= 0 || n1 <= 0 || n2 <= 0) return fail_arg(URED_E_SHAPE, "ured_nn_launch_shape: empty problem");
    if (nn_uses_tensor_cores(flags)) {
        const int t1 = (n1 + kTcM - 1) / kTcM, t2 = (flags & URED_FLAG_ONE_DIRECTION) ? 0 : (n2 + kTcM - 1) / kTcM;
        const int g = tc_group_tiles(B, t1, t2);
        const int nbig = (flags & URED_FLAG_ONE_DIRECTION) ? n2 : (n1 > n2 ? n1 : n2);
        if (variant) *variant = URED_NN_VARIANT_TENSOR;
        if (queries_per_cta) *queries_per_cta = g * kTcM;
        if (threads) *threads = kTcLaunchThreads;
        if (nsplit) *nsplit = (nbig + kTcMaxC - 1) / kTcMaxC;     // candidate ranges, scanned one after the other by the same CTA
        if (items) *items = (int)((long long)B * ((t1 + g - 1) / g + (t2 + g - 1) / g));
        if (split_items) *split_items = 0;
        return 0;
    }
    const NNShape sh = choose_nn_shape(B, n1, n2, (flags & URED_FLAG_EXACT_ONLY) != 0);
    const int qt = sh.R * sh.T;
    const long long n_items = (long long)B * ((n1 + qt - 1) / qt + ((flags & URED_FLAG_ONE_DIRECTION) ? 0 : (n2 + qt - 1) / qt));
    const SplitPlan pl = split_plan(sh, n_items);
    if (variant) *variant = sh.variant;
    if (queries_per_cta) *queries_per_cta = qt;
    if (threads) *threads = sh.T;
    if (nsplit) *nsplit = pl.split_items ? pl.nsplit : 1;
    if (items) *items = (int)n_items;
    if (split_items) *split_items = pl.split_items;
    return 0;
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
    // candidates: packed2 always, packed1 unless one direction only; queries: the packed image if given, else the raw cloud
    if ((!packed1 && (!skip2 || !xyz1)) || !packed2) return fail_arg(URED_E_NULL, "ured_nn_packed: NULL input");
    PackedView v1 = view_packed(packed1, n1), v2 = view_packed(packed2, n2);
    NNParams p;
    p.xyz[0] = xyz1; p.xyz[1] = xyz2;
    p.soa[0] = packed1 ? v1.soa : nullptr; p.soa[1] = v2.soa;
    p.dist[0] = dist1; p.dist[1] = dist2;
    p.idx[0] = idx1; p.idx[1] = idx2;
    p.n[0] = n1; p.n[1] = n2;
    p.np[0] = v1.np; p.np[1] = v2.np;
    p.rep1 = rep1; p.mod2 = mod2;
    p.len[0] = len1; p.len[1] = len2;
    const bool exact = (flags & URED_FLAG_EXACT_ONLY) != 0;
    const bool one_dir = (flags & URED_FLAG_ONE_DIRECTION) != 0;  // only cloud-1 points search cloud 2 (K=1 kNN)
    // screening on the tensor cores: every candidate cloud of the call fits the resident operand image
    if (nn_uses_tensor_cores(flags)) {
        TCParams tp;
        tp.nn = p;
        tp.nn.qtiles[0] = tp.nn.qtiles[1] = 0; tp.nn.nsplit = 1; tp.nn.full_items = tp.nn.split_items = 0;
        tp.nn.part_dist = nullptr; tp.nn.part_idx = nullptr;
        const int t1 = (n1 + kTcM - 1) / kTcM, t2 = one_dir ? 0 : (n2 + kTcM - 1) / kTcM;
        tp.gtiles = tc_group_tiles(B, t1, t2);
        tp.groups[0] = (t1 + tp.gtiles - 1) / tp.gtiles;
        tp.groups[1] = (t2 + tp.gtiles - 1) / tp.gtiles;
        const long long items = (long long)B * (tp.groups[0] + tp.groups[1]);
        if (items > 0x7fffffffll) return fail_arg(URED_E_SHAPE, "too many work items for one launch");
        tp.items = (int)items;
        const int grid = (int)(items < 148 ? items : 148);   // one persistent CTA per SM (B200: 148)
        const int cfg = env_int("URED_TC_CONFIG", 0);
#define URED_TC_LAUNCH(NSETS, NACC, N, SPLIT)                                                                                           \
    do {                                                                                                                                \
        URED_CUDA(cudaFuncSetAttribute(nn_tc_kernel<NSETS, NACC, N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem), "nn_tc smem attribute"); \
        nn_tc_kernel<NSETS, NACC, N, SPLIT><<<grid, 32 * (4 * NSETS + (SPLIT ? 1 : NSETS) + 4), kTcSmem, st>>>(tp);                     \
    } while (0)
        // Measured on cfg2 (profiles/README.md, round 2): two sets taking the 256-column tiles in turn, one accumulator and one
        // MMA warp each: 0.488 ms; one set, two accumulators: 0.583; both sets on every tile (column halves): 0.549; narrower
        // tiles with more accumulators (4 x 128, 3 x 160): 0.61-0.86 -- the per-tile hand-off costs dominate.
        switch (cfg) {
            case 1: URED_TC_LAUNCH(1, 2, 256, false); break;
            case 2: URED_TC_LAUNCH(2, 2, 256, true); break;
            default: URED_TC_LAUNCH(2, 2, 256, false); break;
        }
#undef URED_TC_LAUNCH
        URED_COUNT_LAUNCH();
        return check_cuda(cudaGetLastError(), "nn_tc_kernel launch");
    }
    const NNShape sh = choose_nn_shape(B, n1, n2, exact);
    const int QT = sh.R * sh.T;
    p.qtiles[0] = (n1 + QT - 1) / QT;
    p.qtiles[1] = one_dir ? 0 : (n2 + QT - 1) / QT;
    const long long items = (long long)B * (p.qtiles[0] + p.qtiles[1]);
    if (items > 0x7fffffffll / 8) return fail_arg(URED_E_SHAPE, "too many work items for one launch");
    const SplitPlan pl = split_plan(sh, items);
    p.nsplit = pl.nsplit; p.full_items = pl.full_items; p.split_items = pl.split_items;
    p.part_dist = nullptr; p.part_idx = nullptr;
    if (pl.split_items) {
        const size_t need = split_scratch_bytes(sh, items);
        if (!scratch) return fail_arg(URED_E_NULL, "ured_nn_packed: scratch buffer required for this shape (ured_nn_scratch_bytes)");
        if ((uintptr_t)scratch % 256 || scratch_bytes < need) return fail_arg(URED_E_WORKSPACE, "ured_nn_packed: scratch too small or misaligned");
        // [dist parts | idx parts], each [split][split item][query slot of the tile]
        p.part_dist = (float *)scratch;
        p.part_idx = (int *)(p.part_dist + (size_t)pl.nsplit * pl.split_items * QT);
    }
    rc = launch_nn_variant(sh.variant, exact, p, B, st);
    if (rc || pl.split_items == 0) return rc;
    MergeParams mp;
    mp.pd = p.part_dist; mp.pi = p.part_idx;
    mp.dist[0] = dist1; mp.dist[1] = dist2;
    mp.idx[0] = idx1; mp.idx[1] = idx2;
    mp.n[0] = n1; mp.n[1] = n2;
    mp.qtiles[0] = p.qtiles[0]; mp.qtiles[1] = p.qtiles[1];
    mp.nsplit = pl.nsplit; mp.full_items = pl.full_items; mp.split_items = pl.split_items; mp.QT = QT;
    const size_t threads = (size_t)pl.split_items * QT;
    merge_splits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(mp);
    URED_COUNT_LAUNCH();
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

int ured_dcd_forward_ex(const float *dist1, const float *dist2, const int *idx1, const int *idx2, int B, int n1, int n2,
                        int rep1, int mod2, const int *len1, const int *len2, float alpha, float n_lambda, float frac_12,
                        float frac_21, unsigned flags, float *loss, float *cd_p, float *cd_t, float *ew1, float *ew2,
                        float *fscore, float f_threshold, void *stream) {
    if (B < 0 || n1 < 0 || n2 < 0) return fail_arg(URED_E_SHAPE, "negative size");
    if (rep1 < 1 || mod2 < 1) return fail_arg(URED_E_SHAPE, "rep1 and mod2 must be >= 1");
    if (B == 0) return 0;
    if (n1 == 0 || n2 == 0) return fail_arg(URED_E_SHAPE, "ured_dcd_forward: empty cloud");
    if (!dist1 || !dist2 || !idx1 || !idx2) return fail_arg(URED_E_NULL, "ured_dcd_forward: NULL input");
    const bool want_loss = loss || ew1 || ew2;
    const size_t staged_bytes = (size_t)(n1 + n2) * 8;   // histograms / distances (4 B) + terms (4 B) per point
    const bool staged = staged_bytes <= 200 * 1024;
    const size_t smem = staged ? staged_bytes : (want_loss ? (size_t)(n1 + n2) * sizeof(int) : 0);  // unstaged: cd_p / cd_t / fscore alone need no histogram
    if (smem > 200 * 1024) return fail_arg(URED_E_RANGE, "ured_dcd_forward: n1 + n2 > 51200 points per pair not supported");
    // (per device, not per thread or process: set before every launch that needs it -- the call is cheap and legal during capture)
    const bool wide = B <= 148 && env_int("URED_DCD_WIDE", 1);   // (see the launch below)
    if (smem > kSmemOptIn) {
        const void *fn = staged ? (wide ? (const void *)dcd_fwd_kernel<true, 1024> : (const void *)dcd_fwd_kernel<true, kDcdThreads>)
                                : (wide ? (const void *)dcd_fwd_kernel<false, 1024> : (const void *)dcd_fwd_kernel<false, kDcdThreads>);
        URED_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "dcd smem attribute");
    }
    DcdLens lens;
    lens.len1 = len1; lens.len2 = len2; lens.rep1 = rep1; lens.mod2 = mod2; lens.non_reg = (flags & URED_FLAG_NON_REG) ? 1 : 0;
    DcdOut out;
    out.loss = loss; out.cd_p = cd_p; out.cd_t = cd_t; out.ew1 = ew1; out.ew2 = ew2;
    out.fscore = fscore; out.f_threshold = f_threshold; out.B = B;
    // A lone CTA's time is its threads' serial work: while every pair gets an SM to itself, 1024 threads per pair finish sooner
    // (32 pairs: 13 us with 256 threads); with more pairs 256-thread CTAs keep the whole batch in one wave.
#define URED_DCD_LAUNCH(ST, TH) dcd_fwd_kernel<ST, TH><<<B, TH, smem, (cudaStream_t)stream>>>(dist1, dist2, idx1, idx2, n1, n2, alpha, n_lambda, \
                                                                                         frac_12, frac_21, out, lens)
    if (staged) { if (wide) URED_DCD_LAUNCH(true, 1024); else URED_DCD_LAUNCH(true, kDcdThreads); }
    else { if (wide) URED_DCD_LAUNCH(false, 1024); else URED_DCD_LAUNCH(false, kDcdThreads); }
#undef URED_DCD_LAUNCH
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "dcd_fwd_kernel launch");
}

int ured_dcd_forward(const float *dist1, const float *dist2, const int *idx1, const int *idx2, int B, int n1, int n2,
                     int rep1, int mod2, const int *len1, const int *len2, float alpha, float n_lambda, float frac_12,
                     float frac_21, unsigned flags, float *loss, float *cd_p, float *cd_t, float *ew1, float *ew2, void *stream) {
    return ured_dcd_forward_ex(dist1, dist2, idx1, idx2, B, n1, n2, rep1, mod2, len1, len2, alpha, n_lambda, frac_12, frac_21, flags,
                               loss, cd_p, cd_t, ew1, ew2, nullptr, 0.0f, stream);
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
    if (!xyz1 || !xyz2 || !idx1) return fail_arg(URED_E_NULL, "ured_dcd_backward: NULL input");
    if (!idx2 && (g_dist2 || g_loss || g_cd_p || g_cd_t)) return fail_arg(URED_E_NULL, "ured_dcd_backward: NULL idx2 (only the one-direction backward may omit it)");
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
    p.one_dir = idx2 ? 0 : 1;   // (ured_nn_backward_one_direction passes no idx2)
    // one CTA per (pair, cloud): 14 B per point of the other cloud + 4 B per own point (6 CTAs per SM at 2048 + 2048)
    const size_t side_need = (size_t)max(n1, n2) * 14 + (size_t)min(n1, n2) * 4;
    // Measured (profiles/README.md, round 2).  640 pairs: the pair kernel is faster (62 us against 91 us -- every own term is
    // evaluated twice here).  32 pairs: a lone CTA's time is its threads' serial work, so 64 CTAs of 1024 threads finish in
    // 15 us where 32 CTAs of 512 threads take 24 us (and 64 CTAs of 256 threads 26 us).  So: this kernel, wide, while its
    // 2 B CTAs get an SM each; also when a pair does not fit one CTA's shared memory but each side does.
    const int pair_cta = env_int("URED_GRAD_PAIR_CTA", -1);   // tests / A-B runs: 1 forces the per-pair kernel, 0 this one
    const size_t pair_need = (size_t)(n1 + n2) * 18;
    const bool few = 2ll * B <= 148;                          // 148 SMs (B200)
    const bool prefer_side = pair_cta >= 0 ? pair_cta == 0 : (few || pair_need > 200 * 1024);
    if (prefer_side && !shared1 && !shared2 && side_need <= 200 * 1024 && max(n1, n2) < 65536 && !(flags_env_general())) {
        const int wide_env = env_int("URED_GRAD_SIDE_WIDE", -1);
        const bool big = wide_env >= 0 ? wide_env != 0 : (few || side_need > 72 * 1024);   // (fewer than 3 narrow CTAs per SM would fit)
        if (side_need > kSmemOptIn) {
            if (big) URED_CUDA(cudaFuncSetAttribute(grad_side_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "grad smem attribute");
            else URED_CUDA(cudaFuncSetAttribute(grad_side_kernel<256, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "grad smem attribute");
        }
        if (big) grad_side_kernel<1024, 1><<<2 * (unsigned)B, 1024, side_need, st>>>(p);
        else grad_side_kernel<256, 6><<<2 * (unsigned)B, 256, side_need, st>>>(p);
        URED_COUNT_LAUNCH();
        return check_cuda(cudaGetLastError(), "grad_side_kernel launch");
    }
    const size_t smem_need = (size_t)(n1 + n2) * 18;  // own terms (12 B) + segment ends (4 B) + lists (2 B) per point: 3 CTAs per SM at 2048 + 2048
    if (!shared1 && !shared2 && smem_need <= 200 * 1024 && !(flags_env_general())) {
        if (smem_need > kSmemOptIn)
            URED_CUDA(cudaFuncSetAttribute(grad_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "grad smem attribute");
        grad_gather_kernel<<<B, kGradSmemThreads, smem_need, st>>>(p);
        URED_COUNT_LAUNCH();
        return check_cuda(cudaGetLastError(), "grad_gather_kernel launch");
    }
    int gx = (n1 + n2 + kGradThreads - 1) / kGradThreads;
    if (gx > 64) gx = 64;
    dim3 grid(B, gx);
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

int ured_nn_backward_one_direction(const float *xyz1, const float *xyz2, int B, int n1, int n2, const int *len2,
                                   const float *graddist1, const int *idx1, float *gradxyz1, float *gradxyz2, void *stream) {
    return ured_dcd_backward(xyz1, xyz2, B, n1, n2, 1, B > 0 ? B : 1, nullptr, len2, nullptr, nullptr, idx1, nullptr, nullptr, nullptr, 0.0f,
                             nullptr, nullptr, nullptr, graddist1, nullptr, gradxyz1, gradxyz2, stream);
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

int ured_probe_ffma(float *sink, int blocks, int iters, double *flop, void *stream) {
    if (!sink || blocks < 1 || iters < 1) return fail_arg(URED_E_SHAPE, "ured_probe_ffma: bad argument");
    ffma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, iters, 0.999f, 1e-3f);
    URED_COUNT_LAUNCH();
    if (flop) *flop = (double)blocks * 256.0 * (double)iters * 64.0 * 2.0;
    return check_cuda(cudaGetLastError(), "ffma_probe_kernel launch");
}

// ---- EMD (auction) -------------------------------------------------------------------------------------------------
size_t ured_emd_workspace_bytes(int B, int n) {
    if (B <= 0 || n <= 0) return 256;
    return align_up((size_t)B * align_up((size_t)n * 7 * 4 + 64, 256), 256);
}

int ured_emd_forward(const float *xyz1, const float *xyz2, int B, int n, float eps, int iters, float *dist, int *assignment,
                     void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || n < 0 || iters < 0) return fail_arg(URED_E_SHAPE, "ured_emd_forward: negative size");
    if (B == 0 || n == 0) return 0;
    if (!xyz1 || !xyz2 || !dist || !assignment || !workspace) return fail_arg(URED_E_NULL, "ured_emd_forward: NULL pointer");
    if ((uintptr_t)workspace % 256 || workspace_bytes < ured_emd_workspace_bytes(B, n)) return fail_arg(URED_E_WORKSPACE, "ured_emd_forward: workspace too small or misaligned");
    if ((long long)B * kEmdCluster > 0x7fffffffll) return fail_arg(URED_E_SHAPE, "ured_emd_forward: too many pairs");
    EmdParams p;
    p.xyz1 = xyz1; p.xyz2 = xyz2; p.dist = dist; p.assignment = assignment;
    p.ws = (unsigned char *)workspace; p.ws_stride = align_up((size_t)n * 7 * 4 + 64, 256);
    p.n = n; p.iters = iters; p.eps = eps;
    const size_t smem = (size_t)4 * (n < kEmdChunk ? n : kEmdChunk) * sizeof(float);
    if (smem > kSmemOptIn)
        URED_CUDA(cudaFuncSetAttribute(emd_auction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kEmdChunk * (int)sizeof(float)), "emd smem attribute");
    emd_auction_kernel<<<(unsigned)(B * kEmdCluster), kEmdThreads, smem, (cudaStream_t)stream>>>(p);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "emd_auction_kernel launch");
}

int ured_emd_backward(const float *xyz1, const float *xyz2, int B, int n, const float *graddist, const int *assignment,
                      float *gradxyz1, void *stream) {
    if (B < 0 || n < 0) return fail_arg(URED_E_SHAPE, "ured_emd_backward: negative size");
    if (B == 0 || n == 0) return 0;
    if (!xyz1 || !xyz2 || !graddist || !assignment || !gradxyz1) return fail_arg(URED_E_NULL, "ured_emd_backward: NULL pointer");
    const size_t total = (size_t)B * n;
    emd_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz1, xyz2, graddist, assignment, gradxyz1, n, total);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "emd_grad_kernel launch");
}

// ---- peer exchange (sharded retrieval) -----------------------------------------------------------------------------
size_t ured_xchg_bytes(int world, int rows, int k) {
    if (world < 1 || world > kXchgMaxWorld || rows < 1 || rows > kXchgMaxRows || k < 1 || k > kXchgMaxK) return 0;
    XchgLayout lay;
    lay.world = world; lay.rows = rows; lay.k = k;
    return lay.total();
}

int ured_xchg_alloc(size_t bytes, void **dev_ptr) {
    if (!dev_ptr || bytes == 0) return fail_arg(URED_E_NULL, "ured_xchg_alloc: NULL pointer or zero size");
    // An IPC-exportable allocation has to be a whole cudaMalloc block (not a slice of a caching allocator's pool):
    // this buffer is the one piece of device memory the library allocates itself.
    URED_CUDA(cudaMalloc(dev_ptr, bytes), "cudaMalloc(exchange buffer)");
    URED_CUDA(cudaMemset(*dev_ptr, 0, bytes), "cudaMemset(exchange buffer)");
    return check_cuda(cudaDeviceSynchronize(), "exchange buffer init");
}
int ured_xchg_free(void *dev_ptr) { return dev_ptr ? check_cuda(cudaFree(dev_ptr), "cudaFree(exchange buffer)") : 0; }

int ured_xchg_export(void *dev_ptr, void *handle64) {
    if (!dev_ptr || !handle64) return fail_arg(URED_E_NULL, "ured_xchg_export: NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == URED_XCHG_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    URED_CUDA(cudaIpcGetMemHandle(&h, dev_ptr), "cudaIpcGetMemHandle");
    memcpy(handle64, &h, sizeof(h));
    return 0;
}
int ured_xchg_import(const void *handle64, void **peer_ptr) {
    if (!handle64 || !peer_ptr) return fail_arg(URED_E_NULL, "ured_xchg_import: NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    return check_cuda(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
int ured_xchg_close(void *peer_ptr) { return peer_ptr ? check_cuda(cudaIpcCloseMemHandle(peer_ptr), "cudaIpcCloseMemHandle") : 0; }

int ured_xchg_status(const void *own_buf, int *status, unsigned *epoch, void *stream) {
    if (!own_buf || !status) return fail_arg(URED_E_NULL, "ured_xchg_status: NULL pointer");
    unsigned hdr[3];
    URED_CUDA(cudaMemcpyAsync(hdr, own_buf, sizeof(hdr), cudaMemcpyDeviceToHost, (cudaStream_t)stream), "exchange status copy");
    URED_CUDA(cudaStreamSynchronize((cudaStream_t)stream), "exchange status sync");
    *status = (int)hdr[2];
    if (epoch) *epoch = hdr[0];
    return 0;
}

int ured_topk_exchange(const float *scores, int rows, int cols, int k, int idx_offset, void *const *bufs, int world, int rank,
                       int buf_rows, float *out_scores, int *out_ids, unsigned timeout_ms, void *stream) {
    if (rows < 0 || cols < 0 || k < 1) return fail_arg(URED_E_SHAPE, "ured_topk_exchange: bad size");
    if (world < 1 || world > kXchgMaxWorld || rank < 0 || rank >= world) return fail_arg(URED_E_RANGE, "ured_topk_exchange: world <= 16, 0 <= rank < world");
    if (k > kXchgMaxK) return fail_arg(URED_E_RANGE, "ured_topk_exchange: k <= 64");
    if (rows != buf_rows || rows > kXchgMaxRows) return fail_arg(URED_E_RANGE, "ured_topk_exchange: rows must equal the buffer's row count (<= 512)");
    if (rows == 0) return 0;
    if (!bufs || !out_scores || !out_ids || (cols > 0 && !scores)) return fail_arg(URED_E_NULL, "ured_topk_exchange: NULL pointer");
    XchgParams p;
    memset(&p, 0, sizeof(p));
    for (int r = 0; r < world; r++) {
        if (!bufs[r]) return fail_arg(URED_E_NULL, "ured_topk_exchange: NULL peer buffer");
        p.buf[r] = (unsigned char *)bufs[r];
    }
    p.lay.world = world; p.lay.rows = buf_rows; p.lay.k = k;
    p.rank = rank;
    p.scores = scores; p.cols = cols; p.idx_offset = idx_offset;
    p.out_scores = out_scores; p.out_ids = out_ids;
    p.timeout_ns = (unsigned long long)(timeout_ms ? timeout_ms : 5000u) * 1000000ull;
    topk_exchange_kernel<<<rows, kXchgThreads, 0, (cudaStream_t)stream>>>(p);
    URED_COUNT_LAUNCH();
    return check_cuda(cudaGetLastError(), "topk_exchange_kernel launch");
}

}  // extern "C"
