// Prototype NN inner loops (exact diff-form vs 3-FFMA screen) used to size the real kernels.
#include <cuda_runtime.h>
#include <cfloat>
// prototype: inner loops to inspect SASS
template<int R, int G, int T>
__global__ void __launch_bounds__(T) exact_k(const float* __restrict__ q, const float4* __restrict__ sX, float* out, int M){
    extern __shared__ __align__(16) float sm[];
    float* X = sm; float* Y = sm + M; float* Z = sm + 2*M;
    for (int i = threadIdx.x; i < 3*M; i += blockDim.x) sm[i] = ((const float*)sX)[i];
    __syncthreads();
    float2 nqx[R], nqy[R], nqz[R]; float best[R]; int bc[R];
    #pragma unroll
    for (int r=0;r<R;r++){ int j = blockIdx.x*T*R + r*T + threadIdx.x;
        float a=-q[j*3],b=-q[j*3+1],c=-q[j*3+2]; nqx[r]=make_float2(a,a); nqy[r]=make_float2(b,b); nqz[r]=make_float2(c,c); best[r]=__int_as_float(0x7f800000); bc[r]=0;}
    for (int c0=0;c0<M;c0+=G){
        float cm[R];
        #pragma unroll
        for (int r=0;r<R;r++) cm[r]=__int_as_float(0x7f800000);
        #pragma unroll
        for (int k=0;k<G;k+=4){
            float4 x4=*(const float4*)(X+c0+k), y4=*(const float4*)(Y+c0+k), z4=*(const float4*)(Z+c0+k);
            #pragma unroll
            for (int r=0;r<R;r++){
                float2 dx0=__fadd2_rn(make_float2(x4.x,x4.y),nqx[r]), dx1=__fadd2_rn(make_float2(x4.z,x4.w),nqx[r]);
                float2 dy0=__fadd2_rn(make_float2(y4.x,y4.y),nqy[r]), dy1=__fadd2_rn(make_float2(y4.z,y4.w),nqy[r]);
                float2 dz0=__fadd2_rn(make_float2(z4.x,z4.y),nqz[r]), dz1=__fadd2_rn(make_float2(z4.z,z4.w),nqz[r]);
                float2 t0=__fmul2_rn(dy0,dy0), t1=__fmul2_rn(dy1,dy1);
                t0=__ffma2_rn(dx0,dx0,t0); t1=__ffma2_rn(dx1,dx1,t1);
                t0=__ffma2_rn(dz0,dz0,t0); t1=__ffma2_rn(dz1,dz1,t1);
                cm[r]=fminf(fminf(cm[r],t0.x),t0.y); cm[r]=fminf(fminf(cm[r],t1.x),t1.y);
            }
        }
        #pragma unroll
        for (int r=0;r<R;r++){ bool p = cm[r] < best[r]; best[r]= p?cm[r]:best[r]; bc[r]= p?c0:bc[r]; }
    }
    #pragma unroll
    for (int r=0;r<R;r++){ int j = blockIdx.x*T*R + r*T + threadIdx.x; out[j]=best[r]+bc[r]; }
}
template<int R, int G, int T>
__global__ void __launch_bounds__(T) screen_k(const float* __restrict__ q, const float4* __restrict__ sX, float* out, int M){
    extern __shared__ __align__(16) float sm[];
    float* X = sm; float* Y = sm + M; float* Z = sm + 2*M; float* W = sm+3*M;
    for (int i = threadIdx.x; i < 4*M; i += blockDim.x) sm[i] = ((const float*)sX)[i];
    __syncthreads();
    float2 qx[R], qy[R], qz[R]; float best[R], s2[R]; int bc[R];
    #pragma unroll
    for (int r=0;r<R;r++){ int j = blockIdx.x*T*R + r*T + threadIdx.x;
        float a=q[j*3],b=q[j*3+1],c=q[j*3+2]; qx[r]=make_float2(a,a); qy[r]=make_float2(b,b); qz[r]=make_float2(c,c); best[r]=__int_as_float(0x7f800000); s2[r]=best[r]; bc[r]=0;}
    for (int c0=0;c0<M;c0+=G){
        float cm[R];
        #pragma unroll
        for (int r=0;r<R;r++) cm[r]=__int_as_float(0x7f800000);
        #pragma unroll
        for (int k=0;k<G;k+=4){
            float4 x4=*(const float4*)(X+c0+k), y4=*(const float4*)(Y+c0+k), z4=*(const float4*)(Z+c0+k), w4=*(const float4*)(W+c0+k);
            #pragma unroll
            for (int r=0;r<R;r++){
                float2 t0=__ffma2_rn(make_float2(z4.x,z4.y),qz[r],make_float2(w4.x,w4.y));
                float2 t1=__ffma2_rn(make_float2(z4.z,z4.w),qz[r],make_float2(w4.z,w4.w));
                t0=__ffma2_rn(make_float2(y4.x,y4.y),qy[r],t0); t1=__ffma2_rn(make_float2(y4.z,y4.w),qy[r],t1);
                t0=__ffma2_rn(make_float2(x4.x,x4.y),qx[r],t0); t1=__ffma2_rn(make_float2(x4.z,x4.w),qx[r],t1);
                cm[r]=fminf(fminf(cm[r],t0.x),t0.y); cm[r]=fminf(fminf(cm[r],t1.x),t1.y);
            }
        }
        #pragma unroll
        for (int r=0;r<R;r++){ bool p = cm[r] < best[r]; s2[r]=fminf(s2[r],fmaxf(cm[r],best[r])); best[r]= fminf(cm[r],best[r]); bc[r]= p?c0:bc[r]; }
    }
    #pragma unroll
    for (int r=0;r<R;r++){ int j = blockIdx.x*T*R + r*T + threadIdx.x; out[j]=best[r]+bc[r]+s2[r]; }
}

#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
template <typename K> void bench(const char* name, K kern, int R, int T, int smem_arrays, const float* q, const float4* c, float* out, int M, long nq) {
    int grid = (int)(nq / (T * R));
    size_t smem = (size_t)smem_arrays * M * 4;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; w++) kern<<<grid, T, smem>>>(q, c, out, M);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { CK(cudaEventRecord(e0)); kern<<<grid, T, smem>>>(q, c, out, M); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    double pairs = (double)nq * M;
    printf("%-28s grid %6d occ %2d  %.3f ms  %.2f Tpair/s  (%.1f%% of 9.31)\n", name, grid, occ, best, pairs / best * 1e-9, pairs / best * 1e-9 / 9.31 * 100);
}
int main() {
    const int M = 2048; const long nq = 640L * 2 * 2048;
    std::vector<float> hq(nq * 3), hc(4 * M);
    srand(1); for (auto& v : hq) v = rand() / (float)RAND_MAX; for (auto& v : hc) v = rand() / (float)RAND_MAX;
    float *q, *c, *out; CK(cudaMalloc(&q, hq.size() * 4)); CK(cudaMalloc(&c, hc.size() * 4)); CK(cudaMalloc(&out, nq * 4));
    CK(cudaMemcpy(q, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(c, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
#define B(K, R, G, T, A) bench(#K "<R" #R ",G" #G ",T" #T ">", K<R, G, T>, R, T, A, q, (const float4*)c, out, M, nq)
    B(exact_k, 2, 16, 128, 3); B(exact_k, 4, 16, 128, 3); B(exact_k, 4, 32, 128, 3); B(exact_k, 8, 16, 128, 3); B(exact_k, 4, 16, 256, 3); B(exact_k, 8, 32, 256, 3);
    B(screen_k, 2, 16, 128, 4); B(screen_k, 4, 16, 128, 4); B(screen_k, 4, 32, 128, 4); B(screen_k, 8, 16, 128, 4); B(screen_k, 8, 32, 128, 4); B(screen_k, 4, 16, 256, 4); B(screen_k, 8, 32, 256, 4);
    return 0;
}
