/*
 * c_abi_demo.c -- the drop-in boundary used from plain C, with no Python, torch or pybind in the process.
 *
 *   gcc -std=c99 -O2 -I include -I /usr/local/cuda/include examples/c_abi_demo.c \
 *       -L <package dir> -lured_chamfer -L /usr/local/cuda/lib64 -lcudart -lm -o c_abi_demo
 *   LD_LIBRARY_PATH=<package dir>:/usr/local/cuda/lib64 ./c_abi_demo
 *
 * This is what a cgo / JNI / N-API binding would do: device buffers owned by the caller, sizes and raw pointers in,
 * an error code out, work enqueued on the caller's stream.  The program checks the result against a brute-force
 * loop written with the reference's arithmetic (chamfer3D.cu:32-39: d = fma(dz,dz,fma(dx,dx,dy*dy)), lowest index wins).
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ured_chamfer.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA: %s\n", cudaGetErrorString(e_)); return 2; } } while (0)

static float frand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return (float)(*s >> 8) / 16777216.0f; }

int main(void) {
    const int B = 3, N = 700, M = 1300;
    unsigned seed = 12345u;
    float *h1 = malloc(sizeof(float) * B * N * 3), *h2 = malloc(sizeof(float) * B * M * 3);
    for (int i = 0; i < B * N * 3; i++) h1[i] = frand(&seed);
    for (int i = 0; i < B * M * 3; i++) h2[i] = frand(&seed);

    float *d1x, *d2x, *dist1, *dist2;
    int *idx1, *idx2;
    void *ws;
    size_t ws_bytes = ured_chamfer_workspace_bytes(B, N, M);
    CK(cudaMalloc((void **)&d1x, sizeof(float) * B * N * 3));
    CK(cudaMalloc((void **)&d2x, sizeof(float) * B * M * 3));
    CK(cudaMalloc((void **)&dist1, sizeof(float) * B * N));
    CK(cudaMalloc((void **)&dist2, sizeof(float) * B * M));
    CK(cudaMalloc((void **)&idx1, sizeof(int) * B * N));
    CK(cudaMalloc((void **)&idx2, sizeof(int) * B * M));
    CK(cudaMalloc(&ws, ws_bytes));
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));
    CK(cudaMemcpyAsync(d1x, h1, sizeof(float) * B * N * 3, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d2x, h2, sizeof(float) * B * M * 3, cudaMemcpyHostToDevice, stream));

    int rc = ured_chamfer_forward(d1x, d2x, B, N, M, NULL, NULL, dist1, dist2, idx1, idx2, ws, ws_bytes, 0u, stream);
    if (rc != 0) { fprintf(stderr, "ured_chamfer_forward failed (%d): %s\n", rc, ured_last_error_string()); return 3; }

    float *o1 = malloc(sizeof(float) * B * N);
    int *oi1 = malloc(sizeof(int) * B * N);
    CK(cudaMemcpyAsync(o1, dist1, sizeof(float) * B * N, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(oi1, idx1, sizeof(int) * B * N, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));

    int bad = 0;
    for (int b = 0; b < B; b++)
        for (int j = 0; j < N; j++) {
            const float *q = h1 + ((size_t)b * N + j) * 3;
            float best = 0.0f;
            int best_i = 0;
            for (int k = 0; k < M; k++) {
                const float *c = h2 + ((size_t)b * M + k) * 3;
                float dx = c[0] - q[0], dy = c[1] - q[1], dz = c[2] - q[2];
                float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                if (k == 0 || d < best) { best = d; best_i = k; }
            }
            if (best != o1[b * N + j] || best_i != oi1[b * N + j]) bad++;
        }
    /* an argument error comes back as a negative code and a message, never as a crash */
    int rc2 = ured_chamfer_forward(d1x, d2x, B, N, M, NULL, NULL, dist1, dist2, idx1, idx2, NULL, 0, 0u, stream);
    printf("c_abi_demo: %d x (%d vs %d) points, mismatches = %d, kernels launched = %llu, bad-call code = %d (%s)\n",
           B, N, M, bad, ured_kernel_launches(), rc2, ured_last_error_string());
    return bad == 0 && rc2 == URED_E_NULL ? 0 : 1;
}
