// Micro-benchmark: does a packed FFMA2 leave an issue slot for other pipes?  Per loop trip a warp executes
// 12 FFMA2 (bcast(R) * UR pair, as in the activation walks) plus N_ALU independent integer adds (ALU pipe).
// If FFMA2 occupies the dispatch port for both of its FMA-pipe cycles, time grows with N_ALU from the first add;
// if the second cycle is a free issue slot, up to 12 adds per trip ride along.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_bench issue_bench.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

struct P { float2 c[12]; };

template <int N_ALU, bool PACKED>
__global__ void k(float* out, int iters, const __grid_constant__ P p) {
    float2 a[12];
    float xs[12];
    int z[16];
#pragma unroll
    for (int i = 0; i < 12; ++i) { a[i] = make_float2(threadIdx.x * 0.001f + i, i); xs[i] = 1.0f + threadIdx.x * 1e-6f * i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            if (PACKED) {
                a[i] = __ffma2_rn(make_float2(xs[i], xs[i]), p.c[i], a[i]);
            } else {
                a[i].x = fmaf(xs[i], p.c[i].x, a[i].x);
                a[i].y = fmaf(xs[i], p.c[i].y, a[i].y);
            }
            if (i < N_ALU) asm volatile("add.s32 %0, %0, %1;" : "+r"(z[i % 16]) : "r"(it));
        }
        if (N_ALU > 12) {
#pragma unroll
            for (int i = 12; i < N_ALU; ++i) asm volatile("add.s32 %0, %0, %1;" : "+r"(z[i % 16]) : "r"(it));
        }
    }
    float s = 0.f;
    int zi = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += a[i].x + a[i].y;
#pragma unroll
    for (int i = 0; i < 16; ++i) zi += z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + zi;
}

template <int N_ALU, bool PACKED>
double run(int iters) {
    float* out;
    const int blocks = 148 * 5, threads = 128;      // 20 warps / SM, like the channels-last walk
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    P p;
    for (int i = 0; i < 12; ++i) p.c[i] = make_float2(1e-3f * (i + 1), -1e-3f * (i + 1));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<N_ALU, PACKED><<<blocks, threads>>>(out, iters, p);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<N_ALU, PACKED><<<blocks, threads>>>(out, iters, p);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(out);
    // cycles per trip per scheduler: 5 warps per SMSP
    return ms * 1e-3 * 1.9e9 / iters / 5.0;
}

int main() {
    const int iters = 20000;
    printf("cycles per loop trip per warp-slot (12 FFMA2 = 24 FMA-pipe cycles; 5 warps per scheduler, 1.9 GHz assumed)\n");
    printf("packed  FFMA2 x12 + ALU adds 0/4/8/12/16/24: %.1f %.1f %.1f %.1f %.1f %.1f\n", run<0, true>(iters), run<4, true>(iters),
           run<8, true>(iters), run<12, true>(iters), run<16, true>(iters), run<24, true>(iters));
    printf("scalar  FFMA  x24 + ALU adds 0/4/8/12/16/24: %.1f %.1f %.1f %.1f %.1f %.1f\n", run<0, false>(iters), run<4, false>(iters),
           run<8, false>(iters), run<12, false>(iters), run<16, false>(iters), run<24, false>(iters));
    return 0;
}
