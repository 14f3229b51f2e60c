// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue/pipe throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float c0, float c1) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    const float2 cc = make_float2(c0, c1);
    const float2 dd = make_float2(c1, c0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], c0, c1);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float2 v = make_float2(a[2 * i], a[2 * i + 1]);
                    v = __ffma2_rn(v, cc, dd);
                    a[2 * i] = v.x;
                    a[2 * i + 1] = v.y;
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(int iters) {
    float* out;
    const int blocks = 148 * 8, threads = 256;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)blocks * threads * iters * 8 * 16;
    cudaFree(out);
    return fmas / (ms * 1e-3);
}

int main() {
    const int iters = 4000;
    const double f1 = run<0>(iters), f2 = run<1>(iters);
    printf("FFMA : %.2f TFMA/s (%.1f FMA/clk/SM at 1.9 GHz)\n", f1 / 1e12, f1 / 148 / 1.9e9);
    printf("FFMA2: %.2f TFMA/s (%.1f FMA/clk/SM at 1.9 GHz)\n", f2 / 1e12, f2 / 148 / 1.9e9);
    return 0;
}
