// Micro-benchmark: FFMA vs FFMA2 (fma.rn.f32x2) throughput on sm_100a, by operand form.
//   mode 0: FFMA  R, R, R            mode 1: FFMA2 pairs, all vector registers
//   mode 2: FFMA2 acc += bcast(R.F32) * UR-pair   (taps in uniform registers: what afa_kernels.cuh issues)
//   mode 3: FFMA2 acc += bcast(R.F32) * R-pair    (taps forced into vector registers)
//   mode 4: FFMA2 acc += R-pair * UR-pair
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

struct P { float2 c[8]; };

template <int MODE>
__global__ void k(float* out, int iters, float c0, float c1, const __grid_constant__ P p) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    const float2 cc = make_float2(c0, c1);
    const float2 dd = make_float2(c1, c0);
    float2 rc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rc[i] = make_float2(p.c[i].x + threadIdx.x * 1e-9f, p.c[i].y);   // thread-dependent => vector regs
    float xs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xs[i] = 1.0f + threadIdx.x * 1e-6f * i;
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
                    if (MODE == 1) v = __ffma2_rn(v, cc, dd);
                    if (MODE == 2) v = __ffma2_rn(make_float2(xs[(i + r) & 7], xs[(i + r) & 7]), p.c[i], v);
                    if (MODE == 3) v = __ffma2_rn(make_float2(xs[(i + r) & 7], xs[(i + r) & 7]), rc[i], v);
                    if (MODE == 4) v = __ffma2_rn(v, p.c[i], dd);
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
    P p;
    for (int i = 0; i < 8; ++i) p.c[i] = make_float2(1e-3f * (i + 1), -1e-3f * (i + 1));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f, p);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f, p);
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
    const char* names[5] = {"FFMA  R,R,R          ", "FFMA2 R2,R2,R2       ", "FFMA2 bcast(R),UR2,R2", "FFMA2 bcast(R),R2,R2 ", "FFMA2 R2,UR2,R2      "};
    double f[5] = {run<0>(iters), run<1>(iters), run<2>(iters), run<3>(iters), run<4>(iters)};
    for (int i = 0; i < 5; ++i)
        printf("%s : %.2f TFMA/s (%.1f FMA/clk/SM at 1.9 GHz)\n", names[i], f[i] / 1e12, f[i] / 148 / 1.9e9);
    return 0;
}
