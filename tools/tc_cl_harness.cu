// tc_cl_harness.cu -- stand-alone bring-up / timing harness for the channels-last tensor-core Activation1d kernel
// (csrc/afa_tc_cl_kernels.cuh).  Not product code and not a parity test (those live in tests/ and go through the C ABI): it
// links the kernel's translation units directly, checks the output against a double-precision loop written from the
// reference's op definitions (resample.py:29-38, activations.py:113-126, filter.py:94-101), the zero rows behind T and a
// guard band behind the tensor, and times L2-cold launches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I <csrc> tools/tc_cl_harness.cu <csrc>/afa_tc.cu
//        <csrc>/afa_tc_cl.cu -o tools/_build/tc_cl_harness
//   tc_cl_harness B C T [--tpad N] [--ny N] [--iters I] [--flags F] [--bias 0|1] [--check-cols N]
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "afa_b200.h"
#include "afa_internal.h"

namespace afa_internal {
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    return code;
}
int cuda_error(cudaError_t e, const char* what) {
    fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e));
    return (int)e;
}
void count_launch() {}
static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }
void pdl_set(int on) { g_pdl = on; }
}  // namespace afa_internal

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(2); } } while (0)

static const double kTaps[12] = {0.0020289646927267313, 0.009389465674757957, -0.0255434587597847, -0.057657383382320404,
                                 0.12857258319854736, 0.44320979714393616, 0.44320979714393616, 0.12857258319854736,
                                 -0.057657383382320404, -0.0255434587597847, 0.009389465674757957, 0.0020289646927267313};

static float bf16_round(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
    u = r & 0xFFFF0000u;
    memcpy(&f, &u, 4);
    return f;
}
static uint16_t bf16_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)(u >> 16);
}
static float bf16_to_f(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// one (batch, channel) column: y (T) in double from x + bias
static void ref_col(const std::vector<double>& x, int T, double a, double ib, std::vector<double>& y) {
    std::vector<double> s(2 * T);
    for (int n = 0; n < 2 * T; ++n) {
        double acc = 0;
        for (int k = 0; k < 12; ++k) {          // u[n] = 2 sum f[k] x[clamp((n + 5 - k) / 2)] over even n + 5 - k   (resample.py:32-36)
            const int num = n + 5 - k;
            if (num & 1) continue;
            const int ii = (num >= 0) ? num / 2 : -((-num) / 2);
            acc += 2.0 * kTaps[k] * x[clampi(ii, 0, T - 1)];
        }
        const double sn = sin(a * acc);
        s[n] = acc + ib * sn * sn;
    }
    y.assign(T, 0.0);
    for (int t = 0; t < T; ++t) {
        double acc = 0;
        for (int k = 0; k < 12; ++k) acc += kTaps[k] * s[clampi(2 * t + k - 5, 0, 2 * T - 1)];
        y[t] = acc;
    }
}

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: tc_cl_harness B C T [--tpad N] [--ny N] [--iters I] [--flags F] [--bias 0|1] [--check-cols N]\n");
        return 1;
    }
    const int B = atoi(argv[1]), C = atoi(argv[2]), T = atoi(argv[3]);
    int ny = 0, iters = 0, flags = 1, check_cols = 48, tpad = 0, use_bias = 1;
    for (int i = 4; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--ny")) ny = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--iters")) iters = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--flags")) flags = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--check-cols")) check_cols = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--tpad")) tpad = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--bias")) use_bias = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--pdl")) afa_internal::pdl_set(atoi(argv[i + 1]));
    }
    const int Tout = T + tpad;
    afa_internal::tc_cl_set_tuning(2, ny);
    const int64_t NX = (int64_t)B * T * C, NYE = (int64_t)B * Tout * C;
    const int64_t guard = 4096;
    { int32_t info[6]; if (!afa_internal::tc_cl_kernel_info(info)) printf("kernel: %d registers, %d B smem, %d threads, occupancy API: %d CTAs/SM\n", info[0], info[1], info[2], info[4]); }
    if (!afa_internal::tc_cl_eligible((void*)16, (int64_t)T * C, nullptr, (void*)16, (int64_t)Tout * C, Tout, B, C, T, AFA_DTYPE_BF16)) {
        printf("not eligible for the channels-last tensor-core path\n");
        return 1;
    }
    std::vector<uint16_t> hx(NX);
    std::vector<float> halpha(C), hbeta(C), hbias(C);
    uint64_t seed = 1234;
    auto rnd = [&]() {
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        return ((seed >> 11) + 0.5) / 9007199254740992.0;
    };
    auto gauss = [&]() { return sqrt(-2.0 * log(rnd())) * cos(6.283185307179586 * rnd()); };
    for (int64_t i = 0; i < NX; ++i) hx[i] = bf16_bits(bf16_round((float)gauss()));
    for (int c = 0; c < C; ++c) {
        halpha[c] = (float)(0.5 * gauss());
        hbeta[c] = (float)(0.5 * gauss());
        hbias[c] = use_bias ? (float)(0.3 * gauss()) : 0.f;
    }
    float tu[12], td[12];
    for (int i = 0; i < 12; ++i) tu[i] = td[i] = (float)kTaps[i];

    const size_t xbytes = (size_t)NX * 2, ybytes = (size_t)(NYE + guard) * 2;
    int nsets = iters > 0 ? (int)std::max<size_t>(2, (size_t)(400ull << 20) / (xbytes + ybytes) + 1) : 1;
    if (nsets > 64) nsets = 64;
    std::vector<void*> dx(nsets), dy(nsets);
    for (int s = 0; s < nsets; ++s) {
        CK(cudaMalloc(&dx[s], xbytes));
        CK(cudaMalloc(&dy[s], ybytes));
        CK(cudaMemcpy(dx[s], hx.data(), xbytes, cudaMemcpyHostToDevice));
        CK(cudaMemset(dy[s], 0x7B, ybytes));
    }
    float *dalpha, *dbeta, *dbias;
    CK(cudaMalloc(&dalpha, C * 4));
    CK(cudaMalloc(&dbeta, C * 4));
    CK(cudaMalloc(&dbias, C * 4));
    CK(cudaMemcpy(dalpha, halpha.data(), C * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbeta, hbeta.data(), C * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbias, hbias.data(), C * 4, cudaMemcpyHostToDevice));
    auto launch = [&](int s) {
        return afa_internal::tc_cl_fwd_launch(dx[s], (int64_t)T * C, use_bias ? dbias : nullptr, dy[s], (int64_t)Tout * C, Tout, dalpha, dbeta,
                                              tu, td, B, C, T, flags, 0);
    };
    int rc = launch(0);
    if (rc) { printf("launch failed rc=%d\n", rc); return 2; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("KERNEL FAILED: %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<uint16_t> hy(NYE + guard);
    CK(cudaMemcpy(hy.data(), dy[0], ybytes, cudaMemcpyDeviceToHost));

    // ---- check: a spread of (batch, channel) columns (all if few), the zero rows, the guard band
    const int64_t cols = (int64_t)B * C;
    std::vector<int64_t> cc;
    if (cols <= check_cols) for (int64_t r = 0; r < cols; ++r) cc.push_back(r);
    else for (int i = 0; i < check_cols; ++i) cc.push_back((int64_t)((double)i * (cols - 1) / (check_cols - 1)));
    double emax = 0, refmax = 0;
    int64_t worst_col = -1, worst_t = -1;
    std::vector<double> xc(T), yr;
    for (int64_t col : cc) {
        const int b = (int)(col / C), c = (int)(col % C);
        for (int t = 0; t < T; ++t) xc[t] = (double)bf16_to_f(hx[((int64_t)b * T + t) * C + c]) + (double)hbias[c];
        double al = halpha[c], be = (flags & 2) ? halpha[c] : hbeta[c];
        if (flags & 1) { al = exp(al); be = exp(be); }
        ref_col(xc, T, al, 1.0 / (be + 1e-9), yr);
        for (int t = 0; t < T; ++t) {
            const double got = bf16_to_f(hy[((int64_t)b * Tout + t) * C + c]);
            const double d = fabs(got - yr[t]);
            refmax = std::max(refmax, fabs(yr[t]));
            if (!(d <= emax)) { emax = d; worst_col = col; worst_t = t; }
        }
    }
    int64_t bad_zero = 0, bad_guard = 0;
    for (int b = 0; b < B; ++b)
        for (int t = T; t < Tout; ++t)
            for (int c = 0; c < C; ++c) bad_zero += hy[((int64_t)b * Tout + t) * C + c] != 0;
    for (int64_t i = 0; i < guard; ++i) bad_guard += hy[NYE + i] != 0x7B7B;
    printf("B=%d C=%d T=%d T_out=%d bias=%d: E(y) = %.3e (max |err| %.4g at col %lld t %lld, max |ref| %.4g), zero rows wrong: %lld, guard touched: %lld -> %s\n",
           B, C, T, Tout, use_bias, emax / std::max(refmax, 1e-30), emax, (long long)worst_col, (long long)worst_t, refmax, (long long)bad_zero,
           (long long)bad_guard, (emax / std::max(refmax, 1e-30) <= 1e-2 && !bad_zero && !bad_guard) ? "OK" : "FAIL");
    if (worst_col >= 0 && emax / std::max(refmax, 1e-30) > 1e-2) {
        const int b = (int)(worst_col / C), c = (int)(worst_col % C);
        for (int t = 0; t < T; ++t) xc[t] = (double)bf16_to_f(hx[((int64_t)b * T + t) * C + c]) + (double)hbias[c];
        double al = halpha[c], be = (flags & 2) ? halpha[c] : hbeta[c];
        if (flags & 1) { al = exp(al); be = exp(be); }
        ref_col(xc, T, al, 1.0 / (be + 1e-9), yr);
        const int t0 = (int)std::max<int64_t>(0, worst_t - 6), t1 = (int)std::min<int64_t>(T, worst_t + 7);
        for (int t = t0; t < t1; ++t) printf("  t=%d got %.5f ref %.5f\n", t, bf16_to_f(hy[((int64_t)b * Tout + t) * C + c]), yr[t]);
        // where do errors live? first few wrong positions of this column
        int shown = 0;
        for (int t = 0; t < T && shown < 12; ++t) {
            const double got = bf16_to_f(hy[((int64_t)b * Tout + t) * C + c]);
            if (fabs(got - yr[t]) > 2e-2 * std::max(refmax, 1e-30)) { printf("  wrong at t=%d (t%%64=%d): got %.5f ref %.5f\n", t, t % 64, got, yr[t]); ++shown; }
        }
    }
    if (iters > 0) {
        for (int i = 0; i < nsets; ++i) launch(i);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) launch(i % nsets);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / iters;
        printf("time: %.2f us per launch over %d launches (%d buffer sets), %.1f GB/s algorithmic (2 x %lld elements x 2 B)\n", us, iters, nsets,
               (double)(NX + NYE) * 2 / us / 1e3, (long long)NX);
    }
    return 0;
}
