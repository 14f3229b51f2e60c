// afa_capi.cu -- C ABI (include/afa_b200.h) over the sm_100a kernels in afa_kernels.cuh.
// Host side only: argument checks, tap folding, kernel selection, launch on the caller's stream.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "afa_b200.h"
#include "afa_kernels.cuh"

#ifndef AFA_CHUNK_LIST
#define AFA_CHUNK_LIST(X) X(5) X(9)
#endif

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
int g_tune_chunks[2] = {0, 0};
int g_tune_threads[2] = {0, 0};

int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

constexpr int kNT = 128;

struct Plan {
    int dtype;
    int chunks;
    int vec;
    int L;
    bool aligned;
    uint32_t nseg;
    uint32_t total_segs;
    uint32_t grid;
};

int default_chunks(int which, int dtype) {
    (void)which;
    (void)dtype;
    return 9;
}

int make_plan(int which, const void* p0, const void* p1, const void* p2, int64_t batch, int64_t channels, int64_t T,
              int dtype, Plan* pl) {
    if (dtype != AFA_DTYPE_F32 && dtype != AFA_DTYPE_BF16) return fail(AFA_ERR_BAD_DTYPE, "dtype %d is not AFA_DTYPE_F32/BF16", dtype);
    if (batch < 0 || channels <= 0 || T < 0) return fail(AFA_ERR_BAD_ARG, "bad shape batch=%lld channels=%lld T=%lld", (long long)batch, (long long)channels, (long long)T);
    if (T >= (1ll << 30)) return fail(AFA_ERR_TOO_LARGE, "T=%lld exceeds 2^30", (long long)T);
    pl->dtype = dtype;
    pl->vec = dtype == AFA_DTYPE_F32 ? 4 : 8;
    int ch = g_tune_chunks[which] ? g_tune_chunks[which] : default_chunks(which, dtype);
    pl->chunks = ch;
    pl->L = ch * pl->vec;
    const int64_t rows = batch * channels;
    const int64_t nseg = T > 0 ? (T + pl->L - 1) / pl->L : 0;
    const int64_t total_segs = rows * nseg;
    if (total_segs >= (1ll << 31) - kNT) return fail(AFA_ERR_TOO_LARGE, "batch*channels*ceil(T/%d)=%lld exceeds 2^31", pl->L, (long long)total_segs);
    pl->nseg = (uint32_t)nseg;
    pl->total_segs = (uint32_t)total_segs;
    pl->grid = (uint32_t)((total_segs + kNT - 1) / kNT);
    const size_t esz = dtype == AFA_DTYPE_F32 ? 4 : 2;
    const uintptr_t ptr_or = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2;
    if (ptr_or & (esz - 1)) return fail(AFA_ERR_ALIGNMENT, "tensor pointers must be aligned to the element size");
    pl->aligned = (T % pl->vec == 0) && ((ptr_or & 15) == 0);
    return 0;
}

void fold_fwd_taps(const float* up, const float* dn, afa::FwdTaps* t) {
    for (int j = 0; j < 6; ++j) {
        t->ue[j] = 2.0f * up[2 * j + 1];   // ratio * conv_transpose taps            resample.py:33
        t->uo[j] = 2.0f * up[2 * j];
    }
    for (int k = 0; k < 12; ++k) t->dn[k] = dn[k];
}
void fold_bwd_taps(const float* up, const float* dn, afa::BwdTaps* t) {
    for (int j = 0; j < 6; ++j) {
        t->ue[j] = 2.0f * up[2 * j + 1];
        t->uo[j] = 2.0f * up[2 * j];
        t->de[j] = dn[2 * j + 1];
        t->dod[j] = dn[2 * j];
    }
    // adjoint of the replicate pad of the activated signal (filter.py:98): taps that fall on the pad
    t->lo[0] = dn[0] + dn[1] + dn[2] + dn[3] + dn[4];
    t->lo[1] = dn[0] + dn[1] + dn[2];
    t->lo[2] = dn[0];
    t->hi[0] = dn[7] + dn[8] + dn[9] + dn[10] + dn[11];
    t->hi[1] = dn[9] + dn[10] + dn[11];
    t->hi[2] = dn[11];
}

template <typename K>
int prepare(K kernel, size_t smem) {
    // set the attribute once per (kernel, device) per thread
    static thread_local const void* last = nullptr;
    static thread_local int last_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (last == (const void*)kernel && last_dev == dev) return 0;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    last = (const void*)kernel;
    last_dev = dev;
    return 0;
}

template <typename T, int CH, bool AL>
int launch_fwd_t(const afa::FwdArgs& a, uint32_t grid, cudaStream_t st) {
    auto k = afa::afa_fwd_kernel<T, CH, kNT, AL>;
    const size_t smem = afa::Tile<T, CH, kNT>::fwd_smem();
    if (int rc = prepare(k, smem)) return rc;
    k<<<grid, kNT, smem, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_fwd_kernel launch");
}
template <typename T, int CH, bool AL>
int launch_bwd_t(const afa::BwdArgs& a, uint32_t grid, cudaStream_t st) {
    auto k = afa::afa_bwd_kernel<T, CH, kNT, AL>;
    const size_t smem = afa::Tile<T, CH, kNT>::bwd_smem();
    if (int rc = prepare(k, smem)) return rc;
    k<<<grid, kNT, smem, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_bwd_kernel launch");
}

template <typename T>
int launch_fwd(const Plan& pl, const afa::FwdArgs& a, cudaStream_t st) {
#define X(CH)                                                                                   \
    if (pl.chunks == CH)                                                                        \
        return pl.aligned ? launch_fwd_t<T, CH, true>(a, pl.grid, st) : launch_fwd_t<T, CH, false>(a, pl.grid, st);
    AFA_CHUNK_LIST(X)
#undef X
    return fail(AFA_ERR_BAD_ARG, "no kernel compiled for %d chunks per segment", pl.chunks);
}
template <typename T>
int launch_bwd(const Plan& pl, const afa::BwdArgs& a, cudaStream_t st) {
#define X(CH)                                                                                   \
    if (pl.chunks == CH)                                                                        \
        return pl.aligned ? launch_bwd_t<T, CH, true>(a, pl.grid, st) : launch_bwd_t<T, CH, false>(a, pl.grid, st);
    AFA_CHUNK_LIST(X)
#undef X
    return fail(AFA_ERR_BAD_ARG, "no kernel compiled for %d chunks per segment", pl.chunks);
}

template <typename T, int CH>
const void* kernel_ptr(int which, bool aligned) {
    if (which == 0) return aligned ? (const void*)afa::afa_fwd_kernel<T, CH, kNT, true> : (const void*)afa::afa_fwd_kernel<T, CH, kNT, false>;
    return aligned ? (const void*)afa::afa_bwd_kernel<T, CH, kNT, true> : (const void*)afa::afa_bwd_kernel<T, CH, kNT, false>;
}
template <typename T, int CH>
size_t kernel_smem(int which) {
    return which == 0 ? afa::Tile<T, CH, kNT>::fwd_smem() : afa::Tile<T, CH, kNT>::bwd_smem();
}

}  // namespace

extern "C" {

int afa_version(void) { return AFA_VERSION; }
const char* afa_last_error(void) { return g_err; }
int64_t afa_launch_count(void) { return (int64_t)g_launches.load(); }

int afa_set_tuning(int which, int chunks, int threads) {
    if (which < 0 || which > 1) return fail(AFA_ERR_BAD_ARG, "which must be 0 (fwd) or 1 (bwd)");
    bool ok = chunks == 0;
#define X(CH) ok = ok || chunks == CH;
    AFA_CHUNK_LIST(X)
#undef X
    if (!ok) return fail(AFA_ERR_BAD_ARG, "chunks=%d is not one of the compiled segment sizes", chunks);
    if (threads != 0 && threads != kNT) return fail(AFA_ERR_BAD_ARG, "threads=%d: only %d is compiled", threads, kNT);
    g_tune_chunks[which] = chunks;
    g_tune_threads[which] = threads;
    return 0;
}

int afa_activation1d_fwd(const void* x, void* y, const float* alpha, const float* beta, const float* taps_up12,
                         const float* taps_down12, int64_t batch, int64_t channels, int64_t T, int dtype, int flags,
                         void* stream) {
    if (!x || !y || !alpha || !taps_up12 || !taps_down12) return fail(AFA_ERR_BAD_ARG, "null pointer argument");
    if (!(flags & AFA_FLAG_SNAKE) && !beta) return fail(AFA_ERR_BAD_ARG, "beta is required unless AFA_FLAG_SNAKE is set");
    if (x == y) return fail(AFA_ERR_BAD_ARG, "y must not alias x");
    Plan pl;
    if (int rc = make_plan(0, x, y, nullptr, batch, channels, T, dtype, &pl)) return rc;
    if (pl.total_segs == 0) return 0;
    afa::FwdArgs a;
    a.x = x;
    a.y = y;
    a.alpha = alpha;
    a.beta = beta;
    fold_fwd_taps(taps_up12, taps_down12, &a.taps);
    a.total = batch * channels * T;
    a.total_segs = pl.total_segs;
    a.nseg = pl.nseg;
    a.T = (int32_t)T;
    a.C = (int32_t)channels;
    a.flags = flags;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == AFA_DTYPE_F32 ? launch_fwd<float>(pl, a, st) : launch_fwd<__nv_bfloat16>(pl, a, st);
}

size_t afa_bwd_workspace_bytes(int64_t batch, int64_t channels, int64_t T, int dtype) {
    Plan pl;
    if (make_plan(1, nullptr, nullptr, nullptr, batch, channels, T, dtype, &pl)) return 0;
    return (size_t)pl.total_segs * 2 * sizeof(float) + 16;
}

int afa_activation1d_bwd(const void* x, const void* gy, void* gx, float* galpha, float* gbeta, const float* alpha,
                         const float* beta, const float* taps_up12, const float* taps_down12, int64_t batch,
                         int64_t channels, int64_t T, int dtype, int flags, void* workspace, size_t workspace_bytes,
                         void* stream) {
    if (!x || !gy || !gx || !galpha || !alpha || !taps_up12 || !taps_down12) return fail(AFA_ERR_BAD_ARG, "null pointer argument");
    const bool snake = (flags & AFA_FLAG_SNAKE) != 0;
    if (!snake && (!beta || !gbeta)) return fail(AFA_ERR_BAD_ARG, "beta and gbeta are required unless AFA_FLAG_SNAKE is set");
    if (gx == x || gx == gy) return fail(AFA_ERR_BAD_ARG, "gx must not alias x or gy");
    Plan pl;
    if (int rc = make_plan(1, x, gy, gx, batch, channels, T, dtype, &pl)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (pl.total_segs == 0) {
        cudaError_t e = cudaMemsetAsync(galpha, 0, sizeof(float) * channels, st);
        if (e == cudaSuccess && !snake) e = cudaMemsetAsync(gbeta, 0, sizeof(float) * channels, st);
        return e == cudaSuccess ? 0 : cuda_fail(e, "cudaMemsetAsync");
    }
    const size_t need = (size_t)pl.total_segs * 2 * sizeof(float);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 3))
        return fail(AFA_ERR_WORKSPACE, "workspace of %zu bytes needed (got %zu)", need, workspace_bytes);
    afa::BwdArgs a;
    a.x = x;
    a.gy = gy;
    a.gx = gx;
    a.alpha = alpha;
    a.beta = beta;
    a.part = (float*)workspace;
    fold_bwd_taps(taps_up12, taps_down12, &a.taps);
    a.total = batch * channels * T;
    a.total_segs = pl.total_segs;
    a.nseg = pl.nseg;
    a.T = (int32_t)T;
    a.C = (int32_t)channels;
    a.flags = flags;
    int rc = dtype == AFA_DTYPE_F32 ? launch_bwd<float>(pl, a, st) : launch_bwd<__nv_bfloat16>(pl, a, st);
    if (rc) return rc;
    afa::afa_param_grad_finalize<<<(unsigned)channels, afa::kFinalizeThreads, 0, st>>>(
        a.part, galpha, gbeta, pl.total_segs, pl.nseg, (int)batch, (int)channels, snake ? 1 : 0);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_param_grad_finalize launch");
}

int afa_kernel_info(int which, int dtype, int64_t T, int32_t out[6]) {
    if (!out || which < 0 || which > 1) return fail(AFA_ERR_BAD_ARG, "bad argument");
    Plan pl;
    if (int rc = make_plan(which, nullptr, nullptr, nullptr, 1, 1, T, dtype, &pl)) return rc;
    const void* k = nullptr;
    size_t smem = 0;
#define X(CH)                                                                                                 \
    if (pl.chunks == CH) {                                                                                    \
        if (dtype == AFA_DTYPE_F32) { k = kernel_ptr<float, CH>(which, pl.aligned); smem = kernel_smem<float, CH>(which); } \
        else { k = kernel_ptr<__nv_bfloat16, CH>(which, pl.aligned); smem = kernel_smem<__nv_bfloat16, CH>(which); }        \
    }
    AFA_CHUNK_LIST(X)
#undef X
    if (!k) return fail(AFA_ERR_BAD_ARG, "no kernel for %d chunks", pl.chunks);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, k);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncGetAttributes");
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kNT, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    out[0] = fa.numRegs;
    out[1] = (int32_t)(fa.sharedSizeBytes + smem);
    out[2] = kNT;
    out[3] = pl.L;
    out[4] = occ;
    out[5] = (int32_t)g_launches.load();
    return 0;
}

}  // extern "C"
