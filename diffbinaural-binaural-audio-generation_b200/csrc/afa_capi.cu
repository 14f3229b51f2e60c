// afa_capi.cu -- C ABI (include/afa_b200.h) over the sm_100a kernels in afa_kernels.cuh.
// Host side only: argument checks, tap folding, kernel selection, launch on the caller's stream.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "afa_b200.h"
#include "afa_internal.h"
#include "afa_kernels.cuh"
#include "afa_cl_kernels.cuh"
#include "afa_actconv_kernels.cuh"

#ifndef AFA_CHUNK_LIST
#define AFA_CHUNK_LIST(X) X(5) X(9) X(13) X(17)
#endif

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
int g_tune_chunks[3] = {0, 0, 0};   // [2]: channels-last walk, segment length in units of 12 samples
int g_tune_threads[3] = {0, 0, 0};
int g_tail_chunks = 0;    // tail kernel: walk length 12 n + 2 (which=8; 0 = 98)
int g_actconv_xs = 1;     // fused activation+convolution, tcgen05 path: stage the input rows in shared memory (which=4)
int g_tc_mode = 1, g_tc_ny = 0, g_tc_rlog2 = -1;   // tensor-core Activation1d (which = 5, 6)
int g_actconv_path = 1;   // fused activation+convolution: 1 = tcgen05 (falls back to mma.sync when the tile does not fit), 0 = mma.sync

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

constexpr int kNW = 4;            // warps per CTA (each warp is an autonomous pipeline)
constexpr int kNT = kNW * 32;

struct Plan {
    int dtype;
    int chunks;
    int vec;
    int L;
    bool aligned;
    bool half;           // bf16 rows 8-byte but not 16-byte aligned (T % 8 == 4): TMA path with 8-byte smem chunks
    uint32_t nseg;
    uint32_t total_segs;
    uint32_t n_wtiles;
};

afa::FastDiv make_fastdiv(uint32_t d) {
    afa::FastDiv f;
    f.d = d;
    if (d <= 1) {
        f.mul = 0;
        f.shr = 0;
        return f;
    }
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg;                         // ceil(log2 d)
    const unsigned p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);         // exact for 0 <= n < 2^31
    f.shr = p - 32;
    return f;
}

afa::Geometry make_geometry(const Plan& pl, int64_t batch, int64_t channels, int64_t T, int flags);

// Persistent grid: as many CTAs as stay resident, trimmed so that every warp owns the same number of
// warp tiles (no straggler wave).
int grid_for(const void* kernel, size_t smem, uint32_t n_wtiles, uint32_t* grid) {
    struct Cache { const void* k; int dev; int ctas; };
    static thread_local Cache cache[32];
    static thread_local int n_cache = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    int ctas = 0;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].k == kernel && cache[i].dev == dev) ctas = cache[i].ctas;
    if (!ctas) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        int occ = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kNT, smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(MultiProcessorCount)");
        if (occ < 1) return fail(AFA_ERR_BAD_ARG, "kernel does not fit on an SM (smem %zu)", smem);
        ctas = occ * sms;
        if (n_cache < 32) cache[n_cache++] = Cache{kernel, dev, ctas};
    }
    const uint64_t max_warps = (uint64_t)ctas * kNW;
    const uint64_t tiles_per_warp = (n_wtiles + max_warps - 1) / max_warps;
    const uint64_t warps = (n_wtiles + tiles_per_warp - 1) / tiles_per_warp;
    *grid = (uint32_t)((warps + kNW - 1) / kNW);
    return 0;
}


// Segment length heuristic, fitted to same-box sweeps on the model's shapes (profiles/r01_fwd_sweep_segment_length.log):
// long segments amortise the warm-up steps and the per-tile bookkeeping, but every row end costs a slower
// edge-mode warp, so short rows want short segments; small launches want many small tiles.
int default_chunks(int which, int dtype, int64_t elements, int64_t T) {
    const int vec = dtype == AFA_DTYPE_F32 ? 4 : 8;
    if (which == 1) {   // backward: two staged tensors -> fewer resident warps; long segments only pay on long fp32 rows of big launches
        // (same-box sweeps, profiles/r02_bwd_sweep_segment_length.log: bf16 is fastest at 5 chunks on every shape; fp32 one-clip
        //  launches 36-38 us at 5 chunks against 43-44 at 13; fp32 (16, 384, 13776) 223 us at 5 against 238 at 9)
        if (dtype != AFA_DTYPE_F32 || elements < (16ll << 20)) return 5;
        const int64_t wtiles13 = elements / (32 * 13 * vec);
        if (wtiles13 >= 2 * 148 * 8 && T >= 32768) return 13;
        return T >= 16384 ? 9 : 5;
    }
    if (elements < (16ll << 20)) return 9;
    if (dtype == AFA_DTYPE_F32) return T < 8192 ? 13 : 17;
    if (T < 8192) return 9;
    return T < 131072 ? 13 : 17;
}

int make_plan(int which, const void* p0, const void* p1, const void* p2, int64_t batch, int64_t channels, int64_t T,
              int dtype, Plan* pl) {
    if (dtype != AFA_DTYPE_F32 && dtype != AFA_DTYPE_BF16) return fail(AFA_ERR_BAD_DTYPE, "dtype %d is not AFA_DTYPE_F32/BF16", dtype);
    if (batch < 0 || channels <= 0 || T < 0) return fail(AFA_ERR_BAD_ARG, "bad shape batch=%lld channels=%lld T=%lld", (long long)batch, (long long)channels, (long long)T);
    if (T >= (1ll << 30)) return fail(AFA_ERR_TOO_LARGE, "T=%lld exceeds 2^30", (long long)T);
    pl->dtype = dtype;
    pl->vec = dtype == AFA_DTYPE_F32 ? 4 : 8;
    int ch = g_tune_chunks[which] ? g_tune_chunks[which] : default_chunks(which, dtype, batch * channels * T, T);
    const size_t esz = dtype == AFA_DTYPE_F32 ? 4 : 2;
    const uintptr_t ptr_or = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2;
    pl->half = dtype == AFA_DTYPE_BF16 && (T % 8 == 4) && ((ptr_or & 15) == 0);
    if (pl->half) ch = 9;   // kHalfChunks
    pl->chunks = ch;
    pl->L = ch * pl->vec;
    const int64_t rows = batch * channels;
    const int64_t nseg = T > 0 ? (T + pl->L - 1) / pl->L : 0;
    const int64_t total_segs = rows * nseg;
    if (total_segs >= (1ll << 31) - 64) return fail(AFA_ERR_TOO_LARGE, "batch*channels*ceil(T/%d)=%lld exceeds 2^31", pl->L, (long long)total_segs);
    pl->nseg = (uint32_t)nseg;
    pl->total_segs = (uint32_t)total_segs;
    pl->n_wtiles = (uint32_t)((total_segs + 31) / 32);
    if (ptr_or & (esz - 1)) return fail(AFA_ERR_ALIGNMENT, "tensor pointers must be aligned to the element size");
    pl->aligned = ((T % pl->vec == 0) || pl->half) && ((ptr_or & 15) == 0);
    return 0;
}

void fold_pair_taps(const float* up, const float* dn, afa::PairTaps* t) {
    for (int j = 0; j < 6; ++j) {
        t->cu[j] = make_float2(2.0f * up[2 * j], 2.0f * up[2 * j + 1]);   // ratio * conv_transpose taps   resample.py:33
        t->cd[j] = make_float2(dn[2 * j], dn[2 * j + 1]);
    }
}
void fold_fwd_taps(const float* up, const float* dn, afa::FwdTaps* t) {
    fold_pair_taps(up, dn, &t->p);
    for (int j = 0; j < 6; ++j) {
        t->ue[j] = 2.0f * up[2 * j + 1];
        t->uo[j] = 2.0f * up[2 * j];
    }
}
void fold_bwd_taps(const float* up, const float* dn, afa::BwdTaps* t) {
    fold_pair_taps(up, dn, &t->p);
    // adjoint of the replicate pad of the activated signal (filter.py:98): taps that fall on the pad
    t->lo[0] = dn[0] + dn[1] + dn[2] + dn[3] + dn[4];
    t->lo[1] = dn[0] + dn[1] + dn[2];
    t->lo[2] = dn[0];
    t->hi[0] = dn[7] + dn[8] + dn[9] + dn[10] + dn[11];
    t->hi[1] = dn[9] + dn[10] + dn[11];
    t->hi[2] = dn[11];
}

afa::Geometry make_geometry(const Plan& pl, int64_t batch, int64_t channels, int64_t T, int flags) {
    afa::Geometry g;
    g.total = batch * channels * T;
    g.total_segs = pl.total_segs;
    g.n_wtiles = pl.n_wtiles;
    g.nseg = make_fastdiv(pl.nseg);
    g.chan = make_fastdiv((uint32_t)channels);
    g.T = (int32_t)T;
    g.flags = flags;
    g.half = pl.half ? 1 : 0;
    return g;
}

template <typename T, int CH, bool AL, bool HALF = false>
int launch_fwd_t(const afa::FwdArgs& a, cudaStream_t st) {
    auto k = afa::afa_fwd_kernel<T, CH, kNW, AL, HALF>;
    const size_t smem = afa::WarpTile<T, CH>::fwd_smem(kNW);
    uint32_t grid = 0;
    if (int rc = grid_for((const void*)k, smem, a.g.n_wtiles, &grid)) return rc;
    k<<<grid, kNT, smem, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_fwd_kernel launch");
}
template <typename T, int CH, bool AL, bool HALF = false>
int launch_bwd_t(const afa::BwdArgs& a, cudaStream_t st) {
    auto k = afa::afa_bwd_kernel<T, CH, kNW, AL, HALF>;
    const size_t smem = afa::WarpTile<T, CH>::bwd_smem(kNW);
    uint32_t grid = 0;
    if (int rc = grid_for((const void*)k, smem, a.g.n_wtiles, &grid)) return rc;
    k<<<grid, kNT, smem, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_bwd_kernel launch");
}

constexpr int kHalfChunks = 9;   // the half-aligned (bf16, T % 8 == 4) kernels are compiled for this segment size only

template <typename T>
int launch_fwd(const Plan& pl, const afa::FwdArgs& a, cudaStream_t st) {
    if constexpr (sizeof(T) == 2) {
        if (pl.half && pl.aligned) return launch_fwd_t<T, kHalfChunks, true, true>(a, st);
    }
#define X(CH)                                                                                   \
    if (pl.chunks == CH)                                                                        \
        return pl.aligned ? launch_fwd_t<T, CH, true>(a, st) : launch_fwd_t<T, CH, false>(a, st);
    AFA_CHUNK_LIST(X)
#undef X
    return fail(AFA_ERR_BAD_ARG, "no kernel compiled for %d chunks per segment", pl.chunks);
}
template <typename T>
int launch_bwd(const Plan& pl, const afa::BwdArgs& a, cudaStream_t st) {
    if constexpr (sizeof(T) == 2) {
        if (pl.half && pl.aligned) return launch_bwd_t<T, kHalfChunks, true, true>(a, st);
    }
#define X(CH)                                                                                   \
    if (pl.chunks == CH)                                                                        \
        return pl.aligned ? launch_bwd_t<T, CH, true>(a, st) : launch_bwd_t<T, CH, false>(a, st);
    AFA_CHUNK_LIST(X)
#undef X
    return fail(AFA_ERR_BAD_ARG, "no kernel compiled for %d chunks per segment", pl.chunks);
}

template <typename T, int CH>
const void* kernel_ptr(int which, bool aligned) {
    if (which == 0) return aligned ? (const void*)afa::afa_fwd_kernel<T, CH, kNW, true> : (const void*)afa::afa_fwd_kernel<T, CH, kNW, false>;
    return aligned ? (const void*)afa::afa_bwd_kernel<T, CH, kNW, true> : (const void*)afa::afa_bwd_kernel<T, CH, kNW, false>;
}
template <typename T, int CH>
size_t kernel_smem(int which) {
    return which == 0 ? afa::WarpTile<T, CH>::fwd_smem(kNW) : afa::WarpTile<T, CH>::bwd_smem(kNW);
}

}  // namespace

namespace afa_internal {
static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }
void pdl_set(int on) { g_pdl = on; }
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_error(cudaError_t e, const char* what) { return cuda_fail(e, what); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace afa_internal

extern "C" {

int afa_version(void) { return AFA_VERSION; }
const char* afa_last_error(void) { return g_err; }
int64_t afa_launch_count(void) { return (int64_t)g_launches.load(); }

int afa_set_tuning(int which, int chunks, int threads) {
    if (which == 5) {   // tensor-core Activation1d (bf16): chunks = 0 off / 1 heuristic / 2 whenever eligible; threads = forced blocks per lane (0 = heuristic)
        if (chunks < 0 || chunks > 2 || threads < 0 || threads > 4096 || threads % 4)
            return fail(AFA_ERR_BAD_ARG, "tensor-core path: mode 0..2, blocks per lane 0 or a multiple of 4 up to 4096");
        g_tc_mode = chunks;
        g_tc_ny = threads;
        afa_internal::tc_set_tuning(g_tc_mode, g_tc_ny, g_tc_rlog2);
        return 0;
    }
    if (which == 6) {   // tensor-core Activation1d: log2 of the rows per CTA (3..7; -1 = heuristic)
        if (chunks != -1 && (chunks < 3 || chunks > 7)) return fail(AFA_ERR_BAD_ARG, "log2(rows per CTA) must be 3..7 or -1");
        g_tc_rlog2 = chunks;
        afa_internal::tc_set_tuning(g_tc_mode, g_tc_ny, g_tc_rlog2);
        return 0;
    }
    if (which == 9) {   // programmatic dependent launch of the tensor-core Activation1d kernels: 1 on (default), 0 off
        if (chunks != 0 && chunks != 1) return fail(AFA_ERR_BAD_ARG, "programmatic dependent launch must be 0 or 1");
        afa_internal::pdl_set(chunks);
        return 0;
    }
    if (which == 8) {   // tail kernel: walk length = 12 * chunks + 2 samples (0 = built-in 98)
        if (chunks < 0 || chunks > 4096) return fail(AFA_ERR_BAD_ARG, "tail walk length must be 12 * [1, 4096] + 2");
        g_tail_chunks = chunks;
        return 0;
    }
    if (which == 7) {   // channels-last tensor-core Activation1d (bf16, no residual prologue): chunks = 0 off / 1 heuristic / 2 whenever eligible
        if (chunks < 0 || chunks > 2 || threads < 0 || threads > 4096 || threads % 4)
            return fail(AFA_ERR_BAD_ARG, "channels-last tensor-core path: mode 0..2, blocks per CTA 0 or a multiple of 4 up to 4096");
        afa_internal::tc_cl_set_tuning(chunks, threads);
        return 0;
    }
    if (which == 4) {   // fused activation+convolution, tcgen05 path: 1 = input rows staged by a bulk copy (default), 0 = global loads
        if (chunks != 0 && chunks != 1) return fail(AFA_ERR_BAD_ARG, "input staging must be 0 or 1");
        g_actconv_xs = chunks;
        return 0;
    }
    if (which == 3) {   // fused activation+convolution: 1 = tcgen05 (default), 0 = legacy mma.sync
        if (chunks != 0 && chunks != 1) return fail(AFA_ERR_BAD_ARG, "tensor path must be 0 (mma.sync) or 1 (tcgen05)");
        g_actconv_path = chunks;
        return 0;
    }
    if (which == 2) {   // channels-last walk: segment length = 12 * chunks + 2 samples
        if (chunks < 0 || chunks > 4096) return fail(AFA_ERR_BAD_ARG, "channels-last segment length must be 12 * [1, 4096] + 2");
        g_tune_chunks[2] = chunks;
        return 0;
    }
    if (which < 0 || which > 1) return fail(AFA_ERR_BAD_ARG, "which must be 0 (fwd), 1 (bwd), 2 (channels-last fwd), 3 / 4 (fused conv), 5 / 6 (tensor-core fwd), 7 (channels-last tensor-core fwd), 8 (tail), 9 (dependent launch)");
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
    // bf16 tensors with 16-byte aligned rows: both FIR filters run on the tensor cores (afa_tc_kernels.cuh)
    if (afa_internal::tc_eligible(x, y, batch, channels, T, dtype))
        return afa_internal::tc_fwd_launch(x, y, alpha, beta, taps_up12, taps_down12, batch, channels, T, flags,
                                           (cudaStream_t)stream, 0, nullptr);
    afa::FwdArgs a;
    a.x = x;
    a.y = y;
    a.alpha = alpha;
    a.beta = beta;
    fold_fwd_taps(taps_up12, taps_down12, &a.taps);
    a.g = make_geometry(pl, batch, channels, T, flags);
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == AFA_DTYPE_F32 ? launch_fwd<float>(pl, a, st) : launch_fwd<__nv_bfloat16>(pl, a, st);
}

int afa_activation1d_fwd_pitched(const void* x, int64_t x_row_pitch, void* y, int64_t y_row_pitch, const float* alpha,
                                 const float* beta, const float* taps_up12, const float* taps_down12, int64_t batch,
                                 int64_t channels, int64_t T, int dtype, int flags, void* stream) {
    if (!x || !y || !alpha || !taps_up12 || !taps_down12) return fail(AFA_ERR_BAD_ARG, "null pointer argument");
    if (!(flags & AFA_FLAG_SNAKE) && !beta) return fail(AFA_ERR_BAD_ARG, "beta is required unless AFA_FLAG_SNAKE is set");
    if (x == y) return fail(AFA_ERR_BAD_ARG, "y must not alias x");
    if (batch < 0 || channels < 0 || T < 0) return fail(AFA_ERR_BAD_ARG, "negative size");
    if (batch == 0 || channels == 0 || T == 0) return 0;
    if (x_row_pitch == T && y_row_pitch == T)
        return afa_activation1d_fwd(x, y, alpha, beta, taps_up12, taps_down12, batch, channels, T, dtype, flags, stream);
    if (dtype != AFA_DTYPE_BF16 && dtype != AFA_DTYPE_F32) return fail(AFA_ERR_BAD_DTYPE, "dtype must be AFA_DTYPE_F32 or AFA_DTYPE_BF16");
    if (!afa_internal::tc_pitched_ok(x, x_row_pitch, y, y_row_pitch, batch, channels, T, dtype))
        return fail(AFA_ERR_ALIGNMENT,
                    "pitched rows are served by the tensor-core kernel only: bf16, T %% 8 == 0, T >= 64, pitches >= T and multiples of 8 "
                    "elements, 16-byte aligned x and y (and afa_set_tuning(5, ...) not 0); copy to a dense tensor otherwise");
    return afa_internal::tc_fwd_launch(x, y, alpha, beta, taps_up12, taps_down12, batch, channels, T, flags, (cudaStream_t)stream,
                                       0, nullptr, x_row_pitch, y_row_pitch);
}

// slice sums [C][kFinalizeMaxSplit][2] + per-channel counters of the split second reduction stage
static size_t finalize_extra_bytes(int64_t channels) {
    return (size_t)channels * (afa::kFinalizeMaxSplit * 2 * sizeof(float) + sizeof(uint32_t));
}

size_t afa_bwd_workspace_bytes(int64_t batch, int64_t channels, int64_t T, int dtype) {
    // An upper bound over every kernel variant the call may select: the query sees no pointers, and a misaligned base
    // pointer or a forced segment size moves the launch to a variant with shorter segments (more partial sums).  The
    // shortest compiled segment is 5 chunks of 16 bytes.
    Plan pl;
    if (make_plan(1, nullptr, nullptr, nullptr, batch, channels, T, dtype, &pl)) return 0;
    const int64_t lmin = 5 * pl.vec;
    const int64_t nseg = T > 0 ? (T + lmin - 1) / lmin : 0;
    if (batch * channels * nseg == 0) return 16;
    return (size_t)(batch * channels * nseg) * 2 * sizeof(float) + finalize_extra_bytes(channels) + 16;
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
    // second reduction stage (afa_param_grad_finalize): `split` CTAs per channel, enough to fill the machine twice over
    const int64_t per_channel = batch * (int64_t)pl.nseg;
    int split = (int)((4 * 148 + channels - 1) / channels);
    if (split > afa::kFinalizeMaxSplit) split = afa::kFinalizeMaxSplit;
    while (split > 1 && per_channel / split < 16 * afa::kFinalizeThreads) --split;   // a slice must be worth its arrival round trip
    const size_t part_bytes = (size_t)pl.total_segs * 2 * sizeof(float);
    const size_t need = part_bytes + (split > 1 ? finalize_extra_bytes(channels) : 0);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 3))
        return fail(AFA_ERR_WORKSPACE, "workspace of %zu bytes needed (got %zu)", need, workspace_bytes);
    afa::BwdArgs a;
    a.x = x;
    a.gy = gy;
    a.gx = gx;
    a.alpha = alpha;
    a.beta = beta;
    a.part = (float*)workspace;
    float* part2 = split > 1 ? (float*)((char*)workspace + part_bytes) : nullptr;
    a.cnt = split > 1 ? (uint32_t*)(part2 + (size_t)channels * afa::kFinalizeMaxSplit * 2) : nullptr;
    a.n_cnt = (int32_t)channels;
    fold_bwd_taps(taps_up12, taps_down12, &a.taps);
    a.g = make_geometry(pl, batch, channels, T, flags);
    int rc = dtype == AFA_DTYPE_F32 ? launch_bwd<float>(pl, a, st) : launch_bwd<__nv_bfloat16>(pl, a, st);
    if (rc) return rc;
    afa::afa_param_grad_finalize<<<dim3((unsigned)channels, (unsigned)split), afa::kFinalizeThreads, 0, st>>>(
        a.part, galpha, gbeta, pl.total_segs, pl.nseg, (int)batch, (int)channels, snake ? 1 : 0, part2, a.cnt);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_param_grad_finalize launch");
}

int afa_kernel_info(int which, int dtype, int64_t T, int32_t out[6]) {
    return afa_kernel_info_shape(which, dtype, 1, 1, T, out);
}

int afa_kernel_info_shape(int which, int dtype, int64_t batch, int64_t channels, int64_t T, int32_t out[6]) {
    if (out && which == 5) return afa_internal::tc_kernel_info(out);   // the tensor-core forward (one variant)
    if (out && which == 7) return afa_internal::tc_cl_kernel_info(out);   // its channels-last variant
    if (!out || which < 0 || which > 1) return fail(AFA_ERR_BAD_ARG, "bad argument");
    Plan pl;
    if (int rc = make_plan(which, nullptr, nullptr, nullptr, batch, channels, T, dtype, &pl)) return rc;
    const void* k = nullptr;
    size_t smem = 0;
#define X(CH)                                                                                                 \
    if (pl.chunks == CH) {                                                                                    \
        if (dtype == AFA_DTYPE_F32) { k = kernel_ptr<float, CH>(which, pl.aligned); smem = kernel_smem<float, CH>(which); } \
        else { k = kernel_ptr<__nv_bfloat16, CH>(which, pl.aligned); smem = kernel_smem<__nv_bfloat16, CH>(which); }        \
    }
    AFA_CHUNK_LIST(X)
#undef X
    if (pl.half && pl.aligned) {
        k = which == 0 ? (const void*)afa::afa_fwd_kernel<__nv_bfloat16, kHalfChunks, kNW, true, true>
                       : (const void*)afa::afa_bwd_kernel<__nv_bfloat16, kHalfChunks, kNW, true, true>;
        smem = kernel_smem<__nv_bfloat16, kHalfChunks>(which);
    }
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


// ------------------------------------------------------------------------------------------------
// AMP-block entry points on channels-last activations (afa_cl_kernels.cuh)
// ------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// Segment length of the channels-last walk: long segments amortise the 10 extra steps (halo warm-up and
// drain) per segment; small launches want enough threads to fill 148 SMs a few times over.
// Segment length of the channels-last walk, L = 12 n + 2 (the L + 10 steps are n + 1 groups of 12).  Long
// segments amortise the 10 warm-up / drain steps, but the grid runs in waves of (resident CTAs/SM) x 148 CTAs
// whose duration grows with L, and a launch of only 2-3 long waves ends in a ragged tail (measured:
// profiles/r01_cl_sweep_*.log): long segments only where the launch still has several waves.
int cl_segment(int64_t batch, int64_t channels, int64_t T, bool res) {
    if (g_tune_chunks[2]) return 12 * g_tune_chunks[2] + 2;
    const double wave = 148.0 * (res ? 4 : 5);
    auto waves = [&](int L) { return (double)(batch * ((T + L - 1) / L) * channels) / afa::kClThreads / wave; };
    if (res) return waves(98) >= 4.0 ? 98 : (waves(74) >= 2.0 ? 74 : 50);
    if (waves(146) >= 3.0) return 146;
    return waves(98) >= 2.0 ? 98 : 50;
}

template <typename T>
int launch_cl(const afa::ClArgs& a, bool res, cudaStream_t st) {
    const uint32_t grid = (a.total + afa::kClThreads - 1) / afa::kClThreads;
    if (res) afa::afa_cl_fwd_kernel<T, true><<<grid, afa::kClThreads, 0, st>>>(a);
    else afa::afa_cl_fwd_kernel<T, false><<<grid, afa::kClThreads, 0, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_cl_fwd_kernel launch");
}

}  // namespace

extern "C" {

int afa_amp_activation1d_fwd_cl(const void* x, int64_t x_bstride, const void* res, int64_t res_bstride,
                                const float* bias, void* xsum, int64_t xsum_bstride, void* y, int64_t y_bstride,
                                int64_t y_tpad, const float* alpha, const float* beta, const float* taps_up12,
                                const float* taps_down12, int64_t batch, int64_t channels, int64_t T, int dtype,
                                int flags, void* stream) {
    if (!x || !y || !alpha || !taps_up12 || !taps_down12) return fail(AFA_ERR_BAD_ARG, "null pointer argument");
    if (!(flags & AFA_FLAG_SNAKE) && !beta) return fail(AFA_ERR_BAD_ARG, "beta is required unless AFA_FLAG_SNAKE is set");
    if (dtype != AFA_DTYPE_F32 && dtype != AFA_DTYPE_BF16) return fail(AFA_ERR_BAD_DTYPE, "dtype %d is not AFA_DTYPE_F32/BF16", dtype);
    if (batch < 0 || channels <= 0 || T < 0) return fail(AFA_ERR_BAD_ARG, "bad shape batch=%lld channels=%lld T=%lld", (long long)batch, (long long)channels, (long long)T);
    if ((xsum != nullptr) != (res != nullptr)) return fail(AFA_ERR_BAD_ARG, "res and xsum come together: xsum = x + res is the new residual stream");
    if (y == x || y == res || (xsum && (xsum == x || xsum == y || xsum == res)))
        return fail(AFA_ERR_BAD_ARG, "outputs must not alias inputs or each other (segments re-read their neighbours' halo)");
    if (y_tpad == 0) y_tpad = T;
    const int64_t row = T * channels, yrow = y_tpad * channels;
    if (y_tpad < T || x_bstride < row || (res && res_bstride < row) || (xsum && xsum_bstride < row) || y_bstride < yrow)
        return fail(AFA_ERR_BAD_ARG, "batch strides must cover T*channels elements (y: y_tpad*channels)");
    if (yrow >= (1ll << 31)) return fail(AFA_ERR_TOO_LARGE, "T*channels=%lld exceeds 2^31", (long long)yrow);
    const size_t esz = dtype == AFA_DTYPE_F32 ? 4 : 2;
    if (((uintptr_t)x | (uintptr_t)res | (uintptr_t)xsum | (uintptr_t)y) & (esz - 1)) return fail(AFA_ERR_ALIGNMENT, "tensor pointers must be aligned to the element size");
    if (batch == 0 || T == 0) return 0;
    // bf16 without a residual prologue: both FIR filters on the tensor cores (afa_tc_cl_kernels.cuh)
    if (afa_internal::tc_cl_eligible(x, x_bstride, res, y, y_bstride, y_tpad, batch, channels, T, dtype))
        return afa_internal::tc_cl_fwd_launch(x, x_bstride, bias, y, y_bstride, y_tpad, alpha, beta, taps_up12, taps_down12, batch,
                                              channels, T, flags, (cudaStream_t)stream);
    const int L = cl_segment(batch, channels, T, res != nullptr);
    const int64_t nseg = (T + L - 1) / L;
    const int64_t total = batch * nseg * channels;
    if (total >= (1ll << 31) - afa::kClThreads) return fail(AFA_ERR_TOO_LARGE, "batch*channels*ceil(T/%d)=%lld exceeds 2^31", L, (long long)total);
    afa::ClArgs a;
    a.x = x; a.res = res; a.xsum = xsum; a.y = y;
    a.bias = bias; a.alpha = alpha; a.beta = beta;
    fold_fwd_taps(taps_up12, taps_down12, &a.taps);
    a.x_bs = x_bstride; a.res_bs = res_bstride; a.xsum_bs = xsum_bstride; a.y_bs = y_bstride;
    a.total = (uint32_t)total;
    a.chan = make_fastdiv((uint32_t)channels);
    a.batch = make_fastdiv((uint32_t)batch);
    a.nseg = (uint32_t)nseg;
    a.T = (int32_t)T; a.L = L; a.y_tpad = (int32_t)y_tpad; a.flags = flags;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == AFA_DTYPE_F32 ? launch_cl<float>(a, res != nullptr, st) : launch_cl<__nv_bfloat16>(a, res != nullptr, st);
}

int afa_tail_fwd_cl(const void* x, int64_t x_bstride, const float* alpha, const float* beta, const float* taps_up12,
                    const float* taps_down12, const float* w_post, const float* bias_post, int use_tanh, float* wave,
                    int16_t* pcm, int pcm_interleave, float pcm_scale, const int32_t* frame_map, int hop, int64_t T_out,
                    int64_t batch, int64_t channels, int64_t T, int dtype, int flags, void* stream) {
    if (!x || !alpha || !taps_up12 || !taps_down12 || !w_post) return fail(AFA_ERR_BAD_ARG, "null pointer argument");
    if (!wave && !pcm) return fail(AFA_ERR_BAD_ARG, "at least one of wave / pcm is required");
    if (!(flags & AFA_FLAG_SNAKE) && !beta) return fail(AFA_ERR_BAD_ARG, "beta is required unless AFA_FLAG_SNAKE is set");
    if (dtype != AFA_DTYPE_F32 && dtype != AFA_DTYPE_BF16) return fail(AFA_ERR_BAD_DTYPE, "dtype %d is not AFA_DTYPE_F32/BF16", dtype);
    if (batch < 0 || channels <= 0 || T < 0) return fail(AFA_ERR_BAD_ARG, "bad shape batch=%lld channels=%lld T=%lld", (long long)batch, (long long)channels, (long long)T);
    if (channels > 32) return fail(AFA_ERR_BAD_ARG, "the tail kernel maps channels to the lanes of a warp: channels=%lld > 32", (long long)channels);
    if (pcm && (pcm_interleave < 1 || batch % pcm_interleave)) return fail(AFA_ERR_BAD_ARG, "batch=%lld is not a multiple of pcm_interleave=%d", (long long)batch, pcm_interleave);
    if (x_bstride < T * channels) return fail(AFA_ERR_BAD_ARG, "batch stride must cover T*channels elements");
    if (T * channels >= (1ll << 31)) return fail(AFA_ERR_TOO_LARGE, "T*channels exceeds 2^31");
    if (frame_map) {
        if (hop < 1 || T % hop) return fail(AFA_ERR_BAD_ARG, "frame_map needs T=%lld to be a whole number of hop=%d frames", (long long)T, hop);
        if (T_out < T) return fail(AFA_ERR_BAD_ARG, "T_out=%lld is shorter than T=%lld", (long long)T_out, (long long)T);
    } else {
        if (T_out == 0) T_out = T;
        if (T_out != T) return fail(AFA_ERR_BAD_ARG, "T_out must equal T without a frame_map");
        hop = 1;
    }
    if (batch == 0 || T == 0) return 0;
    const int L = g_tail_chunks ? 12 * g_tail_chunks + 2 : 98;      // walk length (12 n + 2); L - 6 outputs per segment (afa_set_tuning(8, n))
    const int64_t nseg = (T + (L - 6) - 1) / (L - 6);
    const int64_t warps = batch * nseg;
    if (warps >= (1ll << 31) / 32) return fail(AFA_ERR_TOO_LARGE, "too many segments");
    afa::TailArgs a;
    a.x = x; a.alpha = alpha; a.beta = beta; a.w = w_post; a.bias = bias_post; a.wave = wave; a.pcm = pcm;
    fold_fwd_taps(taps_up12, taps_down12, &a.taps);
    a.x_bs = x_bstride;
    a.total_warps = (uint32_t)warps;
    a.nseg = make_fastdiv((uint32_t)nseg);
    a.T = (int32_t)T; a.C = (int32_t)channels; a.L = L; a.flags = flags; a.use_tanh = use_tanh;
    a.il = pcm ? pcm_interleave : 1;
    a.pcm_scale = pcm_scale;
    a.frame_map = frame_map;
    a.hop = hop;
    a.n_frames = (int32_t)(T / hop);
    a.T_out = T_out;
    const uint32_t wpb = afa::kClThreads / 32;
    const uint32_t grid = (uint32_t)((warps + wpb - 1) / wpb);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AFA_DTYPE_F32) afa::afa_cl_tail_kernel<float><<<grid, afa::kClThreads, 0, st>>>(a);
    else afa::afa_cl_tail_kernel<__nv_bfloat16><<<grid, afa::kClThreads, 0, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_cl_tail_kernel launch");
}

int afa_resblock_mean(const void* const* xt, const void* const* xres, int num_kernels, const float* bias_sum,
                      float scale, void* out, int64_t rows, int64_t channels, int dtype, void* stream) {
    if (!xt || !xres || !out) return fail(AFA_ERR_BAD_ARG, "null pointer argument");
    if (num_kernels < 1 || num_kernels > afa::kMeanMaxK) return fail(AFA_ERR_BAD_ARG, "num_kernels must be in [1, %d]", afa::kMeanMaxK);
    if (dtype != AFA_DTYPE_F32 && dtype != AFA_DTYPE_BF16) return fail(AFA_ERR_BAD_DTYPE, "dtype %d is not AFA_DTYPE_F32/BF16", dtype);
    if (rows < 0 || channels <= 0) return fail(AFA_ERR_BAD_ARG, "bad shape rows=%lld channels=%lld", (long long)rows, (long long)channels);
    const int64_t n = rows * channels;
    if (n >= (1ll << 31)) return fail(AFA_ERR_TOO_LARGE, "rows*channels=%lld exceeds 2^31", (long long)n);
    if (n == 0) return 0;
    const int vec = dtype == AFA_DTYPE_F32 ? 4 : 8;
    uintptr_t ptr_or = (uintptr_t)out;
    afa::MeanArgs a;
    for (int j = 0; j < afa::kMeanMaxK; ++j) { a.y[j] = nullptr; a.r[j] = nullptr; }
    for (int j = 0; j < num_kernels; ++j) {
        if (!xt[j]) return fail(AFA_ERR_BAD_ARG, "null tensor pointer at index %d", j);
        a.y[j] = xt[j];
        a.r[j] = xres[j];                                   // may be null: xt[j] already contains its residual stream
        ptr_or |= (uintptr_t)xt[j] | (uintptr_t)xres[j];
    }
    const size_t esz = dtype == AFA_DTYPE_F32 ? 4 : 2;
    if (ptr_or & (esz - 1)) return fail(AFA_ERR_ALIGNMENT, "tensor pointers must be aligned to the element size");
    const bool vector = (channels % vec == 0) && ((ptr_or & 15) == 0);
    a.bias_sum = bias_sum;
    a.out = out;
    a.n = n;
    a.cvec = make_fastdiv((uint32_t)(vector ? channels / vec : channels));
    a.K = num_kernels;
    a.scale = scale;
    const int64_t nv = vector ? n / vec : n;
    const uint32_t grid = (uint32_t)((nv + 255) / 256 < 148 * 16 ? (nv + 255) / 256 : 148 * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AFA_DTYPE_F32) {
        if (vector) afa::afa_mean_kernel<float, 4><<<grid, 256, 0, st>>>(a);
        else afa::afa_mean_kernel<float, 1><<<grid, 256, 0, st>>>(a);
    } else {
        if (vector) afa::afa_mean_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>(a);
        else afa::afa_mean_kernel<__nv_bfloat16, 1><<<grid, 256, 0, st>>>(a);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_mean_kernel launch");
}


// ------------------------------------------------------------------------------------------------
// Activation1d as the prologue of the AMPBlock convolution (afa_actconv_kernels.cuh)
// ------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

template <bool RES, int NT8, int KS>
int launch_actconv_t(const afa::ActConvArgs& a, size_t smem, uint32_t grid, cudaStream_t st) {
    auto k = afa::afa_cl_actconv_kernel<RES, NT8, KS>;
    static thread_local const void* configured[8] = {nullptr};
    static thread_local int configured_dev[8] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    bool done = false;
    for (int i = 0; i < 8; ++i) done = done || (configured[i] == (const void*)k && configured_dev[i] == dev + 1);
    if (!done) {
        e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        for (int i = 0; i < 8; ++i)
            if (!configured[i]) { configured[i] = (const void*)k; configured_dev[i] = dev + 1; break; }
    }
    k<<<grid, afa::kAcThreads, smem, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_cl_actconv_kernel launch");
}

template <bool RES, int NPAD, int KS, bool XS = false>
int launch_actconv_tc_t(const afa::ActConvTcArgs& a, size_t smem, uint32_t grid, cudaStream_t st) {
    auto k = afa::afa_cl_actconv_tc_kernel<RES, NPAD, KS, XS>;
    static thread_local const void* configured[8] = {nullptr};
    static thread_local int configured_dev[8] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    bool done = false;
    for (int i = 0; i < 8; ++i) done = done || (configured[i] == (const void*)k && configured_dev[i] == dev + 1);
    if (!done) {
        e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
        for (int i = 0; i < 8; ++i)
            if (!configured[i]) { configured[i] = (const void*)k; configured_dev[i] = dev + 1; break; }
    }
    k<<<grid, afa::kAcThreads, smem, st>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "afa_cl_actconv_tc_kernel launch");
}

int launch_actconv_tc_xs(const afa::ActConvTcArgs& a, size_t smem, uint32_t grid, cudaStream_t st) {
    switch (a.a.C) {
        case 8: return launch_actconv_tc_t<false, 16, 1, true>(a, smem, grid, st);
        case 16: return launch_actconv_tc_t<false, 16, 1, true>(a, smem, grid, st);
        case 24: return launch_actconv_tc_t<false, 32, 2, true>(a, smem, grid, st);
        case 32: return launch_actconv_tc_t<false, 32, 2, true>(a, smem, grid, st);
        case 48: return launch_actconv_tc_t<false, 48, 3, true>(a, smem, grid, st);
        case 64: return launch_actconv_tc_t<false, 64, 4, true>(a, smem, grid, st);
    }
    return fail(AFA_ERR_BAD_ARG, "fused activation+convolution is compiled for channels in {8,16,24,32,48,64}, got %d", a.a.C);
}

template <bool RES>
int launch_actconv_tc(const afa::ActConvTcArgs& a, size_t smem, uint32_t grid, cudaStream_t st) {
    switch (a.a.C) {
        case 8: return launch_actconv_tc_t<RES, 16, 1>(a, smem, grid, st);
        case 16: return launch_actconv_tc_t<RES, 16, 1>(a, smem, grid, st);
        case 24: return launch_actconv_tc_t<RES, 32, 2>(a, smem, grid, st);
        case 32: return launch_actconv_tc_t<RES, 32, 2>(a, smem, grid, st);
        case 48: return launch_actconv_tc_t<RES, 48, 3>(a, smem, grid, st);
        case 64: return launch_actconv_tc_t<RES, 64, 4>(a, smem, grid, st);
    }
    return fail(AFA_ERR_BAD_ARG, "fused activation+convolution is compiled for channels in {8,16,24,32,48,64}, got %d", a.a.C);
}


template <bool RES>
int launch_actconv(const afa::ActConvArgs& a, size_t smem, uint32_t grid, cudaStream_t st) {
    switch (a.C) {
        case 8: return launch_actconv_t<RES, 1, 1>(a, smem, grid, st);
        case 16: return launch_actconv_t<RES, 2, 1>(a, smem, grid, st);
        case 24: return launch_actconv_t<RES, 3, 2>(a, smem, grid, st);
        case 32: return launch_actconv_t<RES, 4, 2>(a, smem, grid, st);
        case 48: return launch_actconv_t<RES, 6, 3>(a, smem, grid, st);
        case 64: return launch_actconv_t<RES, 8, 4>(a, smem, grid, st);
    }
    return fail(AFA_ERR_BAD_ARG, "fused activation+convolution is compiled for channels in {8,16,24,32,48,64}, got %d", a.C);
}

}  // namespace

extern "C" {

int afa_amp_act_conv_supported(int64_t channels, int kernel_size, int dilation, int dtype) {
    if (dtype != AFA_DTYPE_BF16) return 0;
    if (!(channels == 8 || channels == 16 || channels == 24 || channels == 32 || channels == 48 || channels == 64)) return 0;
    if (kernel_size < 1 || kernel_size > 11 || !(kernel_size & 1) || dilation < 1 || dilation > 5) return 0;
    return 1;
}

int afa_amp_act_conv_fwd_cl(const void* x, int64_t x_bstride, const void* res, int64_t res_bstride, const float* bias,
                            void* xsum, int64_t xsum_bstride, const void* addend, int64_t addend_bstride, void* y,
                            int64_t y_bstride, const float* alpha, const float* beta, const float* taps_up12,
                            const float* taps_down12, const void* w_kcc, int kernel_size, int dilation, int64_t batch,
                            int64_t channels, int64_t T, int dtype, int flags, void* stream) {
    if (!x || !y || !alpha || !taps_up12 || !taps_down12 || !w_kcc) return fail(AFA_ERR_BAD_ARG, "null pointer argument");
    if (!(flags & AFA_FLAG_SNAKE) && !beta) return fail(AFA_ERR_BAD_ARG, "beta is required unless AFA_FLAG_SNAKE is set");
    if (dtype != AFA_DTYPE_BF16) return fail(AFA_ERR_BAD_DTYPE, "the fused activation+convolution runs on bf16 activations (tensor-core path); dtype %d", dtype);
    if (batch < 0 || channels <= 0 || T < 0) return fail(AFA_ERR_BAD_ARG, "bad shape batch=%lld channels=%lld T=%lld", (long long)batch, (long long)channels, (long long)T);
    if (!afa_amp_act_conv_supported(channels, kernel_size, dilation, dtype))
        return fail(AFA_ERR_BAD_ARG, "unsupported configuration: channels=%lld (need one of 8,16,24,32,48,64), kernel_size=%d (odd, <= 11), dilation=%d (<= 5)", (long long)channels, kernel_size, dilation);
    if ((xsum != nullptr) != (res != nullptr)) return fail(AFA_ERR_BAD_ARG, "res and xsum come together: xsum = x + res is the new residual stream");
    if (y == x || y == res || (xsum && (xsum == x || xsum == y || xsum == res)))
        return fail(AFA_ERR_BAD_ARG, "outputs must not alias inputs or each other (tiles re-read their neighbours' halo)");
    const int64_t row = T * channels;
    if (x_bstride < row || (res && res_bstride < row) || (xsum && xsum_bstride < row) || y_bstride < row || (addend && addend_bstride < row))
        return fail(AFA_ERR_BAD_ARG, "batch strides must cover T*channels elements");
    if (addend && (addend == y || addend == xsum)) return fail(AFA_ERR_BAD_ARG, "addend must not alias an output");
    if (addend && (((uintptr_t)addend & 15) || (addend_bstride % 8))) return fail(AFA_ERR_ALIGNMENT, "addend must be 16-byte aligned (batch stride a multiple of 8)");
    if (row >= (1ll << 31)) return fail(AFA_ERR_TOO_LARGE, "T*channels=%lld exceeds 2^31", (long long)row);
    if (((uintptr_t)x | (uintptr_t)res | (uintptr_t)xsum) & 1) return fail(AFA_ERR_ALIGNMENT, "tensor pointers must be aligned to the element size");
    if (((uintptr_t)y & 3) || ((uintptr_t)w_kcc & 15) || (y_bstride & 1)) return fail(AFA_ERR_ALIGNMENT, "y must be 4-byte aligned (even batch stride), the weights 16-byte aligned");
    if (batch == 0 || T == 0) return 0;
    const int C = (int)channels;
    afa::ActConvArgs a;
    a.x = x; a.res = res; a.xsum = xsum; a.y = y; a.bias = bias; a.alpha = alpha; a.beta = beta;
    a.w = (const __nv_bfloat16*)w_kcc;
    a.addend = addend; a.addend_bs = addend_bstride;
    fold_fwd_taps(taps_up12, taps_down12, &a.taps);
    a.x_bs = x_bstride; a.res_bs = res_bstride; a.xsum_bs = xsum_bstride; a.y_bs = y_bstride;
    a.T = (int32_t)T; a.C = C; a.flags = flags; a.batch = (int32_t)batch;
    a.k = kernel_size; a.dil = dilation;
    const int P = (kernel_size / 2) * dilation;
    a.n_sub = afa::kAcThreads / C;                                   // (sub-segment, channel) pairs fill the CTA
    if (a.n_sub > 12) a.n_sub = 12;
    const int cpad = (C + 15) / 16 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    // tcgen05 path: column-strip tile, all blocks of 128 output rows issued up front into TMEM
    if (g_actconv_path == 1 && (((uintptr_t)y & 15) == 0) && (y_bstride % 8 == 0)) {
        const int npad = C <= 16 ? 16 : cpad;
        const int nstrip = cpad / 8;
        const size_t wb = (size_t)kernel_size * nstrip * npad * 16;
        // without a residual prologue the tile's input rows are staged in shared memory by one bulk copy (XS variant) --
        // unless the extra buffer would shrink the tile by more than a fifth (wide weights: C = 48, k = 11)
        const bool xs_ok = g_actconv_xs && !res && (((uintptr_t)x & 15) == 0) && (x_bstride % 8 == 0);
        struct Cfg { int lsub, a_rows, tt, n_mb, rows_alloc, xin_rows; size_t smem; bool ok; };
        auto pick = [&](bool xs) {
            Cfg c{};
            for (int n = 7; n >= 3; --n) {
                const int lsub = 12 * n + 2;
                const int a_rows = a.n_sub * lsub;
                if ((int64_t)a_rows - 2 * P > 2 * T && n > 3) continue;       // short rows: keep enough tiles
                const int tt = (a_rows - 2 * P) / 16 * 16;
                if (tt < 16) break;
                const int n_mb = (tt + 127) / 128;
                const int rows_alloc = (n_mb * 128 + 2 * P > a_rows ? n_mb * 128 + 2 * P : a_rows) | 1;   // MMA reach and phase-1 rows
                const int xin_rows = xs ? a_rows + 11 : 0;
                const size_t smem_tc = 128 + wb + (size_t)nstrip * rows_alloc * 16 + (size_t)xin_rows * C * 2;
                if (n_mb > afa::kAcMaxBlocks || n_mb * npad > 512 || smem_tc > 110 * 1024) continue;
                c = Cfg{lsub, a_rows, tt, n_mb, rows_alloc, xin_rows, smem_tc, true};
                break;
            }
            return c;
        };
        Cfg cfg = pick(false);
        bool xs = false;
        if (xs_ok) {
            const Cfg cx = pick(true);
            if (cx.ok && (!cfg.ok || 5 * cx.tt >= 4 * cfg.tt)) { cfg = cx; xs = true; }
        }
        if (cfg.ok) {
            afa::ActConvTcArgs ta;
            a.Lsub = cfg.lsub; a.a_rows = cfg.a_rows; a.TT = cfg.tt; a.a_stride = 8; a.w_stride = 8;
            a.n_tiles = (int32_t)((T + cfg.tt - 1) / cfg.tt);
            ta.a = a;
            ta.rows_alloc = cfg.rows_alloc;
            ta.n_mb = cfg.n_mb;
            ta.xin_rows = cfg.xin_rows;
            int cols = 32;
            while (cols < cfg.n_mb * npad) cols *= 2;
            ta.tmem_cols = cols;
            const int64_t grid_tc = (int64_t)a.n_tiles * batch;
            if (grid_tc >= (1ll << 31)) return fail(AFA_ERR_TOO_LARGE, "too many tiles");
            if (xs) return launch_actconv_tc_xs(ta, cfg.smem, (uint32_t)grid_tc, st);
            return res ? launch_actconv_tc<true>(ta, cfg.smem, (uint32_t)grid_tc, st) : launch_actconv_tc<false>(ta, cfg.smem, (uint32_t)grid_tc, st);
        }
    }
    // legacy mma.sync path: row-major tiles, rows padded by 8 elements (conflict-free ldmatrix)
    a.a_stride = cpad + 8;
    a.w_stride = cpad + 8;
    const size_t w_bytes = (size_t)kernel_size * C * a.w_stride * 2;
    int n = 7;
    while (n > 2 && w_bytes + (size_t)a.n_sub * (12 * n + 2) * a.a_stride * 2 > 100 * 1024) --n;
    while (n > 2 && (int64_t)a.n_sub * (12 * n + 2) - 2 * P > 2 * T) --n;
    a.Lsub = 12 * n + 2;
    a.a_rows = a.n_sub * a.Lsub;
    a.TT = (a.a_rows - 2 * P) / 16 * 16;
    if (a.TT < 16) return fail(AFA_ERR_BAD_ARG, "tile too small for kernel_size=%d dilation=%d at channels=%d", kernel_size, dilation, C);
    a.n_tiles = (int32_t)((T + a.TT - 1) / a.TT);
    const size_t smem = w_bytes + (size_t)a.a_rows * a.a_stride * 2;
    const int64_t grid = (int64_t)a.n_tiles * batch;
    if (grid >= (1ll << 31)) return fail(AFA_ERR_TOO_LARGE, "too many tiles");
    return res ? launch_actconv<true>(a, smem, (uint32_t)grid, st) : launch_actconv<false>(a, smem, (uint32_t)grid, st);
}

}  // extern "C"
