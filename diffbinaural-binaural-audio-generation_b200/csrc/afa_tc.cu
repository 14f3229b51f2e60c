// afa_tc.cu -- host side of the tensor-core Activation1d kernels (afa_tc_kernels.cuh): eligibility, tile-shape
// choice, tensor-map encoding (driver entry point fetched through the runtime: no link against libcuda), launch.
// Called from afa_activation1d_fwd (afa_capi.cu) for bf16 tensors whose rows are 16-byte aligned.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "afa_b200.h"
#include "afa_internal.h"
#include "afa_tc_kernels.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

int g_tc_enable = 1;      // afa_set_tuning(5, ...): 0 = never take the tensor-core path
int g_tc_ny = 0;          // forced y blocks per lane (0 = heuristic)
int g_tc_rlog2 = -1;      // forced log2(rows per CTA) (-1 = heuristic)
int g_tc_mats = 22;       // tap matrices per K slice, up * 10 + down: 2 = bf16 hi + lo (16 mantissa bits), 1 = taps rounded to bf16
int g_tc_dbg_j0 = 0;      // harness: first block of the clock-stamp window

// [rows, T] bf16 row-major, box = R rows x 64 samples, 128-byte swizzle, zero fill outside the tensor.  Callers with
// T % 8 == 4 pass PAIRS of rows as one map row (rows / 2, 2 T): the row pitch of a tensor map is a multiple of 16 bytes.
// A map depends on (base, rows, T, R) only: the last few are kept per host thread, so an eager generator pass -- the same
// buffers call after call -- pays the driver's encode (~1 us each, two per launch) once.
struct MapKey {
    const void* base;
    int64_t rows, T, pitch;
    int R;
};
struct MapSlot {
    MapKey k;
    CUtensorMap tm;
    bool used;
};
int make_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t T, int64_t pitch, int R) {
    static thread_local MapSlot cache[16];
    static thread_local unsigned next = 0;
    for (MapSlot& c : cache)
        if (c.used && c.k.base == base && c.k.rows == rows && c.k.T == T && c.k.pitch == pitch && c.k.R == R) {
            *tm = c.tm;
            return 0;
        }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return afa_internal::set_error(AFA_ERR_BAD_ARG, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)T, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)R};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return afa_internal::set_error(AFA_ERR_BAD_ARG, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    MapSlot& c = cache[next++ % 16];
    c.k = MapKey{base, rows, T, pitch, R};
    c.tm = *tm;
    c.used = true;
    return 0;
}

uint16_t bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
// v ~= hi + lo with hi, lo bf16: 16 mantissa bits for the filter taps, both halves ride the same fp32 accumulation
void split_bf16(float v, uint16_t* hi, uint16_t* lo) {
    *hi = bf16_rne(v);
    const uint32_t hb = (uint32_t)*hi << 16;
    float hf;
    memcpy(&hf, &hb, 4);
    *lo = bf16_rne(v - hf);
}

template <int kUp, int kDn, bool kDebug>
cudaError_t prepare_kernel(int dev) {
    static bool attr_set[64] = {};
    if (dev >= 0 && dev < 64 && attr_set[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(afa_tc::afa_tc_fwd_kernel<kUp, kDn, kDebug>, cudaFuncAttributeMaxDynamicSharedMemorySize, afa_tc::kSmemBytes);
    if (e != cudaSuccess) return e;
    // two CTAs per SM need the largest shared-memory carveout (2 x 88 KB)
    e = cudaFuncSetAttribute(afa_tc::afa_tc_fwd_kernel<kUp, kDn, kDebug>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
    return cudaSuccess;
}

template <int kUp, int kDn, bool kDebug>
cudaError_t launch_kernel(unsigned grid, cudaStream_t st, const CUtensorMap& tmx, const CUtensorMap& tmy, const afa_tc::Args& a) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = prepare_kernel<kUp, kDn, kDebug>(dev);
    if (e != cudaSuccess) return e;
    return afa_internal::launch_pdl(afa_tc::afa_tc_fwd_kernel<kUp, kDn, kDebug>, dim3(grid), dim3(afa_tc::kThreads), afa_tc::kSmemBytes, st, tmx, tmy, a);
}

}  // namespace

namespace afa_internal {

void* tc_encode_tiled() { return (void*)encode_fn(); }
void tc_split_bf16(float v, uint16_t* hi, uint16_t* lo) { split_bf16(v, hi, lo); }
int tc_mats() { return g_tc_mats; }
int tc_mode() { return g_tc_enable; }

void tc_set_mats(int mats) { g_tc_mats = (mats == 11 || mats == 12 || mats == 21) ? mats : 22; }
void tc_set_debug_window(int j0) { g_tc_dbg_j0 = j0; }

void tc_set_tuning(int enable, int ny, int rlog2) {
    g_tc_enable = enable;
    g_tc_ny = ny;
    g_tc_rlog2 = rlog2;
}

// The tensor-core path takes bf16 tensors whose rows start on 16-byte boundaries (tensor-map TMA) and that are
// large enough to fill the machine; everything else stays on the register-walk kernels of afa_kernels.cuh.
bool tc_eligible(const void* x, const void* y, int64_t batch, int64_t channels, int64_t T, int dtype) {
    if (!g_tc_enable || dtype != AFA_DTYPE_BF16) return false;
    if (T < 64 || (T % 4) != 0 || T >= (1ll << 28)) return false;
    if ((((uintptr_t)x | (uintptr_t)y) & 15) != 0) return false;
    const int64_t rows = batch * channels;
    if (rows < 8 || rows >= (1ll << 30)) return false;
    if ((T % 8) != 0 && (rows % 2) != 0) return false;       // rows of 8-byte alignment travel as 16-byte aligned pairs
    if (g_tc_enable == 1) {
        // Built-in choice, fitted to same-box sustained sweeps against the register-walk kernel
        // (profiles/r02_tc_sweep_v10_sustained.log): this kernel wins by 30-70 % on eight-clip launches (3.8-4.2 against
        // 2.2-3.1 TB/s), by 25-30 % on the training shapes (batch 32) and by 3-25 % on single-clip launches; the walk kernel
        // keeps the short launches (< 6 M elements, where this kernel's set-up and pipeline fill weigh most) and rows shorter
        // than 256 samples.
        const int64_t n = rows * T;
        if (T < 256 || n < (6ll << 20)) return false;
    }
    return encode_fn() != nullptr;
}

void tc_plan(int64_t rows_in, int64_t T, int* rlog2_out, int* ny_out, int64_t* n_rgroups, int64_t* n_tstrips) {
    const int64_t halves = (T % 8) != 0 ? 2 : 1;               // tensor rows per tensor-map row
    const int64_t rows = rows_in / halves;
    int rlog2 = 3;
    for (int cand = 7; cand >= 3; --cand) {
        const int64_t R = 1ll << cand;
        const int64_t padded = (rows + R - 1) / R * R;
        if (padded * 10 <= rows * 11) { rlog2 = cand; break; }
    }
    if (g_tc_rlog2 >= 3 && g_tc_rlog2 <= 7) rlog2 = g_tc_rlog2;
    const int64_t R = 1ll << rlog2, G = 128 / R;
    const int64_t rg = (rows + R - 1) / R;
    // Blocks of 16 outputs per lane and CTA (NY, a multiple of 4).  A CTA streams its strip through a recycled chunk ring, so a
    // strip may be any length: its set-up, pipeline fill and drain cost ~5 block-times once, and the grid runs in waves of
    // 2 CTAs per SM.  Candidates: the short strips (4 ... 16) and the strip lengths that make the grid exactly w waves; the
    // cheapest under  waves * (5 + NY)  wins, the longer strip on ties (profiles/r02_tc_sweep_v8_sustained_ny.log,
    // profiles/r02_tc_timeline_power.txt section 7).
    static int slots = 0;
    if (!slots) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = 2 * (sms > 0 ? sms : 148);
    }
    const int64_t tb = (T + 4 * (halves - 1) + 15) / 16;               // blocks per row (the second row of a pair starts 4 samples early)
    auto strips = [&](int64_t ny) { return (halves * ((tb + ny - 1) / ny) + G - 1) / G; };      // CTAs per row group: G lane strips each
    int64_t ny = 16, best = -1;
    auto consider = [&](int64_t cand) {
        cand = (cand + 3) / 4 * 4;
        if (cand < 4) cand = 4;
        if (cand > 4096) cand = 4096;
        const int64_t cost = ((rg * strips(cand) + slots - 1) / slots) * (5 + cand);
        if (best < 0 || cost < best || (cost == best && cand > ny)) { best = cost; ny = cand; }
    };
    for (int cand = 4; cand <= 16; cand += 4) consider(cand);
    for (int w = 1; w <= 8; ++w) {
        const int64_t nts = (int64_t)slots * w / rg;                   // strips per row group that fill w waves
        if (nts >= 1) consider((halves * tb + G * nts - 1) / (G * nts));
    }
    if (g_tc_ny >= 4 && g_tc_ny % 4 == 0) ny = g_tc_ny;
    *rlog2_out = rlog2;
    *ny_out = (int)ny;
    *n_rgroups = rg;
    *n_tstrips = strips(ny);
}

// Rows `x_pitch` / `y_pitch` elements apart (0 = dense): the tensor maps carry the pitch, nothing else changes.
bool tc_pitched_ok(const void* x, int64_t x_pitch, const void* y, int64_t y_pitch, int64_t batch, int64_t channels, int64_t T,
                   int dtype) {
    if (!g_tc_enable || dtype != AFA_DTYPE_BF16 || encode_fn() == nullptr) return false;
    if (T < 64 || (T % 8) != 0 || T >= (1ll << 28)) return false;
    if ((((uintptr_t)x | (uintptr_t)y) & 15) != 0) return false;
    if (x_pitch < T || y_pitch < T || (x_pitch % 8) != 0 || (y_pitch % 8) != 0 || x_pitch >= (1ll << 38) || y_pitch >= (1ll << 38)) return false;
    const int64_t rows = batch * channels;
    return rows >= 1 && rows < (1ll << 30);
}

int tc_fwd_launch(const void* x, void* y, const float* alpha, const float* beta, const float* taps_up12,
                  const float* taps_down12, int64_t batch, int64_t channels, int64_t T, int flags, cudaStream_t st,
                  int debug, float* dbg, int64_t x_pitch, int64_t y_pitch) {
    if (x_pitch <= 0) x_pitch = T;
    if (y_pitch <= 0) y_pitch = T;
    const int64_t rows = batch * channels;
    int rlog2, ny;
    int64_t rg, ts;
    tc_plan(rows, T, &rlog2, &ny, &rg, &ts);
    if (rg * ts >= (1ll << 31)) return set_error(AFA_ERR_TOO_LARGE, "grid of %lld CTAs", (long long)(rg * ts));
    CUtensorMap tmx, tmy;
    const int64_t halves = (T % 8) != 0 ? 2 : 1;
    if (int rc = make_map(&tmx, x, rows / halves, T * halves, x_pitch * halves, 1 << rlog2)) return rc;
    if (int rc = make_map(&tmy, y, rows / halves, T * halves, y_pitch * halves, 1 << rlog2)) return rc;
    afa_tc::Args a;
    memset(&a, 0, sizeof(a));
    a.x = static_cast<const __nv_bfloat16*>(x);
    a.y = static_cast<__nv_bfloat16*>(y);
    a.x_pitch = x_pitch;
    a.y_pitch = y_pitch;
    a.halves = (int32_t)halves;
    a.spr = (int32_t)(((T + 4 * (halves - 1) + 15) / 16 + ny - 1) / ny);      // the second row of a pair is covered 4 samples early
    a.alpha = alpha;
    a.beta = beta;
    const int mats = g_tc_mats;
    for (int i = 0; i < 12; ++i) {
        split_bf16(2.0f * taps_up12[i], &a.up_hi[i], &a.up_lo[i]);      // ratio * conv_transpose taps        resample.py:33
        split_bf16(taps_down12[i], &a.dn_hi[i], &a.dn_lo[i]);
    }
    a.rows = (int32_t)rows;
    a.C = (int32_t)channels;
    a.T = (int32_t)T;
    a.flags = flags;
    a.R_log2 = rlog2;
    a.NY = ny;
    a.n_tstrips = (int32_t)ts;
    a.debug = debug;
    a.dbg = dbg;
    a.dbg_cta = (int32_t)(rg * ts / 2);
    a.dbg_blocks = ny / 2 + 1;
    a.dbg_j0 = g_tc_dbg_j0;
    cudaError_t e;
    const unsigned grid = (unsigned)(rg * ts);
#ifdef AFA_TC_HARNESS
    if (debug) e = mats == 22 ? launch_kernel<2, 2, true>(grid, st, tmx, tmy, a) : launch_kernel<2, 1, true>(grid, st, tmx, tmy, a);
    else if (mats == 11) e = launch_kernel<1, 1, false>(grid, st, tmx, tmy, a);
    else if (mats == 12) e = launch_kernel<1, 2, false>(grid, st, tmx, tmy, a);
    else
#endif
        e = mats == 22 ? launch_kernel<2, 2, false>(grid, st, tmx, tmy, a) : launch_kernel<2, 1, false>(grid, st, tmx, tmy, a);
    count_launch();
    return e == cudaSuccess ? 0 : cuda_error(e, "afa_tc_fwd_kernel launch");
}

template <int kUp, int kDn>
static int kernel_info_t(int32_t out[6]) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, afa_tc::afa_tc_fwd_kernel<kUp, kDn, false>);
    if (e != cudaSuccess) return cuda_error(e, "cudaFuncGetAttributes(afa_tc_fwd_kernel)");
    e = prepare_kernel<kUp, kDn, false>(-1);
    if (e != cudaSuccess) return cuda_error(e, "cudaFuncSetAttribute(afa_tc_fwd_kernel)");
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, afa_tc::afa_tc_fwd_kernel<kUp, kDn, false>, afa_tc::kThreads, afa_tc::kSmemBytes);
    if (e != cudaSuccess) return cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    out[0] = fa.numRegs;
    out[1] = (int32_t)(fa.sharedSizeBytes + afa_tc::kSmemBytes);
    out[2] = afa_tc::kThreads;
    out[3] = 32;                     // outputs per block; blocks per lane are chosen per launch
    out[4] = occ;
    out[5] = 0;
    return 0;
}
int tc_kernel_info(int32_t out[6]) { return g_tc_mats == 22 ? kernel_info_t<2, 2>(out) : kernel_info_t<2, 1>(out); }

}  // namespace afa_internal
