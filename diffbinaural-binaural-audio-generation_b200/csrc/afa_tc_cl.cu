// afa_tc_cl.cu -- host side of the channels-last tensor-core Activation1d forward (afa_tc_cl_kernels.cuh): eligibility, strip
// length, 3-D tensor maps, launch.  Called from afa_amp_activation1d_fwd_cl (afa_capi.cu) for bf16 tensors without a residual
// prologue.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "afa_b200.h"
#include "afa_internal.h"
#include "afa_tc_cl_kernels.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int g_cl_enable = 1;
int g_cl_ny = 0;

// [B, T, C] bf16, channels contiguous: dims (C, T, B), box = 32 channels x 64 time steps, 64-byte swizzle, zero fill outside.
struct MapKey {
    const void* base;
    int64_t C, T, B, bs;
};
struct MapSlot {
    MapKey k;
    CUtensorMap tm;
    bool used;
};
int make_map3(CUtensorMap* tm, const void* base, int64_t C, int64_t T, int64_t B, int64_t bs) {
    static thread_local MapSlot cache[16];
    static thread_local unsigned next = 0;
    for (MapSlot& c : cache)
        if (c.used && c.k.base == base && c.k.C == C && c.k.T == T && c.k.B == B && c.k.bs == bs) {
            *tm = c.tm;
            return 0;
        }
    EncodeTiledFn fn = (EncodeTiledFn)afa_internal::tc_encode_tiled();
    if (!fn) return afa_internal::set_error(AFA_ERR_BAD_ARG, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)bs * 2};
    const cuuint32_t box[3] = {32, 64, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return afa_internal::set_error(AFA_ERR_BAD_ARG, "cuTensorMapEncodeTiled (channels-last) failed (CUresult %d)", (int)r);
    MapSlot& c = cache[next++ % 16];
    c.k = MapKey{base, C, T, B, bs};
    c.tm = *tm;
    c.used = true;
    return 0;
}

template <int kUp, int kDn>
cudaError_t prepare_kernel(int dev) {
    static bool attr_set[64] = {};
    if (dev >= 0 && dev < 64 && attr_set[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(afa_tc::afa_tc_cl_fwd_kernel<kUp, kDn>, cudaFuncAttributeMaxDynamicSharedMemorySize, afa_tc::kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(afa_tc::afa_tc_cl_fwd_kernel<kUp, kDn>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
    return cudaSuccess;
}

template <int kUp, int kDn>
cudaError_t launch_kernel(unsigned grid, cudaStream_t st, const CUtensorMap& tmx, const CUtensorMap& tmy, const afa_tc::ClArgs& a) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = prepare_kernel<kUp, kDn>(dev);
    if (e != cudaSuccess) return e;
    return afa_internal::launch_pdl(afa_tc::afa_tc_cl_fwd_kernel<kUp, kDn>, dim3(grid), dim3(afa_tc::kThreads), afa_tc::kSmemBytes, st, tmx, tmy, a);
}

// Blocks of 16 outputs per unit (NY, a multiple of 4): the cost model of afa_tc.cu's tc_plan (waves of 2 CTAs per SM, set-up +
// pipeline fill + drain ~ 5 block times per CTA).  A CTA takes four consecutive units (batch entry, strip, channel quad).
void cl_plan(int64_t batch, int64_t channels, int64_t T_out, int* ny_out, int64_t* n_cquads, int64_t* n_tstrips) {
    static int slots = 0;
    if (!slots) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = 2 * (sms > 0 ? sms : 148);
    }
    const int64_t cqs = (channels + 31) / 32, rg = batch * cqs;       // units per strip
    const int64_t tb = (T_out + 15) / 16;
    auto strips = [&](int64_t ny) { return (tb + ny - 1) / ny; };
    auto ctas = [&](int64_t ny) { return (rg * strips(ny) + 3) / 4; };
    int64_t ny = 16, best = -1;
    auto consider = [&](int64_t cand) {
        cand = (cand + 3) / 4 * 4;
        if (cand < 4) cand = 4;
        if (cand > 4096) cand = 4096;
        const int64_t cost = ((ctas(cand) + slots - 1) / slots) * (5 + cand);
        if (best < 0 || cost < best || (cost == best && cand > ny)) { best = cost; ny = cand; }
    };
    for (int cand = 4; cand <= 16; cand += 4) consider(cand);
    for (int w = 1; w <= 8; ++w) {
        const int64_t nts = (int64_t)slots * w * 4 / rg;                 // strips that fill w waves
        if (nts >= 1) consider((tb + nts - 1) / nts);
    }
    if (g_cl_ny >= 4 && g_cl_ny % 4 == 0) ny = g_cl_ny;
    *ny_out = (int)ny;
    *n_cquads = cqs;
    *n_tstrips = strips(ny);
}

}  // namespace

namespace afa_internal {

void tc_cl_set_tuning(int enable, int ny) {
    g_cl_enable = enable;
    g_cl_ny = ny;
}

// bf16, no residual prologue, 16-byte aligned rows and batch entries (C % 8 == 0, strides % 8 == 0), T % 4 == 0 (the positions
// of the replicate-pad boundary inside a 64-value block the kernel handles), at most 8 zero rows behind T, and enough work
// to fill the machine; everything else stays on afa_cl_fwd_kernel.
bool tc_cl_eligible(const void* x, int64_t x_bs, const void* res, const void* y, int64_t y_bs, int64_t y_tpad, int64_t batch,
                    int64_t channels, int64_t T, int dtype) {
    if (!g_cl_enable || !tc_mode() || dtype != AFA_DTYPE_BF16 || res != nullptr) return false;
    if (T < 64 || (T % 4) != 0 || T >= (1ll << 28) || y_tpad < T || y_tpad - T > 8) return false;
    if ((channels % 8) != 0 || (x_bs % 8) != 0 || (y_bs % 8) != 0 || channels >= (1ll << 20) || batch >= (1ll << 16)) return false;
    if ((((uintptr_t)x | (uintptr_t)y) & 15) != 0) return false;
    if (g_cl_enable == 1) {
        // units of 32 channels: a channel count that leaves more than a quarter of the lanes of its last unit idle (8, 16, 40 ...),
        // or a launch that cannot fill the machine, stays on the walk kernel
        const int64_t cqs = (channels + 31) / 32;
        if (channels * 4 < cqs * 32 * 3) return false;
        if (batch * channels * T < (4ll << 20)) return false;
    }
    return tc_encode_tiled() != nullptr;
}

int tc_cl_fwd_launch(const void* x, int64_t x_bs, const float* bias, void* y, int64_t y_bs, int64_t y_tpad, const float* alpha,
                     const float* beta, const float* taps_up12, const float* taps_down12, int64_t batch, int64_t channels,
                     int64_t T, int flags, cudaStream_t st) {
    int ny;
    int64_t cgs, ts;
    cl_plan(batch, channels, y_tpad, &ny, &cgs, &ts);
    const int64_t n_ctas = (batch * cgs * ts + 3) / 4;
    if (n_ctas >= (1ll << 29)) return set_error(AFA_ERR_TOO_LARGE, "grid of %lld CTAs", (long long)n_ctas);
    CUtensorMap tmx, tmy;
    if (int rc = make_map3(&tmx, x, channels, T, batch, x_bs)) return rc;
    if (int rc = make_map3(&tmy, y, channels, y_tpad, batch, y_bs)) return rc;
    afa_tc::ClArgs a;
    memset(&a, 0, sizeof(a));
    a.x = static_cast<const __nv_bfloat16*>(x);
    a.bias = bias;
    a.alpha = alpha;
    a.beta = beta;
    float se = 0.f, so = 0.f;
    for (int i = 0; i < 12; ++i) {
        tc_split_bf16(2.0f * taps_up12[i], &a.up_hi[i], &a.up_lo[i]);      // ratio * conv_transpose taps        resample.py:33
        tc_split_bf16(taps_down12[i], &a.dn_hi[i], &a.dn_lo[i]);
    }
    // u[n] = 2 sum_i f[n + 5 - 2 i] x[i]: n even meets the odd-index taps, n odd the even-index ones (same order of additions as
    // afa_cl_kernels.cuh's bias2)
    for (int j = 0; j < 6; ++j) { so += 2.0f * taps_up12[2 * j]; se += 2.0f * taps_up12[2 * j + 1]; }
    a.bias_even = se;
    a.bias_odd = so;
    a.x_bs = x_bs;
    a.C = (int32_t)channels;
    a.T = (int32_t)T;
    a.T_out = (int32_t)y_tpad;
    a.flags = flags;
    a.NY = ny;
    a.n_tstrips = (int32_t)ts;
    a.n_cquads = (int32_t)cgs;
    a.B = (int32_t)batch;
    const unsigned grid = (unsigned)n_ctas;
    const cudaError_t e = tc_mats() == 22 ? launch_kernel<2, 2>(grid, st, tmx, tmy, a) : launch_kernel<2, 1>(grid, st, tmx, tmy, a);
    count_launch();
    return e == cudaSuccess ? 0 : cuda_error(e, "afa_tc_cl_fwd_kernel launch");
}

int tc_cl_kernel_info(int32_t out[6]) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, afa_tc::afa_tc_cl_fwd_kernel<2, 2>);
    if (e != cudaSuccess) return cuda_error(e, "cudaFuncGetAttributes(afa_tc_cl_fwd_kernel)");
    e = prepare_kernel<2, 2>(-1);
    if (e != cudaSuccess) return cuda_error(e, "cudaFuncSetAttribute(afa_tc_cl_fwd_kernel)");
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, afa_tc::afa_tc_cl_fwd_kernel<2, 2>, afa_tc::kThreads, afa_tc::kSmemBytes);
    if (e != cudaSuccess) return cuda_error(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    out[0] = fa.numRegs;
    out[1] = (int32_t)(fa.sharedSizeBytes + afa_tc::kSmemBytes);
    out[2] = afa_tc::kThreads;
    out[3] = 32;
    out[4] = occ;
    out[5] = 0;
    return 0;
}

}  // namespace afa_internal
