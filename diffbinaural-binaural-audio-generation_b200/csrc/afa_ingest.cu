// afa_ingest.cu -- zero-frame compaction of a batch of mel spectrograms on the device.
//
// Reference: BigVGAN/inference_e2e.py:38-74 (detect_and_exclude_zero_frames), called per file and per channel on the host
// (:146-147): frames whose sum of |mel| over the bands is <= 1e-10 are dropped before the generator and come back as silence
// (:77-111, which the fused tail kernel does by scattering hops through a frame map: afa_tail_fwd_cl).  Here the detection,
// the left-packing gather and the frame map of a whole batch are one launch: one CTA per row (clip x channel), no
// per-frame copies, no batched-memcpy API.
//   frame_sums[f] = sum_m |mel[m][f]|   -- the same left-to-right float32 order numpy uses for an axis-0 reduction of a
//                                          C-contiguous [n_mels, T] array, so the mask is bit-identical to the reference's
//   kept frame number p of original frame f: packed[m][p] = mel[m][f], frame_map[p] = f;  n_kept = number of kept frames
#include <cuda_runtime.h>
#include <stdint.h>

#include "afa_b200.h"
#include "afa_internal.h"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) afa_compact_frames_kernel(const float* __restrict__ mel, float* __restrict__ packed,
                                                                      int32_t* __restrict__ frame_map, int32_t* __restrict__ n_kept,
                                                                      int n_mels, int T, float thr) {
    __shared__ int warp_sums[kThreads / 32];
    __shared__ int carry_s;
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* src = mel + (size_t)row * n_mels * T;
    float* dst = packed + (size_t)row * n_mels * T;
    int32_t* fm = frame_map + (size_t)row * T;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int f0 = 0; f0 < T; f0 += kThreads) {
        const int f = f0 + tid;
        int keep = 0;
        if (f < T) {
            float s = 0.f;
            for (int m = 0; m < n_mels; ++m) s += fabsf(__ldg(src + (size_t)m * T + f));
            keep = s > thr ? 1 : 0;                                  // zero_mask = frame_sums <= zero_threshold
        }
        // exclusive scan of the keep flags over the 256 frames of this pass
        int incl = keep;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        int base = carry_s;
        for (int w = 0; w < warp; ++w) base += warp_sums[w];
        const int p = base + incl - keep;
        if (keep) {
            fm[p] = f;
            for (int m = 0; m < n_mels; ++m) dst[(size_t)m * T + p] = __ldg(src + (size_t)m * T + f);
        }
        __syncthreads();
        if (tid == kThreads - 1) carry_s = base + incl;
        __syncthreads();
    }
    const int total = carry_s;
    for (int p = total + tid; p < T; p += kThreads) {
        fm[p] = -1;                                                   // beyond the kept frames: no frame (the tail kernel drops it)
        for (int m = 0; m < n_mels; ++m) dst[(size_t)m * T + p] = 0.f;
    }
    if (tid == 0) n_kept[row] = total;
}

}  // namespace

extern "C" int afa_compact_zero_frames(const float* mel, float* packed, int32_t* frame_map, int32_t* n_kept, int64_t rows,
                                       int n_mels, int64_t T, float zero_threshold, void* stream) {
    if (rows < 0 || n_mels <= 0 || T < 0) return afa_internal::set_error(AFA_ERR_BAD_ARG, "bad shape rows=%lld n_mels=%d T=%lld", (long long)rows, n_mels, (long long)T);
    if (rows >= (1ll << 31) || T >= (1ll << 31)) return afa_internal::set_error(AFA_ERR_TOO_LARGE, "rows / T exceed 2^31");
    if (rows == 0) return 0;                                  // an empty batch is a no-op (its pointers may be null)
    if (!n_kept || (T > 0 && (!mel || !packed || !frame_map))) return afa_internal::set_error(AFA_ERR_BAD_ARG, "null pointer argument");
    if (mel == packed && T > 0) return afa_internal::set_error(AFA_ERR_BAD_ARG, "packed must not alias mel");
    if (T == 0) {
        cudaError_t e = cudaMemsetAsync(n_kept, 0, sizeof(int32_t) * rows, (cudaStream_t)stream);
        return e == cudaSuccess ? 0 : afa_internal::cuda_error(e, "cudaMemsetAsync");
    }
    afa_compact_frames_kernel<<<(unsigned)rows, kThreads, 0, (cudaStream_t)stream>>>(mel, packed, frame_map, n_kept, n_mels, (int)T, zero_threshold);
    afa_internal::count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : afa_internal::cuda_error(e, "afa_compact_frames_kernel launch");
}
