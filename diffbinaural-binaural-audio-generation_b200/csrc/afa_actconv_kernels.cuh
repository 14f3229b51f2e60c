// afa_actconv_kernels.cuh -- Activation1d as the PROLOGUE of the AMPBlock's dilated Conv1d, for the narrow
// stages of the generator (channels <= 64: C = 48 and C = 24 in bigvgan_binaural_22khz_80band_256x), sm_100a.
//
//   y = conv1d_{k, dilation}( down2x(snake(up2x(x (+ res) + bias[c]))) )          (no bias on y: it stays pending)
//   reference: `xt = a(x); xt = c(xt)`  BigVGAN/bigvgan.py:134-138 (AMPBlock1), :234-235 (AMPBlock2), with the
//   folded bias / residual prologue of afa_cl_fwd_kernel (afa_cl_kernels.cuh).
//
// Why: with 24 or 48 channels the convolution is a [T x (k*C)] x [(k*C) x C] product with almost no arithmetic
// per byte; cuDNN runs it at 3-10x its memory time (measured 133-307 us for the 85 MB + 85 MB of one C = 24
// call, gpurun_out/conv_cl.log), while the activation in front of it is FP32-pipe bound and leaves the tensor
// pipe idle.  Fused, the activated tile never leaves the SM and the convolution rides along:
//
//   * a CTA owns TT output rows of one batch entry.  Phase 1: its 256 threads = (sub-segment, channel) pairs run
//     the channels-last activation walk (walk_cl, SINK 2) over the TT + 2*(k/2)*dilation rows the convolution
//     reaches and write them as bf16 into a shared-memory tile [rows][C padded to 16] (zero outside the row =
//     the convolution's zero padding).  Phase 2: the 8 warps sweep the tile with legacy tensor-core MMAs
//     (mma.sync.m16n8k16 bf16 -> f32: M = 16 time steps, N = 8 output channels, K = 16 input channels, one
//     accumulation chain over the k taps), operands via ldmatrix, weights staged once per CTA in shared
//     memory as [tap][c_out][c_in].  tcgen05 is deliberately not used here: the product is 24-48 wide, the
//     FP32 pipe of phase 1 is the bound, and mma.sync needs no TMEM / descriptor machinery for a side job.
//   * several CTAs per SM overlap one CTA's phase 1 (FMA pipe) with another's phase 2 (tensor pipe).
#pragma once
#include "afa_cl_kernels.cuh"

namespace afa {

constexpr int kAcThreads = 256;

struct ActConvArgs {
    const void* x;
    const void* res;
    void* xsum;
    void* y;
    const float* bias;
    const float* alpha;
    const float* beta;
    const __nv_bfloat16* w;     // [k][C][C] = [tap][c_out][c_in] bf16
    FwdTaps taps;
    int64_t x_bs, res_bs, xsum_bs, y_bs;
    int32_t T, C, flags, batch;
    int32_t k, dil;             // kernel size (odd), dilation
    int32_t n_sub, Lsub;        // phase 1: sub-segments per tile, rows per sub-segment (12 n + 2)
    int32_t TT;                 // output rows per tile (multiple of 16)
    int32_t n_tiles;            // tiles per batch entry
    int32_t a_stride, w_stride; // shared-memory row strides in elements (C padded to 16, + 8)
    int32_t a_rows;             // rows of the activation tile = n_sub * Lsub
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NT8: output-channel tiles of 8 (C / 8);  KS: input-channel steps of 16 (ceil(C / 16))
template <bool RES, int NT8, int KS>
__global__ void __launch_bounds__(kAcThreads, 2) afa_cl_actconv_kernel(const __grid_constant__ ActConvArgs args) {
    using T = __nv_bfloat16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* wsm = reinterpret_cast<T*>(smem_raw);                                        // [k * C][w_stride]
    T* asm_ = wsm + (size_t)args.k * args.C * args.w_stride;                        // [a_rows][a_stride]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = args.C, Tlen = args.T;
    // CTA order: batch fastest, then tile -- with the last and the first tile of the rows first (their edge-mode
    // walks are the slow ones and must not be the tail of the grid)
    const int tslot = (int)(blockIdx.x / (uint32_t)args.batch);
    const int b = (int)(blockIdx.x - (uint32_t)tslot * (uint32_t)args.batch);
    const int tile = tslot == 0 ? args.n_tiles - 1 : tslot - 1;
    const int P = (args.k / 2) * args.dil;
    const int t_out0 = tile * args.TT;                 // first output row of this tile
    const int tile_t0 = t_out0 - P;                    // time index of activation-tile row 0

    // ---- stage the weights (bf16, [tap][c_out][c_in]) and clear the K padding of both tiles
    {
        const int rows = args.k * C;
        const int cvec = C / 8;                         // 16-byte chunks per weight row
        for (int i = tid; i < rows * cvec; i += kAcThreads) {
            const int r = i / cvec, v = i - r * cvec;
            *reinterpret_cast<uint4*>(wsm + (size_t)r * args.w_stride + v * 8) =
                __ldg(reinterpret_cast<const uint4*>(args.w + (size_t)r * C + v * 8));
        }
        const int cpad = KS * 16;
        if (cpad > C) {
            const int padv = (cpad - C) / 8;
            for (int i = tid; i < rows * padv; i += kAcThreads) {
                const int r = i / padv, v = i - r * padv;
                *reinterpret_cast<uint4*>(wsm + (size_t)r * args.w_stride + C + v * 8) = make_uint4(0, 0, 0, 0);
            }
            for (int i = tid; i < args.a_rows * padv; i += kAcThreads) {
                const int r = i / padv, v = i - r * padv;
                *reinterpret_cast<uint4*>(asm_ + (size_t)r * args.a_stride + C + v * 8) = make_uint4(0, 0, 0, 0);
            }
        }
    }

    // ---- phase 1: the activation walk of (sub-segment, channel) pairs into the tile
    {
        const int pair = tid;                           // channel fastest: a warp's lanes read consecutive channels
        const int s = pair / C, c = pair - s * C;
        const bool active = s < args.n_sub;
        const int t0 = tile_t0 + s * args.Lsub;
        const bool fast = !active || (t0 >= 5 && t0 + args.Lsub + 5 < Tlen);
        const bool all_fast = __syncthreads_and(fast ? 1 : 0) != 0;
        if (active) {
            const T* px = static_cast<const T*>(args.x) + (int64_t)b * args.x_bs + c;
            const T* pr = RES ? static_cast<const T*>(args.res) + (int64_t)b * args.res_bs + c : nullptr;
            T* ps = RES ? static_cast<T*>(args.xsum) + (int64_t)b * args.xsum_bs + c : nullptr;
            const ChanParams cp = load_chan_params(args.alpha, args.beta, c, args.flags);
            const float bias = args.bias ? __ldg(args.bias + c) : 0.f;
            TileSink ts;
            ts.col = asm_ + c;
            ts.stride = args.a_stride;
            ts.tile_t0 = tile_t0;
            ts.own_lo = t_out0;
            ts.own_hi = min(t_out0 + args.TT, Tlen);
            const uint32_t amask = __activemask();
            if (all_fast) walk_cl<T, 0, RES, 2>(px, pr, ps, nullptr, C, t0, args.Lsub, Tlen, cp.a_eff, cp.ib, bias, args.taps, nullptr, amask, &ts);
            else walk_cl<T, 1, RES, 2>(px, pr, ps, nullptr, C, t0, args.Lsub, Tlen, cp.a_eff, cp.ib, bias, args.taps, nullptr, amask, &ts);
        }
    }
    __syncthreads();

    // ---- phase 2: y[t][co] = sum_j sum_ci W[j][co][ci] * a[t + j*dil - P][ci]; tile row of a[t + j*dil - P] = (t - t_out0) + j*dil
    {
        const uint32_t a_base = smem_u32(asm_);
        const uint32_t w_base = smem_u32(wsm);
        const int a_stride_b = args.a_stride * 2, w_stride_b = args.w_stride * 2;
        const int n_mt = args.TT / 16;
        T* yb = static_cast<T*>(args.y) + (int64_t)b * args.y_bs;
        // ldmatrix lane roles.  A (16 x 16): lane -> row lane % 16, column half lane / 16.
        const int a_row = lane & 15, a_colb = (lane >> 4) * 16;
        // B x4 = two output-channel tiles: lane -> c_out (lane % 8) + 8 * (lane / 16), k half (lane / 8) % 2;  x2 = one tile
        const int b_row4 = (lane & 7) + ((lane >> 4) << 3), b_colb = ((lane >> 3) & 1) * 16;
        const int b_row2 = lane & 7;
        // MT time tiles per sweep: their accumulation chains are independent (mma.sync latency is hidden by
        // MT * NT8 chains instead of NT8) and they share every weight fragment
        constexpr int MT = (NT8 <= 2) ? 4 : (NT8 <= 4 ? 3 : 2);
        for (int mt0 = warp * MT; mt0 < n_mt; mt0 += (kAcThreads / 32) * MT) {
            if (t_out0 + mt0 * 16 >= Tlen) break;
            float acc[MT][NT8][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < NT8; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
            for (int j = 0; j < args.k; ++j) {
                const uint32_t w_addr = w_base + (uint32_t)(j * C * w_stride_b);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t af[MT][4];
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        // tiles past the end of this CTA's range re-read the last valid one (results discarded)
                        const int mt = min(mt0 + m, n_mt - 1);
                        ldmatrix_x4(a_base + (uint32_t)((mt * 16 + j * args.dil + a_row) * a_stride_b + a_colb + ks * 32),
                                    af[m][0], af[m][1], af[m][2], af[m][3]);
                    }
#pragma unroll
                    for (int n = 0; n + 1 < NT8; n += 2) {
                        uint32_t b0, b1, b2, b3;
                        ldmatrix_x4(w_addr + (uint32_t)((n * 8 + b_row4) * w_stride_b + b_colb + ks * 32), b0, b1, b2, b3);
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            mma_bf16_16816(acc[m][n], af[m][0], af[m][1], af[m][2], af[m][3], b0, b1);
                            mma_bf16_16816(acc[m][n + 1], af[m][0], af[m][1], af[m][2], af[m][3], b2, b3);
                        }
                    }
                    if (NT8 & 1) {
                        uint32_t b0, b1;
                        ldmatrix_x2(w_addr + (uint32_t)(((NT8 - 1) * 8 + b_row2) * w_stride_b + b_colb + ks * 32), b0, b1);
#pragma unroll
                        for (int m = 0; m < MT; ++m) mma_bf16_16816(acc[m][NT8 - 1], af[m][0], af[m][1], af[m][2], af[m][3], b0, b1);
                    }
                }
            }
            // C fragment: rows lane / 4 and lane / 4 + 8, channels 8 n + 2 (lane % 4) + {0, 1}
            const int cc = 2 * (lane & 3);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                if (mt0 + m >= n_mt) break;
                const int r0 = t_out0 + (mt0 + m) * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
                for (int n = 0; n < NT8; ++n) {
                    if (r0 < Tlen) *reinterpret_cast<__nv_bfloat162*>(yb + (int64_t)r0 * C + n * 8 + cc) = __floats2bfloat162_rn(acc[m][n][0], acc[m][n][1]);
                    if (r1 < Tlen) *reinterpret_cast<__nv_bfloat162*>(yb + (int64_t)r1 * C + n * 8 + cc) = __floats2bfloat162_rn(acc[m][n][2], acc[m][n][3]);
                }
            }
        }
    }
}

}  // namespace afa
