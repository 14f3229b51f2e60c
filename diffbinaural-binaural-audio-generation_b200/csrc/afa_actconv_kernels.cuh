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
    const void* addend;         // optional [B, T, C]: y = conv(...) + addend (the residual add `x = xt + x`, bigvgan.py:141)
    FwdTaps taps;
    int64_t x_bs, res_bs, xsum_bs, y_bs, addend_bs;
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
                const T* ab = args.addend ? static_cast<const T*>(args.addend) + (int64_t)b * args.addend_bs : nullptr;
#pragma unroll
                for (int n = 0; n < NT8; ++n) {
                    if (ab) {
                        if (r0 < Tlen) {
                            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ab + (int64_t)r0 * C + n * 8 + cc));
                            acc[m][n][0] += f.x; acc[m][n][1] += f.y;
                        }
                        if (r1 < Tlen) {
                            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ab + (int64_t)r1 * C + n * 8 + cc));
                            acc[m][n][2] += f.x; acc[m][n][3] += f.y;
                        }
                    }
                    if (r0 < Tlen) *reinterpret_cast<__nv_bfloat162*>(yb + (int64_t)r0 * C + n * 8 + cc) = __floats2bfloat162_rn(acc[m][n][0], acc[m][n][1]);
                    if (r1 < Tlen) *reinterpret_cast<__nv_bfloat162*>(yb + (int64_t)r1 * C + n * 8 + cc) = __floats2bfloat162_rn(acc[m][n][2], acc[m][n][3]);
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Same kernel with phase 2 on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// ncu on the mma.sync version (C = 24, k = 7): tensor pipe 28 % + FMA pipe 40 % busy, back to back -- legacy
// HMMA.16816 retires ~1000 MAC/clk/SM on this part, an eighth of the tcgen05 rate, so the "side job" was a third
// of the kernel.  Here one thread issues, per block of 128 output rows, k * KS instructions
// tcgen05.mma.cta_group::1.kind::f16  (M = 128 time steps, N = c_out padded to 16, K = 16 input channels):
//   * operands straight from shared memory through matrix descriptors, K-major, no swizzle: the canonical layout
//     ((8, m), 2) : ((16 B, SBO), LBO) with SBO = 128 B makes a row's address LINEAR in the row index
//     (row * 16 B + k_chunk * LBO), so phase 1 writes the tile as column strips [c / 8][row][8] and the k taps of the
//     (dilated) convolution are nothing but descriptor start addresses shifted by j * dilation rows;
//   * all row blocks of the tile are issued up front into disjoint TMEM column ranges (n_mb * NPAD <= 256 columns),
//     one tcgen05.commit -> mbarrier per block; the 8 warps drain finished blocks (tcgen05.ld 32x32b: a thread
//     receives one time step's c_out accumulators = one contiguous row of y) while later blocks still compute.
// ------------------------------------------------------------------------------------------------
struct ActConvTcArgs {
    ActConvArgs a;
    int32_t rows_alloc;   // rows per column strip of the activation tile (odd: strips fall into different banks)
    int32_t n_mb;         // blocks of 128 output rows per tile
    int32_t tmem_cols;    // power of two >= n_mb * NPAD
    int32_t xin_rows;     // XS variant: rows of x staged per tile (a_rows + 11), else 0
};

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes) {
    // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version = 1 [46,48), no swizzle
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
// bounded wait: a wrong descriptor must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spins > (1u << 26)) __trap();
    }
}

constexpr int kAcMaxBlocks = 8;
#ifndef AFA_TC_DEBUG
#define AFA_TC_DEBUG 0
#endif

// NPAD: c_out padded to a multiple of 16 (the MMA's N);  KS: input-channel steps of 16;  XS: the tile's input rows (one
// contiguous range of the channels-last array) are staged in shared memory by ONE bulk TMA copy issued at kernel entry,
// so phase 1 reads x with short-latency shared-memory loads instead of waiting on global loads (ncu: long_scoreboard was
// the top stall of phase 1 with only 16 warps per SM to hide it)
template <bool RES, int NPAD, int KS, bool XS = false>
__global__ void __launch_bounds__(kAcThreads, 2) afa_cl_actconv_tc_kernel(const __grid_constant__ ActConvTcArgs targs) {
    using T = __nv_bfloat16;
    const ActConvArgs& args = targs.a;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NSTRIP = KS * 2;                                   // 16-byte column strips (8 channels each)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);          // [kAcMaxBlocks]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 64);
    T* wsm = reinterpret_cast<T*>(smem_raw + 128);                   // [k][NSTRIP][NPAD][8]
    T* asm_ = wsm + (size_t)args.k * NSTRIP * NPAD * 8;              // [NSTRIP][rows_alloc][8]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = args.C, Tlen = args.T;
    const int tslot = (int)(blockIdx.x / (uint32_t)args.batch);
    const int b = (int)(blockIdx.x - (uint32_t)tslot * (uint32_t)args.batch);
    const int tile = tslot == 0 ? args.n_tiles - 1 : tslot - 1;
    const int P = (args.k / 2) * args.dil;
    const int t_out0 = tile * args.TT;
    const int tile_t0 = t_out0 - P;
    const int rows_alloc = targs.rows_alloc;
    T* xin = asm_ + (size_t)NSTRIP * rows_alloc * 8;                 // XS: [xin_rows][C]
    uint64_t* xbar = reinterpret_cast<uint64_t*>(smem_raw + 72);
    const int xrow0 = max(0, tile_t0 - 5);                           // first staged row of x
    if (XS && tid == 0) {
        const int xrow1 = min(Tlen, tile_t0 + args.a_rows + 6);
        const uint32_t bytes = (uint32_t)(xrow1 - xrow0) * (uint32_t)C * 2u;
        mbar_init(xbar, 1);
        fence_mbar_init();
        mbar_expect_tx(xbar, bytes);
        tma_load_1d(xin, static_cast<const T*>(args.x) + (int64_t)b * args.x_bs + (int64_t)xrow0 * C, bytes, xbar);
    }

    // ---- TMEM allocation (warp 0), mbarriers, weights in the B layout, zero padding
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)targs.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < targs.n_mb; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    {
        const int cstrips = C / 8;
        const int total = args.k * NSTRIP * NPAD;
        for (int i = tid; i < total; i += kAcThreads) {
            const int n = i % NPAD, kc = (i / NPAD) % NSTRIP, j = i / (NPAD * NSTRIP);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (n < C && kc < cstrips) v = __ldg(reinterpret_cast<const uint4*>(args.w + ((size_t)j * C + n) * C + kc * 8));
            *reinterpret_cast<uint4*>(wsm + (size_t)i * 8) = v;
        }
        for (int kc = cstrips; kc < NSTRIP; ++kc)
            for (int r = tid; r < rows_alloc; r += kAcThreads)
                *reinterpret_cast<uint4*>(asm_ + ((size_t)kc * rows_alloc + r) * 8) = make_uint4(0, 0, 0, 0);
    }

    // ---- phase 1: the activation walk of (sub-segment, channel) pairs into the column strips
    {
        const int s = tid / C, c = tid - s * C;
        const bool active = s < args.n_sub;
        const int t0 = tile_t0 + s * args.Lsub;
        const bool fast = !active || (t0 >= 5 && t0 + args.Lsub + 5 < Tlen);
        const bool all_fast = __syncthreads_and(fast ? 1 : 0) != 0;
        if (XS) mbar_wait_bounded(xbar, 0);                      // the staged input rows have landed
        if (active) {
            // XS: element t of this channel at px[t * C] inside the staged rows (the pointer may lie before the buffer)
            const T* px = XS ? xin + c - (int64_t)xrow0 * C : static_cast<const T*>(args.x) + (int64_t)b * args.x_bs + c;
            const T* pr = RES ? static_cast<const T*>(args.res) + (int64_t)b * args.res_bs + c : nullptr;
            T* ps = RES ? static_cast<T*>(args.xsum) + (int64_t)b * args.xsum_bs + c : nullptr;
            const ChanParams cp = load_chan_params(args.alpha, args.beta, c, args.flags);
            const float bias = args.bias ? __ldg(args.bias + c) : 0.f;
            TileSink ts;
            ts.col = asm_ + ((size_t)(c >> 3) * rows_alloc) * 8 + (c & 7);
            ts.stride = 8;
            ts.tile_t0 = tile_t0;
            ts.own_lo = t_out0;
            ts.own_hi = min(t_out0 + args.TT, Tlen);
            const uint32_t amask = __activemask();
            if (all_fast) walk_cl<T, 0, RES, 2, XS>(px, pr, ps, nullptr, C, t0, args.Lsub, Tlen, cp.a_eff, cp.ib, bias, args.taps, nullptr, amask, &ts);
            else walk_cl<T, 1, RES, 2, XS>(px, pr, ps, nullptr, C, t0, args.Lsub, Tlen, cp.a_eff, cp.ib, bias, args.taps, nullptr, amask, &ts);
        }
    }
    // generic-proxy writes of the tiles -> visible to the tensor core's async proxy; TMEM address -> everyone
    fence_proxy_async_smem();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // ---- phase 2, issue: all row blocks up front; block mb is issued by lane 0 of warp mb (one thread per MMA stream,
    //      eight streams side by side -- a single issuer spent ~100 cycles per MMA building descriptors and was the
    //      longest thing in the kernel at k = 11).  Descriptors differ only in their start-address field: add rows.
    if (lane == 0 && warp < targs.n_mb && !(AFA_TC_DEBUG & 1)) {
        // cute::UMMA::InstrDescriptor: D = f32 [4,6) = 1, A = bf16 [7,10) = 1, B = bf16 [10,13) = 1, both K-major, N >> 3 [17,23), M >> 4 [24,29)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t lbo_a = (uint32_t)rows_alloc * 16u, lbo_b = (uint32_t)NPAD * 16u;
        const int mb = warp;
        const uint64_t a0 = umma_desc_kmajor(smem_u32(asm_) + (uint32_t)(mb * 128) * 16u, lbo_a);
        const uint64_t b0 = umma_desc_kmajor(smem_u32(wsm), lbo_b);
        const uint32_t d_tmem = tmem_base + (uint32_t)(mb * NPAD);
        uint32_t a_row = 0, b_row = 0;                               // start-address offsets in 16-byte units
        for (int j = 0; j < args.k; ++j) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
                umma_bf16_ss(d_tmem, a0 + (uint64_t)(a_row + (uint32_t)(ks * 2) * (uint32_t)rows_alloc),
                             b0 + (uint64_t)(b_row + (uint32_t)(ks * 2 * NPAD)), idesc, (j | ks) != 0 ? 1u : 0u);
            a_row += (uint32_t)args.dil;
            b_row += (uint32_t)(NSTRIP * NPAD);
        }
        umma_commit(&bars[mb]);
    }
    __syncwarp();

    // ---- phase 2, drain: warp w reads TMEM lanes 32 (w % 4) ..; the two warp groups take alternate row blocks
    {
        const int q = warp & 3, grp = warp >> 2;
        T* yb = static_cast<T*>(args.y) + (int64_t)b * args.y_bs;
        for (int mb = grp; mb < targs.n_mb; mb += 2) {
            if (!(AFA_TC_DEBUG & 1)) mbar_wait_bounded(&bars[mb], 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int r = mb * 128 + q * 32 + lane;                  // output row within the tile = TMEM lane
            const int t = t_out0 + r;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * NPAD);
            uint32_t v[NPAD];
#pragma unroll
            for (int c8 = 0; c8 < NPAD / 8; ++c8) if (!(AFA_TC_DEBUG & 2)) tmem_ld8(taddr + c8 * 8, *reinterpret_cast<uint32_t(*)[8]>(&v[c8 * 8]));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (r < args.TT && t < Tlen && !(AFA_TC_DEBUG & 4)) {
                T* dst = yb + (int64_t)t * C;
                const T* add = args.addend ? static_cast<const T*>(args.addend) + (int64_t)b * args.addend_bs + (int64_t)t * C : nullptr;
#pragma unroll
                for (int c8 = 0; c8 < NPAD / 8; ++c8) {
                    if (c8 * 8 < C) {
                        if (add) {                                   // the residual add of the block, in fp32 before the one rounding
                            float f[8];
                            IO<T>::load_chunk(add + c8 * 8, f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[c8 * 8 + e] = __float_as_uint(__uint_as_float(v[c8 * 8 + e]) + f[e]);
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[c8 * 8 + 2 * e]), __uint_as_float(v[c8 * 8 + 2 * e + 1]));
                            pk[e] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        *reinterpret_cast<uint4*>(dst + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)targs.tmem_cols) : "memory");
    }
}

}  // namespace afa
