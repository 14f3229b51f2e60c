// afa_kernels.cuh -- fused Activation1d (2x Kaiser-sinc upsample -> Snake/SnakeBeta -> 2x low-pass
// downsample) forward and backward kernels for sm_100a.
//
// What it replaces (reference, /root/reference/BigVGAN): alias_free_activation/act.py:25-30, i.e.
// resample.py:29-38 + activations.py:51-62/113-126 + filter.py:94-101, and their autograd backward.
//
// Design (DESIGN.md has the long form):
//   * Rows (batch*channel) are independent.  Every row is cut into SEGMENTS of L = CH*VEC samples
//     (VEC = elements per 16 bytes, CH odd).  One thread owns one segment and walks it in time order,
//     keeping the 6-tap polyphase windows and the 7 live output accumulators in register RINGS, so the
//     2x-rate intermediate never leaves the register file.
//   * Segments are numbered row-major; a WARP owns 32 consecutive segments (a "warp tile"), which is
//     one CONTIGUOUS range of the flat [rows*T] array.  That range (+8 halo elements either side) is
//     staged in shared memory by ONE 1-D bulk TMA copy (cp.async.bulk + mbarrier, SASS UBLKCP), the
//     result is written IN PLACE over the staged input and leaves with one bulk TMA store.
//   * Warps are autonomous and persistent: each has two stages and two mbarriers of its own, loops
//     over warp tiles round-robin, and prefetches tile i+1 (TMA) while it walks tile i.  There is no
//     CTA-wide barrier after start-up and no global load/store instruction in the steady state.
//   * Lane i touches shared memory at a stride of CH 16-byte chunks; CH odd makes every quarter-warp
//     LDS.128/STS.128 hit 8 distinct 16-byte bank groups -> conflict-free without swizzle.
//   * The walk is a ROLLED loop whose body is S steps (S = ring size) with static register indices:
//     ~7 KB of SASS per mode instead of the 57 KB of a fully unrolled walk (v1 was I-cache bound).
//   * Replicate padding lives in two places: the x clamp and the clamp of the
//     ACTIVATED 2x signal.  Warps whose segments all sit >= 5 samples from both row ends take the
//     branch-free MODE 0 walk; warps touching a row end take MODE 1 (same schedule + selects); rows
//     whose length is not a multiple of VEC (no 16-byte alignment, TMA illegal) run the MODE 2 kernel
//     (scalar staging), still inside this library.
//   * No tensor cores: a depthwise 12-tap stencil is not a contraction (north_star).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace afa {

constexpr int kHalo = 8;  // elements staged either side of a warp tile's flat range (>= 5 needed)
constexpr int kBarBytes = 64;  // mbarrier block at the start of dynamic shared memory (2 per warp)

// Resident CTAs per SM the aligned kernels are compiled for (register cap = 65536 / (128 * n)).
#ifndef AFA_FWD_MINB_F32
#define AFA_FWD_MINB_F32 5
#endif
#ifndef AFA_FWD_MINB_BF16
#define AFA_FWD_MINB_BF16 4
#endif
#ifndef AFA_BWD_MINB
#define AFA_BWD_MINB 4
#endif
// Shared-memory stages per warp in the forward kernel: 2 = prefetch tile i+1 while walking tile i;
// 1 = no intra-warp prefetch (half the shared memory per warp -> more resident warps hide the reload).
#ifndef AFA_FWD_STAGES
#define AFA_FWD_STAGES 2
#endif
#ifndef AFA_BWD_STAGES
#define AFA_BWD_STAGES 2
#endif

// Filter taps as kernel parameters, pre-paired for packed f32x2 math (they become uniform-register
// operands of FFMA2).  The 2x-rate signal is handled as pairs (s[2m-1], s[2m]): both members read the
// same six inputs x[m-3..m+2] and feed the same six outputs y[m-3..m+2], so every tap is ONE FFMA2 whose
// other operand is a scalar broadcast -- no register-pair alignment constraints anywhere.
//   cu[j] = (2*f_up[2j], 2*f_up[2j+1]) : (u[2m-1], u[2m])  = sum_j cu[j] * x[m+2-j]
//   cd[j] = (  f_dn[2j],   f_dn[2j+1]) : (yo, ye)[m+2-j]  += cd[j] * (s[2m-1], s[2m]),  y = yo + ye
// The backward reuses both: (ds[2m-1], ds[2m]) = sum_j cd[j] * gy[m+2-j] and
// (dxo, dxe)[m+2-j] += cu[j] * (du[2m-1], du[2m]).
struct PairTaps {
    float2 cu[6];
    float2 cd[6];
};
struct FwdTaps {
    PairTaps p;
    float ue[6];          // 2*f_up[2j+1]  (scalar copies for the edge constants s[0], s[2T-1])
    float uo[6];          // 2*f_up[2j]
};
struct BwdTaps {
    PairTaps p;
    float lo[3];          // folded taps of the left  replicate pad of s: gy[0], gy[1], gy[2]
    float hi[3];          // folded taps of the right replicate pad of s: gy[T-1], gy[T-2], gy[T-3]
};

// n / d for 0 <= n < 2^31 by multiply-high (host computes mul/shr; d == 1 -> mul = 0)
struct FastDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return mul ? (__umulhi(n, mul) >> shr) : n; }
};

struct Geometry {
    int64_t total;        // rows * T
    uint32_t total_segs;  // rows * nseg
    uint32_t n_wtiles;    // ceil(total_segs / 32)
    FastDiv nseg;         // segments per row
    FastDiv chan;         // channels
    int32_t T;
    int32_t flags;
    int32_t half;         // rows are only 8-byte aligned (bf16, T % 8 == 4): 8-byte shared-memory chunk I/O, split stores
};

struct FwdArgs {
    const void* x;
    void* y;
    const float* alpha;
    const float* beta;
    FwdTaps taps;
    Geometry g;
};
struct BwdArgs {
    const void* x;
    const void* gy;
    void* gx;
    const float* alpha;
    const float* beta;
    float* part;  // [2][total_segs] per-segment parameter-gradient partials
    uint32_t* cnt;  // per-channel arrival counters of the split second stage (afa_param_grad_finalize), or nullptr: zeroed by CTA 0
    int32_t n_cnt;
    BwdTaps taps;
    Geometry g;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared, completes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, bulk-group completion
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk stores have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// ------------------------------------------------------------------------------------------------
// element I/O on shared memory
// ------------------------------------------------------------------------------------------------
template <typename T>
struct IO;
template <>
struct IO<float> {
    static constexpr int VEC = 4;
    static __device__ __forceinline__ void load_chunk(const float* p, float* o) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    static __device__ __forceinline__ void store_chunk(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void load_chunk(const float* p, float* o, bool) { load_chunk(p, o); }   // fp32 rows: 16B or nothing
    static __device__ __forceinline__ void store_chunk(float* p, const float* v, bool) { store_chunk(p, v); }
    static __device__ __forceinline__ void store_half_chunk(float*, const float*) {}
    static __device__ __forceinline__ float load1(const float* p) { return *p; }
    static __device__ __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <>
struct IO<__nv_bfloat16> {
    static constexpr int VEC = 8;
    static __device__ __forceinline__ void load_chunk(const __nv_bfloat16* p, float* o) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = __uint_as_float(w[i] << 16);
            o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store_chunk(__nv_bfloat16* p, const float* v) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // `half`: the chunk is only 8-byte aligned -> two 8-byte accesses
    static __device__ __forceinline__ void load_chunk(const __nv_bfloat16* p, float* o, bool half) {
        if (!half) {
            load_chunk(p, o);
        } else {
            const uint2 a = *reinterpret_cast<const uint2*>(p);
            const uint2 b = *reinterpret_cast<const uint2*>(p + 4);
            const uint32_t w[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                o[2 * i] = __uint_as_float(w[i] << 16);
                o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            }
        }
    }
    static __device__ __forceinline__ void store_chunk(__nv_bfloat16* p, const float* v, bool half) {
        if (!half) {
            store_chunk(p, v);
        } else {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
            *reinterpret_cast<uint2*>(p + 4) = make_uint2(w[2], w[3]);
        }
    }
    // rows that are 8-byte aligned end in the middle of a chunk: store only its first 4 samples
    static __device__ __forceinline__ void store_half_chunk(__nv_bfloat16* p, const float* v) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    }
    static __device__ __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// Snake / SnakeBeta: u + sin^2(alpha*u) / (beta + 1e-9)            activations.py:60, :124
__device__ __forceinline__ float snake_f(float u, float a, float ib) {
    const float sn = __sinf(u * a);
    return fmaf(ib, sn * sn, u);
}

// effective parameters of one channel from the raw ones           activations.py:119-123
__device__ __forceinline__ void effective_params(float a_raw, float b_raw, int flags, float& a_eff, float& b_eff,
                                                 float& ib) {
    float a = a_raw;
    float b = (flags & 2) ? a_raw : b_raw;
    if (flags & 1) {
        a = expf(a);
        b = expf(b);
    }
    a_eff = a;
    b_eff = b;
    ib = 1.0f / (b + 0.000000001f);
}

// What a walk needs to know to prefetch the warp's NEXT tile once its previous bulk store has
// drained (issued by lane 0 after the first S loop steps of the walk), and to prepare that tile's
// channel parameters (all lanes, same place: the global-load and exp/rcp latencies hide under the walk).
struct Prefetch {
    const void* src0;
    const void* src1;    // second tensor (backward: gy), or nullptr
    uint32_t dst0, dst1; // shared-memory addresses of the stage being refilled
    uint32_t bar;        // shared-memory address of that stage's mbarrier
    uint32_t bytes;      // per tensor; 0 = nothing to prefetch
};
struct ChanParams {
    float a_eff, b_eff, ib;
};
struct NextChan {
    const float* alpha;
    const float* beta;
    int32_t c;           // channel of this lane's segment in the next tile, or -1
    int32_t flags;
};
__device__ __forceinline__ void issue_prefetch(const Prefetch& pf) {
    if (pf.bytes) {
        tma_store_wait_read();   // the stage being refilled was the source of the previous bulk store
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pf.bar),
                     "r"(pf.src1 ? 2 * pf.bytes : pf.bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         pf.dst0),
                     "l"(pf.src0), "r"(pf.bytes), "r"(pf.bar)
                     : "memory");
        if (pf.src1)
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pf.dst1),
                "l"(pf.src1), "r"(pf.bytes), "r"(pf.bar)
                : "memory");
    }
}
__device__ __forceinline__ ChanParams load_chan_params(const float* alpha, const float* beta, int c, int flags) {
    ChanParams p;
    const float a_raw = __ldg(alpha + c);
    const float b_raw = (flags & 2) ? a_raw : __ldg(beta + c);
    effective_params(a_raw, b_raw, flags, p.a_eff, p.b_eff, p.ib);
    return p;
}
// the mid-walk hook: lane 0 refills the other stage, every lane prepares its next channel parameters
__device__ __forceinline__ void mid_walk_hook(const Prefetch& pf, const NextChan& nc, ChanParams& next, int lane) {
    if (lane == 0) issue_prefetch(pf);
    if (nc.c >= 0) next = load_chan_params(nc.alpha, nc.beta, nc.c, nc.flags);
}

template <int VEC>
struct RingCfg {
    static constexpr int S = (VEC == 4) ? 12 : 16;  // ring size = steps per loop body (>= 6 + VEC, multiple of VEC)
};

__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }

// ------------------------------------------------------------------------------------------------
// forward walk of one segment, in place.  MODE 0: interior (branch-free); 1: touches a row end
// (16B-aligned rows); 2: rows not 16B-aligned (scalar smem I/O).  `row0` points at this row's sample 0
// inside the staged tile (may lie outside the tile; only in-range positions are touched).
// Step q (1 <= q <= L+5) handles the 2x-rate pair (s[2m-1], s[2m]), m = t0 - 3 + q, and completes
// y[t0 + q - 6].  Register rings with compile-time slots hold the pending upsampler pairs and the two partial
// sums of each pending y (odd-phase taps, even-phase taps: two fixed 6-term FMA chains, then one add), so
// results do not depend on how rows are cut into segments.
// ------------------------------------------------------------------------------------------------
template <typename T, int CH, int MODE>
__device__ __forceinline__ void walk_fwd(T* __restrict__ row0, int t0, int Tlen, float a, float ib, const FwdTaps& tp,
                                         const Prefetch& pf, const NextChan& nc, ChanParams& next, int lane,
                                         uint32_t mask, bool half) {
    using io = IO<T>;
    constexpr int VEC = io::VEC;
    constexpr int L = CH * VEC;
    constexpr int S = RingCfg<VEC>::S;
    constexpr bool kStaticTail = (VEC == 4 && MODE == 0);   // last five steps straight-line (fp32 interior walk only, see below)
    T* seg = row0 + t0;
    // Both FIRs run in TRANSPOSED (scatter) form: a new input updates the six pending outputs it feeds, so
    // the 12 FFMA2 of a step are independent of each other (dependences only reach back >= 1 step).
    //   up[(m) % S]   : pending (u[2m-1], u[2m]);   x[m+2] is its last contribution
    //   ac[(o) % S]   : pending (yo, ye) of y[t0+o]
    float2 up[S], ac[S];
    float xb[VEC];            // the current 16-byte chunk of x
    float hold[8], yb[VEC];
    float s_first = 0.f, s_last = 0.f, x_last = 0.f;

    // ---- preload x[t0-8 .. t0-1] and the edge constants; nothing below reads a neighbour's first 8
    //      samples again before the final __syncwarp, and nobody overwrites them before it (hold[]).
    float xpre[8];
    if (MODE != 2) {
#pragma unroll
        for (int c = 0; c < 8 / VEC; ++c) io::load_chunk(seg - 8 + c * VEC, &xpre[c * VEC], half);
    } else {
#pragma unroll
        for (int j = 3; j < 8; ++j) xpre[j] = io::load1(row0 + min(max(t0 - 8 + j, 0), Tlen - 1));
    }
    if (MODE != 0) {
        if (t0 == 0) {  // left replicate pad of x, and s[0] which the left pad of s repeats      filter.py:98
            const float x0 = io::load1(row0);
            const float x1 = io::load1(row0 + min(1, Tlen - 1));
            const float x2 = io::load1(row0 + min(2, Tlen - 1));
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) xpre[j] = x0;
            }
            float u = tp.ue[0] * x2;
            u = fmaf(tp.ue[1], x1, u);
#pragma unroll
            for (int j = 2; j < 6; ++j) u = fmaf(tp.ue[j], x0, u);
            s_first = snake_f(u, a, ib);
        }
        if (Tlen - 1 <= t0 + L + 5) {  // the row ends inside this walk's reach
            x_last = io::load1(row0 + Tlen - 1);
            float u = 0.f;             // s[2T-1]: the value the right pad of s repeats
#pragma unroll
            for (int j = 0; j < 6; ++j) u = fmaf(tp.uo[j], io::load1(row0 + min(max(Tlen + 2 - j, 0), Tlen - 1)), u);
            s_last = snake_f(u, a, ib);
        }
    }
    __syncwarp(mask);

    // one step; Q is the static part of the step index (ring slots), q the dynamic step number.
    // Steps q <= 0 only feed x[t0-5 .. t0-1] into the pending upsampler outputs (warm-up).
    // `epi`: one of the last five steps (static): work that only feeds outputs of the NEXT segment is skipped
    auto step = [&](const int Q, const int q, const bool first_iter, const bool epi) {
        // --- the new input x[m+2] = x[t0 + q - 1]
        float xv;
        if (first_iter && Q <= 0) {
            xv = xpre[Q + 7];
        } else if (MODE != 2) {
            if ((Q + 7) % VEC == 0) {
                io::load_chunk(seg + q - 1, xb, half);
                if (MODE == 1) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) xb[e] = (t0 + q - 1 + e > Tlen - 1) ? x_last : xb[e];
                }
            }
            xv = xb[(Q + 7) % VEC];
        } else {
            xv = io::load1(row0 + min(t0 + q - 1, Tlen - 1));
        }
        // --- upsampler, transposed form: (u[2(m+j)-1], u[2(m+j)]) += cu[j] * x[m+2]     resample.py:32-36
        const float2 xx = bcast2(xv);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (!(first_iter && Q + j < 1) && !(epi && Q + j > L + 5)) {
                float2& pend = up[(Q + j + 4 * S) % S];
                pend = (j == 5) ? __fmul2_rn(tp.p.cu[5], xx) : __ffma2_rn(tp.p.cu[j], xx, pend);
            }
        }
        if (first_iter && Q < 1) return;
        // --- (u[2m-1], u[2m]) is complete: Snake / SnakeBeta on the pair             activations.py:124
        const float2 u2 = up[(Q + 4 * S) % S];
        const float2 th = __fmul2_rn(u2, bcast2(a));
        const float2 sn = make_float2(__sinf(th.x), __sinf(th.y));
        float2 s2 = __ffma2_rn(bcast2(ib), __fmul2_rn(sn, sn), u2);
        if (MODE != 0) {
            const int m = t0 - 3 + q;
            if (Q <= 3 && first_iter) {   // s[n] for n < 0 repeats s[0]
                if (m <= 0) s2.x = s_first;
                if (m < 0) s2.y = s_first;
            }
            if (m > Tlen) s2.x = s_last;   // s[n] for n >= 2T repeats s[2T-1]
            if (m >= Tlen) s2.y = s_last;
        }
        // --- low-pass, transposed form: (yo, ye)[m+2-j] += cd[j] * (s[2m-1], s[2m])     filter.py:98-99
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (!(first_iter && Q - 1 - j < 0) && !(epi && Q - 1 - j >= L)) {
                float2& acc = ac[(Q - 1 - j + 4 * S) % S];
                acc = (j == 0) ? __fmul2_rn(tp.p.cd[0], s2) : __ffma2_rn(tp.p.cd[j], s2, acc);
            }
        }
        // --- y[t0 + q - 6] is complete
        if (!(first_iter && Q < 6)) {
            const int o = q - 6;
            const float2 done = ac[(Q - 6 + 4 * S) % S];
            const float yv = done.x + done.y;
            if (MODE != 2) {
                yb[(Q - 6 + 4 * S) % VEC] = yv;
                if ((Q - 6 + 4 * S) % VEC == VEC - 1) {
                    const int c0 = o - (VEC - 1);                     // first sample of the finished chunk
                    if (first_iter && Q - 6 - (VEC - 1) < 8) {         // (static) a neighbour may still need these x
#pragma unroll
                        for (int e = 0; e < VEC; ++e) hold[Q - 6 - (VEC - 1) + e] = yb[e];
                    } else if (MODE == 0 || t0 + c0 + VEC <= Tlen) {
                        io::store_chunk(seg + c0, yb, half);
                    } else if (t0 + c0 < Tlen) {
                        io::store_half_chunk(seg + c0, yb);      // half-aligned rows end mid-chunk
                    }
                }
            } else {
                if (first_iter && Q - 6 < 8) hold[Q - 6] = yv;
                else if (t0 + o < Tlen) io::store1(seg + o, yv);
            }
        }
    };

    // q = -4 .. S+5 static (warm-up + the outputs that wait in hold[]), q = S+6 .. L rolled (S steps per trip, ring
    // slots static inside the body), q = L+1 .. L+5 static again (dead work towards the next segment removed)
    // (the static tail pays for the fp32 interior walk only; for bf16 (16-step ring) and for edge-mode warps the extra
    //  straight-line code costs more in instruction-cache misses than the removed dead work saves -- measured,
    //  profiles/r01_ab_epilogue.log)
#ifndef AFA_FWD_BODY_TRIPS
#define AFA_FWD_BODY_TRIPS 1
#endif
    constexpr int BODY = S * ((VEC == 4 && MODE == 0) ? AFA_FWD_BODY_TRIPS : 1);   // steps per rolled-loop trip
    constexpr int NR = kStaticTail ? L - S - 5 : L - S, NF = NR / BODY, RM = NR % BODY;
    static_assert(NR >= 0, "segment shorter than the ring");
#pragma unroll
    for (int q = -4; q < 6 + S; ++q) step(q, q, true, false);
    mid_walk_hook(pf, nc, next, lane);
#pragma unroll 1
    for (int it = 0; it < NF + (RM ? 1 : 0); ++it) {
        const int qb = 6 + S + it * BODY;
#pragma unroll
        for (int k = 0; k < BODY; ++k) {
            if (RM != 0 && k == RM) {
                if (it == NF) break;
            }
            step(6 + k, qb + k, false, false);
        }
    }
    if (kStaticTail) {
#pragma unroll
        for (int q = L + 1; q < L + 6; ++q) step(q, q, false, true);
    }
    __syncwarp(mask);  // every lane has finished reading its right halo: now the first 8 outputs may land
    if (MODE != 2) {
#pragma unroll
        for (int c = 0; c < 8 / VEC; ++c) {
            if (MODE == 0 || t0 + c * VEC + VEC <= Tlen) io::store_chunk(seg + c * VEC, &hold[c * VEC], half);
            else if (t0 + c * VEC < Tlen) io::store_half_chunk(seg + c * VEC, &hold[c * VEC]);
        }
    } else {
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (t0 + o < Tlen) io::store1(seg + o, hold[o]);
    }
}

// ------------------------------------------------------------------------------------------------
// backward walk of one segment: recomputes u from x, forms ds from gy, du = ds*(1 + a*ib*sin(2 a u)),
// scatters du into dx (written in place over x), and accumulates the segment's share of
// d/dalpha_eff, d/dbeta_eff.  Same packed-pair arithmetic and step numbering as the forward.
// ------------------------------------------------------------------------------------------------
template <typename T, int CH, int MODE>
__device__ __forceinline__ void walk_bwd(T* __restrict__ row0, const T* __restrict__ grow0, int t0, int Tlen, float a,
                                         float ib, const BwdTaps& tp, const Prefetch& pf, const NextChan& nc,
                                         ChanParams& next, int lane, uint32_t mask, bool half, float& ga_out,
                                         float& gb_out) {
    using io = IO<T>;
    constexpr int VEC = io::VEC;
    constexpr int L = CH * VEC;
    constexpr int S = RingCfg<VEC>::S;
    constexpr bool kStaticTail = (VEC == 4 && MODE == 0);
    T* seg = row0 + t0;
    const T* gseg = grow0 + t0;
    // transposed-form FIRs, as in the forward: pending (u[2m-1], u[2m]), pending (ds[2m-1], ds[2m]),
    // pending (dxo, dxe) of dx[t0+o]
    float2 up[S], dp[S], ac[S];
    float xb[VEC], gb_[VEC];
    float hold[8], yb[VEC];
    float d_lo = 0.f, d_hi = 0.f, x_last = 0.f;

    float xpre[8], gpre[8];
    if (MODE != 2) {
#pragma unroll
        for (int c = 0; c < 8 / VEC; ++c) {
            io::load_chunk(seg - 8 + c * VEC, &xpre[c * VEC], half);
            io::load_chunk(gseg - 8 + c * VEC, &gpre[c * VEC], half);
        }
    } else {
#pragma unroll
        for (int j = 3; j < 8; ++j) {
            const int t = t0 - 8 + j;
            xpre[j] = io::load1(row0 + min(max(t, 0), Tlen - 1));
            gpre[j] = (t >= 0) ? io::load1(grow0 + max(t, 0)) : 0.f;
        }
    }
    if (MODE != 0) {
#pragma unroll
        for (int i = 0; i < S; ++i) ac[i] = make_float2(0.f, 0.f);   // accumulators left of the row start only see FMAs
        if (t0 == 0) {
            if (MODE == 1) {
                const float x0 = io::load1(row0);
#pragma unroll
                for (int j = 0; j < 8; ++j) { xpre[j] = x0; gpre[j] = 0.f; }   // replicate x; gy does not extend
            }
            // adjoint of the left replicate pad of s: taps that fell on the pad fold onto s[0]
            const float g0 = io::load1(grow0);
            const float g1 = (Tlen > 1) ? io::load1(grow0 + min(1, Tlen - 1)) : 0.f;
            const float g2 = (Tlen > 2) ? io::load1(grow0 + min(2, Tlen - 1)) : 0.f;
            d_lo = fmaf(tp.lo[2], g2, fmaf(tp.lo[1], g1, tp.lo[0] * g0));
        }
        if (Tlen - 1 <= t0 + L + 5) {
            x_last = io::load1(row0 + Tlen - 1);
#pragma unroll
            for (int i = 0; i < 3; ++i) {   // adjoint of the right pad of s folds onto s[2T-1]
                const int t = Tlen - 1 - i;
                const float g = io::load1(grow0 + max(t, 0));
                d_hi = fmaf(tp.hi[i], (t >= 0) ? g : 0.f, d_hi);
            }
        }
    }
    __syncwarp(mask);

    const float a2 = 2.0f * a;
    const float iba = ib * a;
    float2 ga2 = make_float2(0.f, 0.f), gb2 = make_float2(0.f, 0.f);

    auto step = [&](const int Q, const int q, const bool first_iter, const bool epi) {
        // --- the new inputs x[m+2], gy[m+2]  (index t0 + q - 1)
        float xv, gv;
        if (first_iter && Q <= 0) {
            xv = xpre[Q + 7];
            gv = gpre[Q + 7];
        } else if (MODE != 2) {
            if ((Q + 7) % VEC == 0) {
                io::load_chunk(seg + q - 1, xb, half);
                io::load_chunk(gseg + q - 1, gb_, half);
                if (MODE == 1) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const bool past = t0 + q - 1 + e > Tlen - 1;
                        xb[e] = past ? x_last : xb[e];
                        gb_[e] = past ? 0.f : gb_[e];
                    }
                }
            }
            xv = xb[(Q + 7) % VEC];
            gv = gb_[(Q + 7) % VEC];
        } else {
            const int t = t0 + q - 1;
            xv = io::load1(row0 + min(t, Tlen - 1));
            const float g = io::load1(grow0 + min(t, Tlen - 1));
            gv = (t < Tlen) ? g : 0.f;
        }
        const float2 xx = bcast2(xv), gg = bcast2(gv);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (!(first_iter && Q + j < 1) && !(epi && Q + j > L + 5)) {
                float2& pu = up[(Q + j + 4 * S) % S];
                float2& pd = dp[(Q + j + 4 * S) % S];
                pu = (j == 5) ? __fmul2_rn(tp.p.cu[5], xx) : __ffma2_rn(tp.p.cu[j], xx, pu);
                pd = (j == 5) ? __fmul2_rn(tp.p.cd[5], gg) : __ffma2_rn(tp.p.cd[j], gg, pd);
            }
        }
        if (first_iter && Q < 1) return;
        const float2 u2 = up[(Q + 4 * S) % S];
        float2 d2 = dp[(Q + 4 * S) % S];
        const int m = t0 - 3 + q;
        if (MODE != 0) {
            // ds does not exist outside [0, 2T): what fell on the replicate pads was folded into d_lo / d_hi
            if (Q <= 3 && first_iter) {
                if (m <= 0) d2.x = 0.f;
                if (m < 0) d2.y = 0.f;
            }
            if (m > Tlen) d2.x = 0.f;
            if (m >= Tlen) d2.y = 0.f;
            if (Q == 3 && first_iter) {
                if (t0 == 0) d2.y += d_lo;      // ds[0]
            }
            if (m == Tlen) d2.x += d_hi;         // ds[2T-1]
        }
        const float2 ph = __fmul2_rn(bcast2(a2), u2);
        const float2 sn = make_float2(__sinf(ph.x), __sinf(ph.y));
        const float2 cs = make_float2(__cosf(ph.x), __cosf(ph.y));
        const float2 p2 = __fmul2_rn(d2, sn);
        const float2 du2 = __ffma2_rn(p2, bcast2(iba), d2);
        // parameter-gradient partials: s[2m-1] belongs to this segment for q in [4, L+3], s[2m] for q in [3, L+2]
        // (q <= S+5 and q > L are static steps, so the boundary cases cost nothing in the rolled loop)
        const bool own_x = first_iter ? (Q >= 4) : (epi ? (Q <= L + 3) : (kStaticTail || q <= L + 3));
        const bool own_y = first_iter ? (Q >= 3) : (epi ? (Q <= L + 2) : (kStaticTail || q <= L + 2));
        if (own_x && own_y) {
            ga2 = __ffma2_rn(p2, u2, ga2);
            gb2 = __fadd2_rn(gb2, __ffma2_rn(make_float2(-d2.x, -d2.y), cs, d2));
        } else if (own_y) {
            ga2.y = fmaf(p2.y, u2.y, ga2.y);
            gb2.y += fmaf(-d2.y, cs.y, d2.y);
        } else if (own_x) {
            ga2.x = fmaf(p2.x, u2.x, ga2.x);
            gb2.x += fmaf(-d2.x, cs.x, d2.x);
        }
        constexpr int LOW = (MODE == 0) ? 0 : -3;   // edge modes also keep dx_ext[-3..-1] (folded below)
        constexpr int HIGH = (MODE == 0) ? L : L + 3;   // ... and dx_ext[T..T+2] when the row ends in this segment
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (!(first_iter && Q - 1 - j < LOW) && !(epi && Q - 1 - j >= HIGH)) {
                float2& acc = ac[(Q - 1 - j + 4 * S) % S];
                acc = (j == 0) ? __fmul2_rn(tp.p.cu[0], du2) : __ffma2_rn(tp.p.cu[j], du2, acc);
            }
        }
        if (!(first_iter && Q < 6)) {
            const int o = q - 6;
            const float2 done = ac[(Q - 6 + 4 * S) % S];
            float yv = done.x + done.y;
            if (MODE != 0) {
                // adjoint of the x replicate pad: fold dx_ext beyond the row onto its first / last sample
                if (first_iter && Q == 6) {
                    if (t0 == 0) {
                        const float2 e1 = ac[(4 * S - 1) % S], e2 = ac[(4 * S - 2) % S], e3 = ac[(4 * S - 3) % S];
                        yv += (e3.x + e3.y) + (e2.x + e2.y) + (e1.x + e1.y);
                    }
                }
                if (t0 + o == Tlen - 1) {
                    const float2 e1 = ac[(Q - 5 + 4 * S) % S], e2 = ac[(Q - 4 + 4 * S) % S], e3 = ac[(Q - 3 + 4 * S) % S];
                    yv += (e1.x + e1.y) + (e2.x + e2.y) + (e3.x + e3.y);
                }
            }
            if (MODE != 2) {
                yb[(Q - 6 + 4 * S) % VEC] = yv;
                if ((Q - 6 + 4 * S) % VEC == VEC - 1) {
                    const int c0 = o - (VEC - 1);
                    if (first_iter && Q - 6 - (VEC - 1) < 8) {
#pragma unroll
                        for (int e = 0; e < VEC; ++e) hold[Q - 6 - (VEC - 1) + e] = yb[e];
                    } else if (MODE == 0 || t0 + c0 + VEC <= Tlen) {
                        io::store_chunk(seg + c0, yb, half);
                    } else if (t0 + c0 < Tlen) {
                        io::store_half_chunk(seg + c0, yb);      // half-aligned rows end mid-chunk
                    }
                }
            } else {
                if (first_iter && Q - 6 < 8) hold[Q - 6] = yv;
                else if (t0 + o < Tlen) io::store1(seg + o, yv);
            }
        }
    };

    constexpr int NR = kStaticTail ? L - S - 5 : L - S, NF = NR / S, RM = NR % S;
    static_assert(NR >= 0, "segment shorter than the ring");
#pragma unroll
    for (int q = -4; q < 6 + S; ++q) step(q, q, true, false);
    mid_walk_hook(pf, nc, next, lane);
#pragma unroll 1
    for (int it = 0; it < NF + (RM ? 1 : 0); ++it) {
        const int qb = 6 + S + it * S;
#pragma unroll
        for (int k = 0; k < S; ++k) {
            if (RM != 0 && k == RM) {
                if (it == NF) break;
            }
            step(6 + k, qb + k, false, false);
        }
    }
    if (kStaticTail) {
#pragma unroll
        for (int q = L + 1; q < L + 6; ++q) step(q, q, false, true);
    }
    __syncwarp(mask);
    if (MODE != 2) {
#pragma unroll
        for (int c = 0; c < 8 / VEC; ++c) {
            if (MODE == 0 || t0 + c * VEC + VEC <= Tlen) io::store_chunk(seg + c * VEC, &hold[c * VEC], half);
            else if (t0 + c * VEC < Tlen) io::store_half_chunk(seg + c * VEC, &hold[c * VEC]);
        }
    } else {
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (t0 + o < Tlen) io::store1(seg + o, hold[o]);
    }
    ga_out = (ga2.x + ga2.y) * ib;               // sum ds * ib * u * sin(2 a u)
    gb_out = -0.5f * ib * ib * (gb2.x + gb2.y);  // -sum ds * sin^2(a u) * ib^2,  sin^2 = (1 - cos 2au)/2
}

// ------------------------------------------------------------------------------------------------
// warp-tile bookkeeping
// ------------------------------------------------------------------------------------------------
struct TileDesc {
    int64_t flat_lo, flat_hi;  // this warp tile's output range in the flat [rows*T] array
    int64_t ld_lo, ld_hi;      // staged range (flat range + halo, clipped to the array; 16-byte aligned for the TMA)
    int64_t ld_extra;          // half-aligned rows: 4 trailing elements of the array the TMA cannot fetch, or -1
    int64_t row_base;          // row * T of this lane's row
    uint32_t gid, row;
    int32_t t0;
    bool active;               // this lane owns a segment
};

template <int L, bool HALF>
__device__ __forceinline__ TileDesc describe_tile(uint32_t wt, int lane, const Geometry& g) {
    TileDesc d;
    const uint32_t gid_first = wt * 32u;
    const uint32_t last_lane = min(31u, g.total_segs - 1u - gid_first);
    d.gid = gid_first + (uint32_t)lane;
    d.active = (uint32_t)lane <= last_lane;
    const uint32_t gidc = d.active ? d.gid : gid_first + last_lane;
    d.row = g.nseg.div(gidc);
    const uint32_t seg = gidc - d.row * g.nseg.d;
    d.t0 = (int32_t)(seg * (uint32_t)L);
    d.row_base = (int64_t)d.row * g.T;
    const int64_t lo = d.row_base + d.t0;
    const int64_t hi = d.row_base + min((int64_t)g.T, (int64_t)d.t0 + L);
    d.flat_lo = __shfl_sync(0xffffffffu, lo, 0);
    d.flat_hi = __shfl_sync(0xffffffffu, hi, (int)last_lane);
    d.ld_lo = max((int64_t)0, d.flat_lo - kHalo);
    d.ld_hi = min(g.total, d.flat_hi + kHalo);
    d.ld_extra = -1;
    if (HALF) {   // element offsets are multiples of 4 (8 bytes); the bulk copy needs multiples of 8 (16 bytes)
        d.ld_lo &= ~(int64_t)7;
        const int64_t up = (d.ld_hi + 7) & ~(int64_t)7;
        if (up <= g.total) {
            d.ld_hi = up;
        } else {                  // array ends 8 bytes past a 16-byte boundary: fetch those 8 bytes by hand
            d.ld_extra = g.total - 4;
            d.ld_hi = g.total - 4;
        }
    }
    return d;
}

template <typename T, int CH>
struct WarpTile {
    static constexpr int VEC = IO<T>::VEC;
    static constexpr int L = CH * VEC;
    static constexpr int kStageElems = 32 * L + 2 * kHalo + 8;           // one tensor, one stage (+8: 16-byte rounding of half-aligned tiles)
    static constexpr size_t kStageBytes = sizeof(T) * (size_t)kStageElems;  // multiple of 16
    static constexpr size_t fwd_smem(int nw) { return kBarBytes + (size_t)nw * AFA_FWD_STAGES * kStageBytes; }
    static constexpr size_t bwd_smem(int nw) { return kBarBytes + (size_t)nw * AFA_BWD_STAGES * 2 * kStageBytes; }
};

// Bulk-store a finished tile.  16-byte-aligned rows: one bulk TMA store.  Half-aligned rows (bf16, T % 8 == 4):
// the 16-byte-aligned body goes by bulk store, up to 4 leading / trailing elements by 8-byte stores.
template <typename T>
__device__ __forceinline__ void store_tile(T* gdst, const T* tile, const TileDesc& d, bool half, int lane) {
    if (!half) {
        if (lane == 0) {
            tma_store_1d(gdst + d.flat_lo, tile + (d.flat_lo - d.ld_lo), (uint32_t)((d.flat_hi - d.flat_lo) * (int64_t)sizeof(T)));
            tma_store_commit();
        }
        return;
    }
    const int64_t blo = (d.flat_lo + 7) & ~(int64_t)7, bhi = d.flat_hi & ~(int64_t)7;
    if (lane == 0) {
        if (bhi > blo) tma_store_1d(gdst + blo, tile + (blo - d.ld_lo), (uint32_t)((bhi - blo) * (int64_t)sizeof(T)));
        tma_store_commit();
    } else if (lane == 1) {
        if (blo > d.flat_lo && blo <= d.flat_hi)
            *reinterpret_cast<uint2*>(gdst + d.flat_lo) = *reinterpret_cast<const uint2*>(tile + (d.flat_lo - d.ld_lo));
    } else if (lane == 2) {
        if (bhi < d.flat_hi && bhi >= blo)
            *reinterpret_cast<uint2*>(gdst + bhi) = *reinterpret_cast<const uint2*>(tile + (bhi - d.ld_lo));
    }
}
// The 8 trailing bytes of the array that a 16-byte bulk copy cannot reach (half-aligned rows, last tile only).
template <typename T>
__device__ __forceinline__ void fetch_extra(T* tile, const T* gsrc, const TileDesc& d, int lane) {
    if (d.ld_extra >= 0 && lane == 0)
        *reinterpret_cast<uint2*>(tile + (d.ld_extra - d.ld_lo)) = *reinterpret_cast<const uint2*>(gsrc + d.ld_extra);
}

// ------------------------------------------------------------------------------------------------
// forward kernel: persistent, one autonomous pipeline per warp
// ------------------------------------------------------------------------------------------------
template <typename T, int CH, int NW, bool ALIGNED, bool HALF = false>
__global__ void __launch_bounds__(NW * 32, ALIGNED ? (sizeof(T) == 4 ? AFA_FWD_MINB_F32 : AFA_FWD_MINB_BF16) : 1) afa_fwd_kernel(const __grid_constant__ FwdArgs args) {
    using WT = WarpTile<T, CH>;
    constexpr int L = WT::L;
    static_assert(NW * 2 * sizeof(uint64_t) <= kBarBytes, "barrier block");
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * 2;
    constexpr int kStages = AFA_FWD_STAGES;
    T* stages = reinterpret_cast<T*>(smem + kBarBytes) + (size_t)warp * kStages * WT::kStageElems;
    const Geometry& g = args.g;
    const T* gx = static_cast<const T*>(args.x);
    T* gy = static_cast<T*>(args.y);

    const uint32_t GW = gridDim.x * NW;
    uint32_t wt = blockIdx.x * NW + warp;
    if (wt >= g.n_wtiles) return;

    if (ALIGNED) {
        if (lane == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    ChanParams cp;
    {
        const TileDesc d0 = describe_tile<L, HALF>(wt, lane, g);
        if (ALIGNED && lane == 0) {
            const uint32_t bytes = (uint32_t)((d0.ld_hi - d0.ld_lo) * (int64_t)sizeof(T));
            mbar_expect_tx(&bars[0], bytes);
            tma_load_1d(stages, gx + d0.ld_lo, bytes, &bars[0]);
        }
        cp = load_chan_params(args.alpha, args.beta, (int)(d0.row - g.chan.div(d0.row) * g.chan.d), g.flags);
    }
    uint32_t phase = 0;  // bit s = parity to wait for on stage s
    int st = 0;
    for (;;) {
        const uint32_t nwt = wt + GW;
        const bool has_next = nwt < g.n_wtiles;
        Prefetch pf{nullptr, nullptr, 0u, 0u, 0u, 0u};
        NextChan nc{args.alpha, args.beta, -1, g.flags};
        ChanParams cp_next = cp;
        if (has_next) {
            const TileDesc nxt = describe_tile<L, HALF>(nwt, lane, g);
            nc.c = (int32_t)(nxt.row - g.chan.div(nxt.row) * g.chan.d);
            if (ALIGNED) {
                pf.src0 = gx + nxt.ld_lo;
                pf.dst0 = smem_u32(stages + (size_t)((st ^ 1) % kStages) * WT::kStageElems);
                pf.bar = smem_u32(&bars[(st ^ 1) % kStages]);
                pf.bytes = (uint32_t)((nxt.ld_hi - nxt.ld_lo) * (int64_t)sizeof(T));
            }
        }
        // one stage: the refill of this same stage can only be issued after this tile's store has drained
        Prefetch late = pf;
        if (kStages == 1) pf.bytes = 0;
        const TileDesc cur = describe_tile<L, HALF>(wt, lane, g);
        T* tile = stages + (size_t)st * WT::kStageElems;
        const float a_eff = cp.a_eff, b_eff = cp.b_eff, ib = cp.ib;
        T* row0 = tile + (cur.row_base - cur.ld_lo);
        const uint32_t amask = __ballot_sync(0xffffffffu, cur.active);

        if (ALIGNED) {
            constexpr bool half = HALF;   // compile-time: 16-byte-aligned launches pay nothing for the 8-byte path
            mbar_wait(&bars[st], (phase >> st) & 1u);
            phase ^= (1u << st);
            if (half) {
                fetch_extra(tile, gx, cur, lane);
                __syncwarp();
            }
            const bool fast = cur.active && cur.t0 >= 5 && (cur.t0 + L + 5 < g.T);
            if (__all_sync(0xffffffffu, fast)) {
                walk_fwd<T, CH, 0>(row0, cur.t0, g.T, a_eff, ib, args.taps, pf, nc, cp_next, lane, 0xffffffffu, half);
            } else if (cur.active) {
                walk_fwd<T, CH, 1>(row0, cur.t0, g.T, a_eff, ib, args.taps, pf, nc, cp_next, lane, amask, half);
            }
            fence_proxy_async_smem();
            __syncwarp();
            store_tile(gy, tile, cur, half, lane);
            if (kStages == 1 && lane == 0) issue_prefetch(late);
        } else {
            const int n_in = (int)(cur.ld_hi - cur.ld_lo);
            for (int i = lane; i < n_in; i += 32) tile[i] = gx[cur.ld_lo + i];
            __syncwarp();
            if (cur.active) walk_fwd<T, CH, 2>(row0, cur.t0, g.T, a_eff, ib, args.taps, pf, nc, cp_next, lane, amask, false);
            __syncwarp();
            const int n_out = (int)(cur.flat_hi - cur.flat_lo);
            const T* src = tile + (cur.flat_lo - cur.ld_lo);
            for (int i = lane; i < n_out; i += 32) gy[cur.flat_lo + i] = src[i];
            __syncwarp();
        }
        if (!has_next) break;
        wt = nwt;
        cp = cp_next;
        if (kStages == 2) st ^= 1;
    }
    if (ALIGNED && lane == 0) tma_store_wait_read();  // shared memory must outlive the last bulk store's reads
}

// ------------------------------------------------------------------------------------------------
// backward kernel: same frame, two staged tensors (x, gy), gx written in place over x
// ------------------------------------------------------------------------------------------------
template <typename T, int CH, int NW, bool ALIGNED, bool HALF = false>
__global__ void __launch_bounds__(NW * 32, ALIGNED ? AFA_BWD_MINB : 1) afa_bwd_kernel(const __grid_constant__ BwdArgs args) {
    using WT = WarpTile<T, CH>;
    constexpr int L = WT::L;
    static_assert(NW * 2 * sizeof(uint64_t) <= kBarBytes, "barrier block");
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * 2;
    // per warp: [stage 0: x | gy][stage 1: x | gy]
    constexpr int kStages = AFA_BWD_STAGES;
    T* stages = reinterpret_cast<T*>(smem + kBarBytes) + (size_t)warp * kStages * 2 * WT::kStageElems;
    const Geometry& g = args.g;
    const T* px = static_cast<const T*>(args.x);
    const T* pg = static_cast<const T*>(args.gy);
    T* po = static_cast<T*>(args.gx);

    const uint32_t GW = gridDim.x * NW;
    uint32_t wt = blockIdx.x * NW + warp;
    // the workspace is not zeroed by the caller: the counters of the finalize kernel (next launch on this stream) start at 0
    if (blockIdx.x == 0 && args.cnt != nullptr)
        for (int i = threadIdx.x; i < args.n_cnt; i += NW * 32) args.cnt[i] = 0u;
    if (wt >= g.n_wtiles) return;

    if (ALIGNED) {
        if (lane == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    ChanParams cp;
    {
        const TileDesc d0 = describe_tile<L, HALF>(wt, lane, g);
        if (ALIGNED && lane == 0) {
            const uint32_t bytes = (uint32_t)((d0.ld_hi - d0.ld_lo) * (int64_t)sizeof(T));
            mbar_expect_tx(&bars[0], 2 * bytes);
            tma_load_1d(stages, px + d0.ld_lo, bytes, &bars[0]);
            tma_load_1d(stages + WT::kStageElems, pg + d0.ld_lo, bytes, &bars[0]);
        }
        cp = load_chan_params(args.alpha, args.beta, (int)(d0.row - g.chan.div(d0.row) * g.chan.d), g.flags);
    }
    uint32_t phase = 0;
    int st = 0;
    for (;;) {
        const uint32_t nwt = wt + GW;
        const bool has_next = nwt < g.n_wtiles;
        Prefetch pf{nullptr, nullptr, 0u, 0u, 0u, 0u};
        NextChan nc{args.alpha, args.beta, -1, g.flags};
        ChanParams cp_next = cp;
        if (has_next) {
            const TileDesc nxt = describe_tile<L, HALF>(nwt, lane, g);
            nc.c = (int32_t)(nxt.row - g.chan.div(nxt.row) * g.chan.d);
            if (ALIGNED) {
                pf.src0 = px + nxt.ld_lo;
                pf.src1 = pg + nxt.ld_lo;
                pf.dst0 = smem_u32(stages + (size_t)((st ^ 1) % kStages) * 2 * WT::kStageElems);
                pf.dst1 = pf.dst0 + (uint32_t)WT::kStageBytes;
                pf.bar = smem_u32(&bars[(st ^ 1) % kStages]);
                pf.bytes = (uint32_t)((nxt.ld_hi - nxt.ld_lo) * (int64_t)sizeof(T));
            }
        }
        Prefetch late = pf;                      // one stage: refill only after this tile's store has drained
        if (kStages == 1) pf.bytes = 0;
        const TileDesc cur = describe_tile<L, HALF>(wt, lane, g);
        T* tile_x = stages + (size_t)st * 2 * WT::kStageElems;
        T* tile_g = tile_x + WT::kStageElems;
        const float a_eff = cp.a_eff, b_eff = cp.b_eff, ib = cp.ib;
        T* row0 = tile_x + (cur.row_base - cur.ld_lo);
        const T* grow0 = tile_g + (cur.row_base - cur.ld_lo);
        float ga = 0.f, gb = 0.f;
        const uint32_t amask = __ballot_sync(0xffffffffu, cur.active);

        if (ALIGNED) {
            constexpr bool half = HALF;   // compile-time: 16-byte-aligned launches pay nothing for the 8-byte path
            mbar_wait(&bars[st], (phase >> st) & 1u);
            phase ^= (1u << st);
            if (half) {
                fetch_extra(tile_x, px, cur, lane);
                fetch_extra(tile_g, pg, cur, lane);
                __syncwarp();
            }
            const bool fast = cur.active && cur.t0 >= 5 && (cur.t0 + L + 5 < g.T);
            if (__all_sync(0xffffffffu, fast)) {
                walk_bwd<T, CH, 0>(row0, grow0, cur.t0, g.T, a_eff, ib, args.taps, pf, nc, cp_next, lane, 0xffffffffu, half, ga, gb);
            } else if (cur.active) {
                walk_bwd<T, CH, 1>(row0, grow0, cur.t0, g.T, a_eff, ib, args.taps, pf, nc, cp_next, lane, amask, half, ga, gb);
            }
            fence_proxy_async_smem();
            __syncwarp();
            store_tile(po, tile_x, cur, half, lane);
            if (kStages == 1 && lane == 0) issue_prefetch(late);
        } else {
            const int n_in = (int)(cur.ld_hi - cur.ld_lo);
            for (int i = lane; i < n_in; i += 32) {
                tile_x[i] = px[cur.ld_lo + i];
                tile_g[i] = pg[cur.ld_lo + i];
            }
            __syncwarp();
            if (cur.active) walk_bwd<T, CH, 2>(row0, grow0, cur.t0, g.T, a_eff, ib, args.taps, pf, nc, cp_next, lane, amask, false, ga, gb);
            __syncwarp();
            const int n_out = (int)(cur.flat_hi - cur.flat_lo);
            const T* src = tile_x + (cur.flat_lo - cur.ld_lo);
            for (int i = lane; i < n_out; i += 32) po[cur.flat_lo + i] = src[i];
            __syncwarp();
        }
        if (cur.active) {
            // gradients w.r.t. the RAW parameters: d exp(p)/dp = exp(p)        activations.py:121-123
            if (g.flags & 1) {
                ga *= a_eff;
                gb *= b_eff;
            }
            args.part[cur.gid] = ga;
            args.part[(size_t)g.total_segs + cur.gid] = gb;
        }
        if (!has_next) break;
        wt = nwt;
        cp = cp_next;
        if (kStages == 2) st ^= 1;
    }
    if (ALIGNED && lane == 0) tma_store_wait_read();
}

// Second stage of the deterministic parameter-gradient reduction.  Channel c owns batch * nseg per-segment partials
// (runs of nseg consecutive values, one run per batch entry).  grid = (C, split): CTA (c, k) sums the k-th contiguous
// slice of the channel's flattened (batch, segment) range -- four independent thread-strided serial sums per thread (four
// loads in flight), combined in a fixed order, then a fixed shuffle / shared-memory tree -- and, with split > 1, leaves its two
// sums in part2; the CTA that arrives LAST on the channel's counter adds the `split` slice sums in slice order.  Every sum
// has a fixed order whatever the timing, so results are run-to-run identical.  (One CTA per channel -- the first version --
// left 24 CTAs walking 68 k partials each at C = 24: 40 us behind a 210 us backward kernel.)
constexpr int kFinalizeThreads = 256;
constexpr int kFinalizeMaxSplit = 32;
__global__ void __launch_bounds__(kFinalizeThreads)
afa_param_grad_finalize(const float* __restrict__ part, float* __restrict__ galpha, float* __restrict__ gbeta,
                        uint32_t total_segs, uint32_t nseg, int batch, int C, int snake, float* part2, uint32_t* cnt) {
    const int c = blockIdx.x;
    const int split = gridDim.y, k = blockIdx.y;
    const int tid = threadIdx.x;
    const uint32_t per_channel = (uint32_t)batch * nseg;
    const uint32_t slice = (per_channel + split - 1) / split;
    const uint32_t lo = min(per_channel, (uint32_t)k * slice), hi = min(per_channel, lo + slice);
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
    auto at = [&](uint32_t i) {
        const uint32_t b = i / nseg, s = i - b * nseg;
        return ((size_t)b * C + c) * nseg + s;
    };
    uint32_t i = lo + tid;
    for (; i + 3 * kFinalizeThreads < hi; i += 4 * kFinalizeThreads) {
        float va[4], vb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const size_t idx = at(i + j * kFinalizeThreads);
            va[j] = __ldcs(part + idx);
            vb[j] = __ldcs(part + (size_t)total_segs + idx);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sa[j] += va[j];
            sb[j] += vb[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {                              // at most three left: they continue the sums j = 0, 1, 2
        if (i + j * kFinalizeThreads < hi) {
            const size_t idx = at(i + j * kFinalizeThreads);
            sa[j] += __ldcs(part + idx);
            sb[j] += __ldcs(part + (size_t)total_segs + idx);
        }
    }
    float ta = (sa[0] + sa[1]) + (sa[2] + sa[3]), tb = (sb[0] + sb[1]) + (sb[2] + sb[3]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ta += __shfl_xor_sync(0xffffffffu, ta, off);
        tb += __shfl_xor_sync(0xffffffffu, tb, off);
    }
    __shared__ float red[2][kFinalizeThreads / 32];
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = ta;
        red[1][tid >> 5] = tb;
    }
    __syncthreads();
    if (tid != 0) return;
    ta = 0.f;
    tb = 0.f;
#pragma unroll
    for (int w = 0; w < kFinalizeThreads / 32; ++w) {
        ta += red[0][w];
        tb += red[1][w];
    }
    if (split > 1) {
        float* mine = part2 + ((size_t)c * split + k) * 2;
        __stcg(mine, ta);
        __stcg(mine + 1, tb);
        __threadfence();
        if (atomicAdd(cnt + c, 1u) != (uint32_t)split - 1u) return;
        __threadfence();
        ta = 0.f;
        tb = 0.f;
        for (int q = 0; q < split; ++q) {
            ta += __ldcg(part2 + ((size_t)c * split + q) * 2);
            tb += __ldcg(part2 + ((size_t)c * split + q) * 2 + 1);
        }
        cnt[c] = 0u;
    }
    if (snake) {
        galpha[c] = ta + tb;  // beta aliases alpha                              activations.py:57-60
    } else {
        galpha[c] = ta;
        gbeta[c] = tb;
    }
}

}  // namespace afa
