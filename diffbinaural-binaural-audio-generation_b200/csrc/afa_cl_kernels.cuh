// afa_cl_kernels.cuh -- the AMP-block kernels on CHANNELS-LAST activations ([batch, T, channels],
// channels contiguous) for sm_100a.
//
// Why a second layout: the dense convolutions either side of every Activation1d run on cuDNN's NHWC
// tensor-core kernels; in the reference's [B, C, T] layout every convolution is wrapped in two layout
// conversion kernels and followed by a separate bias kernel (measured: 37 % of the generator pass,
// profiles/r01_vocoder_breakdown.log).  Keeping the activations channels-last end to end removes the
// conversions, and the prologue below removes the bias and residual kernels:
//
//   afa_cl_fwd_kernel      x' = x (+ res);  [xsum = x';]  y = down2x(snake(up2x(x' + bias[c])))
//                          reference: Conv1d bias of c1/c2 + `x = xt + x` + the next Activation1d
//                          (BigVGAN/bigvgan.py:132-141, act.py:25-30)
//   afa_cl_tail_kernel     activation_post -> conv_post (C -> 1, k = 7, zero pad) -> clamp | tanh
//                          -> float wave and/or int16 PCM (interleaved)
//                          (BigVGAN/bigvgan.py:379-385, inference_e2e.py:193-205)
//   afa_mean_kernel        x = (sum_j (xt_j + bias_j + x_j)) / num_kernels        (bigvgan.py:368-376)
//
// Layout of the walk: one THREAD owns one (batch, channel, time-segment); the 32 lanes of a warp are 32
// consecutive channels of the same time step, so every global access of a warp is one contiguous
// 64-byte (bf16) / 128-byte (fp32) run -- no shared memory, no halo exchange: the 5-sample halo of a
// segment is simply re-read (L1/L2 hits, the neighbouring segment's lines).  Inputs arrive through a
// register prefetch ring one loop trip (12 steps) ahead of their use.  The arithmetic of a step is the
// same packed-pair transposed-form walk as the [B, C, T] kernel (afa_kernels.cuh, walk_fwd), with the
// bias folded into the initial value of each pending upsampler pair (the up-filter is linear and
// replicate padding keeps all six taps of a phase: u(x + b) = u(x) + b * sum(taps of the phase)).
#pragma once
#include "afa_kernels.cuh"

namespace afa {

constexpr int kClS = 12;        // ring size = steps per rolled-loop trip = prefetch distance
// 1: every request of a group also asks L2 for the group after it (prefetch.global.L2; interior walks only)
#ifndef AFA_CL_L2_PREFETCH
#define AFA_CL_L2_PREFETCH 0
#endif
constexpr int kClThreads = 128;
// resident CTAs per SM the channels-last walk is compiled for (register cap = 65536 / (128 * n))
#ifndef AFA_CL_MINB_PLAIN
#define AFA_CL_MINB_PLAIN 5
#endif
#ifndef AFA_CL_MINB_RES
#define AFA_CL_MINB_RES 4
#endif

struct ClArgs {
    const void* x;
    const void* res;      // optional second addend (residual stream), same shape
    void* xsum;           // with res: x + res is written here (the new residual stream; `bias` stays pending)
    void* y;
    const float* bias;    // optional [C] fp32
    const float* alpha;
    const float* beta;
    FwdTaps taps;
    int64_t x_bs, res_bs, xsum_bs, y_bs;   // batch strides in elements (rows may be padded in time)
    uint32_t total;       // batch * nseg * C threads
    FastDiv chan, batch;
    uint32_t nseg;
    int32_t T, L, y_tpad, flags;
};

struct TailArgs {
    const void* x;
    const float* alpha;
    const float* beta;
    const float* w;       // conv_post weight [C][7] fp32
    const float* bias;    // conv_post bias (1 value) or nullptr
    float* wave;          // optional [batch][T] fp32
    int16_t* pcm;         // optional int16, element (b, t) at ((b / il) * T + t) * il + b % il
    FwdTaps taps;
    int64_t x_bs;
    uint32_t total_warps; // batch * nseg
    FastDiv nseg;
    int32_t T, C, L, flags, use_tanh, il;
    float pcm_scale;
    // optional scatter of whole frames (zero-frame restoration, inference_e2e.py:80-111): sample t of batch entry b
    // lands at frame_map[b * n_frames + t / hop] * hop + t % hop of an output row of T_out samples
    const int32_t* frame_map;
    int32_t hop, n_frames;
    int64_t T_out;
};

template <typename T> __device__ __forceinline__ float cl_load(const T* p);
template <> __device__ __forceinline__ float cl_load<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float cl_load<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}
// The prefetch ring holds RAW loaded words; they become floats only at the step that consumes them.  The
// conversion is `asm volatile` on purpose: written as plain C++ the compiler hoists all 12 (24) shifts of a trip
// to the top of the loop body, so every load of the previous trip -- including the one issued last -- must have
// landed there (ncu: one SHF carried 10 % of all stall samples, long_scoreboard).  Pinned, each load keeps its
// full 12-step lead.
template <typename T> struct ClRaw;
template <> struct ClRaw<float> {
    using raw_t = float;
    static __device__ __forceinline__ raw_t load(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ raw_t load_smem(const float* p) { return *p; }
    static __device__ __forceinline__ float cvt(raw_t v) {
        float o;
        asm volatile("mov.b32 %0, %1;" : "=f"(o) : "f"(v));
        return o;
    }
};
template <> struct ClRaw<__nv_bfloat16> {
    using raw_t = uint32_t;
    static __device__ __forceinline__ raw_t load(const __nv_bfloat16* p) {
        return (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p));
    }
    static __device__ __forceinline__ raw_t load_smem(const __nv_bfloat16* p) {       // p points into shared memory
        uint32_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)));
        return v;
    }
    static __device__ __forceinline__ float cvt(raw_t v) {
        uint32_t o;
        asm volatile("shl.b32 %0, %1, 16;" : "=r"(o) : "r"(v));
        return __uint_as_float(o);
    }
};
template <typename T> __device__ __forceinline__ void cl_store(T* p, float v);
template <> __device__ __forceinline__ void cl_store<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void cl_store<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// conv_post state of one lane (= one channel) in the tail kernel
struct TailSink {
    float w[7];
    float pc[kClS];
    float bias;
    float* wave;      // row of this batch entry (lane 0 stores)
    int16_t* pcm;
    const int32_t* fmap;   // this batch entry's frame map, or nullptr
    int64_t t_out;         // samples per output row: a frame index outside [0, t_out / hop) drops its samples (the reference
                           // clamps every copy to the original length, inference_e2e.py:94-109)
    int hop;
    int il;
    float pcm_scale;
    int use_tanh;
    int lane;
};

// shared-memory tile of the fused activation -> convolution kernel (afa_actconv_kernels.cuh): this lane's column
struct TileSink {
    __nv_bfloat16* col;   // element of tile row r at col[r * stride]
    int stride;           // elements per tile row
    int tile_t0;          // time index of tile row 0
    int own_lo, own_hi;   // x + res is stored to xsum only for rows in [own_lo, own_hi)
};

// ------------------------------------------------------------------------------------------------
// The walk of one (batch, channel, segment).  px/pr/ps/py point at this channel's column of batch b
// (element t at p[t * Cs]).  Step q (-4 <= q <= L+5) consumes x'[t0+q-1], handles the 2x-rate pair
// (s[2m-1], s[2m]), m = t0-3+q, and completes y[t0+q-6].  L = 12 n + 2: the L + 10 steps are n + 1 GROUPS
// of 12; group g covers steps q in [12g-4, 12g+8).
//
// Memory schedule (what the ncu captures of the first version asked for): the 12 inputs of group g+1 are
// requested in ONE burst at the top of group g, right after group g's own raw words have been turned into
// floats, so every load has a whole group (12 steps, ~2000 cycles under load) to land.  Written with one load
// per step, the compiler clustered the conversions at the top of the loop body anyway and left the loads
// spread over the first half of the body: the last ones had 40 % of a trip to land and one SHF carried 10 %
// of all stall samples (long_scoreboard).
//
// MODE 0: the whole reach [t0-5, t0+L+4] lies inside the row (branch-free); MODE 1: anything else (index
// clamps = replicate pad of x, selects for the replicate pad of the activated signal).
// SINK 0: store y;  SINK 1: feed conv_post (tail kernel; all 32 lanes walk in lockstep);  SINK 2: write the
// activated sample as bf16 into a shared-memory tile (zero outside the row: the convolution's zero padding).
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE, bool RES, int SINK, bool XSM = false>
__device__ __forceinline__ void walk_cl(const T* __restrict__ px, const T* __restrict__ pr, T* __restrict__ ps,
                                        T* __restrict__ py, const int Cs, const int t0, const int L, const int Tlen,
                                        const float a, const float ib, const float bias, const FwdTaps& tp,
                                        TailSink* sink, const uint32_t mask, const TileSink* ts = nullptr) {
    constexpr int S = kClS;
    using raw = ClRaw<T>;
    float2 up[S], ac[S];
    typename raw::raw_t xq[S], rq[RES ? S : 1];   // raw words of the NEXT group, in flight
    float xf[S];                                   // inputs of the CURRENT group: x (+ res)
    float s_first = 0.f, s_last = 0.f;

    // request the 12 inputs x[tb .. tb+11] (group slots 0..11)
    auto request = [&](const int tb) {
        if (MODE == 0) {
            const T* p = px + (int64_t)tb * Cs;
            const T* q = RES ? pr + (int64_t)tb * Cs : nullptr;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                xq[i] = XSM ? raw::load_smem(p) : raw::load(p);      // XSM: x was staged in shared memory by a bulk copy
                p += Cs;
                if (RES) { rq[i] = raw::load(q); q += Cs; }
            }
            if (AFA_CL_L2_PREFETCH && SINK == 0) {     // p / q now point at the group after the one just requested
#pragma unroll
                for (int i = 0; i < S; i += AFA_CL_L2_PREFETCH) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                    p += AFA_CL_L2_PREFETCH * Cs;
                    if (RES) { asm volatile("prefetch.global.L2 [%0];" ::"l"(q)); q += AFA_CL_L2_PREFETCH * Cs; }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < S; ++i) {
                const int64_t o = (int64_t)min(max(tb + i, 0), Tlen - 1) * Cs;
                xq[i] = XSM ? raw::load_smem(px + o) : raw::load(px + o);
                if (RES) rq[i] = raw::load(pr + o);
            }
        }
    };
    auto land = [&]() {
#pragma unroll
        for (int i = 0; i < S; ++i) {
            xf[i] = raw::cvt(xq[i]);
            if (RES) xf[i] += raw::cvt(rq[i]);
        }
    };
    request(t0 - 5);                                   // group 0: steps -4..7
    T* pyr = (SINK == 0) ? py + (int64_t)t0 * Cs : nullptr;
    T* psr = (RES && SINK != 2) ? ps + (int64_t)t0 * Cs : nullptr;
    __nv_bfloat16* ptile = nullptr;
    if constexpr (SINK == 2) ptile = ts->col + (t0 - ts->tile_t0) * ts->stride;

    // bias folded into the initial value of every pending upsampler pair
    float2 bias2;
    {
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) { sx += tp.p.cu[j].x; sy += tp.p.cu[j].y; }
        bias2 = make_float2(bias * sx, bias * sy);
    }
    if (MODE != 0) {
        auto xval = [&](int t) -> float {
            const int64_t o = (int64_t)min(max(t, 0), Tlen - 1) * Cs;
            float v = XSM ? raw::cvt(raw::load_smem(px + o)) : cl_load(px + o);
            if (RES) v += cl_load(pr + o);
            return v + bias;
        };
        if (t0 <= 2) {  // s[0], which the left replicate pad of the activated signal repeats       filter.py:98
            const float x0 = xval(0), x1 = xval(1), x2 = xval(2);
            float u = tp.ue[0] * x2;
            u = fmaf(tp.ue[1], x1, u);
#pragma unroll
            for (int j = 2; j < 6; ++j) u = fmaf(tp.ue[j], x0, u);
            s_first = snake_f(u, a, ib);
        }
        if (Tlen - 1 <= t0 + L + 5) {  // s[2T-1], which the right pad repeats
            float u = 0.f;
#pragma unroll
            for (int j = 0; j < 6; ++j) u = fmaf(tp.uo[j], xval(Tlen + 2 - j), u);
            s_last = snake_f(u, a, ib);
        }
    }

    // Q: static part of the step number (ring slots); `own`: x'[t0+q-1] belongs to this segment (xsum store)
    auto step = [&](const int Q, const int q, const bool first_iter, const bool own) {
        const float xv = xf[(Q + 4 + 4 * S) % S];
        if (RES) {
            if constexpr (SINK == 2) {        // tiles overlap by the convolution's reach: store only this tile's own rows
                const int t = t0 + q - 1;
                if (!(first_iter && Q < 1) && t >= ts->own_lo && t < ts->own_hi) cl_store(ps + (int64_t)t * Cs, xv);
            } else if (!(first_iter && Q < 1) && own) {
                if (MODE == 0 || t0 + q - 1 < Tlen) cl_store(psr, xv);
                psr += Cs;
            }
        }
        // --- upsampler, transposed form                                            resample.py:32-36
        const float2 xx = bcast2(xv);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (!(first_iter && Q + j < 1)) {
                float2& pend = up[(Q + j + 4 * S) % S];
                pend = __ffma2_rn(tp.p.cu[j], xx, (j == 5) ? bias2 : pend);
            }
        }
        if (first_iter && Q < 1) return;
        // --- Snake / SnakeBeta on the finished pair                                 activations.py:124
        const float2 u2 = up[(Q + 4 * S) % S];
        const float2 th = __fmul2_rn(u2, bcast2(a));
        const float2 sn = make_float2(__sinf(th.x), __sinf(th.y));
        float2 s2 = __ffma2_rn(bcast2(ib), __fmul2_rn(sn, sn), u2);
        if (MODE != 0) {
            const int m = t0 - 3 + q;
            if (m <= 0) s2.x = s_first;      // s[n], n < 0, repeats s[0]
            if (m < 0) s2.y = s_first;
            if (m > Tlen) s2.x = s_last;     // s[n], n >= 2T, repeats s[2T-1]
            if (m >= Tlen) s2.y = s_last;
        }
        // --- low-pass, transposed form                                              filter.py:98-99
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (!(first_iter && Q - 1 - j < 0)) {
                float2& acc = ac[(Q - 1 - j + 4 * S) % S];
                acc = (j == 0) ? __fmul2_rn(tp.p.cd[0], s2) : __ffma2_rn(tp.p.cd[j], s2, acc);
            }
        }
        if (first_iter && Q < 6) return;
        // --- y[t0 + q - 6] is complete
        const float2 done = ac[(Q - 6 + 4 * S) % S];
        const float yv = done.x + done.y;
        if constexpr (SINK == 0) {
            if (MODE == 0 || t0 + q - 6 < Tlen) cl_store(pyr, yv);
            pyr += Cs;
        } else if constexpr (SINK == 2) {
            const int tv = t0 + q - 6;
            *ptile = __float2bfloat16_rn((MODE == 0 || (tv >= 0 && tv < Tlen)) ? yv : 0.f);
            ptile += ts->stride;
        } else {
            // conv_post (k = 7, zero padding), transposed form: a[tv] feeds out[tv+3-j] with w[j]   bigvgan.py:380
            const int tv = t0 + q - 6;
            const float av = (tv >= 0 && tv < Tlen) ? yv : 0.f;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                if (!(first_iter && Q - j < 6)) {
                    float& p = sink->pc[(Q - 3 - j + 4 * S) % S];
                    p = (j == 0) ? sink->w[0] * av : fmaf(sink->w[j], av, p);
                }
            }
            if (first_iter && Q < 12) return;
            float v = sink->pc[(Q - 9 + 4 * S) % S];       // out[tv - 3] has all seven taps of this channel
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            const int to = tv - 3;
            if (sink->lane == 0 && to < Tlen) {
                v += sink->bias;
                v = sink->use_tanh ? tanhf(v) : fminf(fmaxf(v, -1.0f), 1.0f);     // bigvgan.py:382-385
                int64_t pos = to;
                bool inside = true;
                if (sink->fmap) {
                    const int f = to / sink->hop;
                    const int64_t fm = __ldg(sink->fmap + f);
                    pos = fm * sink->hop + (to - f * sink->hop);
                    inside = fm >= 0 && pos < sink->t_out;
                }
                if (inside && sink->wave) sink->wave[pos] = v;
                if (inside && sink->pcm) sink->pcm[pos * sink->il] = (int16_t)(v * sink->pcm_scale);   // astype("int16") truncates
            }
        }
    };

    // group 0 (steps -4..7) and group 1 (steps 8..19) are straight-line code (the warm-up and the first outputs have
    // static special cases); groups 2..n run in the rolled loop.  Every group body is
    //     request(group g+1)  ->  12 steps on xf  ->  land() (raw words of g+1 -> xf)
    // i.e. the loads' only consumer sits at the END of the body: the scheduler then issues them first and they
    // have the whole body to land.  (With the consumer at the top of the next body, ptxas sank the loads to the
    // bottom of this one -- zero lead -- and a __syncwarp "fence" only pinned their order among memory operations.)
    const int n_groups = (L + 10) / S;                 // L = 12 n + 2  ->  n + 1 >= 2 groups
    land();
    request(t0 - 5 + S);
#pragma unroll
    for (int q = -4; q < 8; ++q) step(q, q, true, q <= L);
    land();
    if (n_groups > 2) request(t0 - 5 + 2 * S);
#pragma unroll
    for (int q = 8; q < 8 + S; ++q) step(q, q, true, q <= L);
    if (n_groups > 2) land();
    // the rolled loop has NO branch inside (one basic block: loads, steps, conversions); the last group is peeled
#pragma unroll 1
    for (int g = 2; g < n_groups - 1; ++g) {
        const int qb = g * S - 4;
        request(t0 - 5 + (g + 1) * S);
#pragma unroll
        for (int k = 0; k < S; ++k) step(8 + k, qb + k, false, true);
        land();
    }
    if (n_groups > 2) {
        const int qb = (n_groups - 1) * S - 4;
#pragma unroll
        for (int k = 0; k < S; ++k) step(8 + k, qb + k, false, k < S - 5);
    }
}

// ------------------------------------------------------------------------------------------------
// forward kernel: one thread per (batch, segment, channel), channel fastest
// ------------------------------------------------------------------------------------------------
template <typename T, bool RES>
__global__ void __launch_bounds__(kClThreads, RES ? AFA_CL_MINB_RES : AFA_CL_MINB_PLAIN) afa_cl_fwd_kernel(const __grid_constant__ ClArgs args) {
    const uint32_t g = blockIdx.x * kClThreads + threadIdx.x;
    const bool active = g < args.total;
    const uint32_t gc = active ? g : args.total - 1u;
    // thread order: channel fastest, then batch, then segment -- with the LAST segment of the rows first and
    // the first one second: the edge-mode walks (and the zero fill) are the slow ones, so they must not be the
    // last CTAs of the grid (ncu: one straggling SM doubled the launch time when they were)
    const uint32_t sc = args.chan.div(gc);              // slot * batch + b
    const uint32_t c = gc - sc * args.chan.d;
    const uint32_t slot = args.batch.div(sc);
    const uint32_t b = sc - slot * args.batch.d;
    const uint32_t s = slot == 0 ? args.nseg - 1u : slot - 1u;
    const int L = args.L, Tlen = args.T, Cs = (int)args.chan.d;
    const int t0 = (int)s * L;

    const T* px = static_cast<const T*>(args.x) + (int64_t)b * args.x_bs + c;
    const T* pr = RES ? static_cast<const T*>(args.res) + (int64_t)b * args.res_bs + c : nullptr;
    T* ps = RES ? static_cast<T*>(args.xsum) + (int64_t)b * args.xsum_bs + c : nullptr;
    T* py = static_cast<T*>(args.y) + (int64_t)b * args.y_bs + c;
    const ChanParams cp = load_chan_params(args.alpha, args.beta, (int)c, args.flags);
    const float bias = args.bias ? __ldg(args.bias + c) : 0.f;

    const bool fast = !active || (t0 >= 5 && t0 + L + 5 + (AFA_CL_L2_PREFETCH ? kClS : 0) < Tlen);
    const uint32_t amask = __ballot_sync(0xffffffffu, active);
    if (__all_sync(0xffffffffu, fast)) {
        if (active) walk_cl<T, 0, RES, 0>(px, pr, ps, py, Cs, t0, L, Tlen, cp.a_eff, cp.ib, bias, args.taps, nullptr, amask);
    } else if (active) {
        walk_cl<T, 1, RES, 0>(px, pr, ps, py, Cs, t0, L, Tlen, cp.a_eff, cp.ib, bias, args.taps, nullptr, amask);
    }
    // rows [T, y_tpad) of y are the zero padding a polyphase (dilated) convolution reads next
    if (active && s == args.nseg - 1u) {
        for (int t = Tlen; t < args.y_tpad; ++t) cl_store(py + (int64_t)t * Cs, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// tail kernel: one WARP per (batch, segment), lane = channel (C <= 32).  The walk covers the
// segment's outputs plus the 3 activated samples either side that conv_post (k = 7) reaches.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kClThreads, 4) afa_cl_tail_kernel(const __grid_constant__ TailArgs args) {
    const int lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * (kClThreads / 32) + (threadIdx.x >> 5);
    if (w >= args.total_warps) return;                   // whole warps leave together
    const uint32_t b = args.nseg.div(w);
    const uint32_t s = w - b * args.nseg.d;
    const int C = args.C, Tlen = args.T;
    const int c = min(lane, C - 1);                      // lanes >= C shadow the last channel with zero weights
    const int Lout = args.L - 6;                         // outputs per segment; the walk is 6 longer
    const int t0 = (int)s * Lout - 3;

    const T* px = static_cast<const T*>(args.x) + (int64_t)b * args.x_bs + c;
    const ChanParams cp = load_chan_params(args.alpha, args.beta, c, args.flags);
    TailSink sink;
#pragma unroll
    for (int j = 0; j < 7; ++j) sink.w[j] = (lane < C) ? __ldg(args.w + c * 7 + j) : 0.f;
    sink.bias = args.bias ? __ldg(args.bias) : 0.f;
    sink.wave = args.wave ? args.wave + (int64_t)b * args.T_out : nullptr;
    sink.il = args.il;
    sink.pcm = args.pcm ? args.pcm + ((int64_t)(b / (uint32_t)args.il) * args.T_out) * args.il + (b % (uint32_t)args.il) : nullptr;
    sink.fmap = args.frame_map ? args.frame_map + (int64_t)b * args.n_frames : nullptr;
    sink.t_out = args.T_out;
    sink.hop = args.hop;
    sink.pcm_scale = args.pcm_scale;
    sink.use_tanh = args.use_tanh;
    sink.lane = lane;
    walk_cl<T, 1, false, 1>(px, nullptr, nullptr, nullptr, C, t0, args.L, Tlen, cp.a_eff, cp.ib, 0.f, args.taps, &sink, 0xffffffffu);
}

// ------------------------------------------------------------------------------------------------
// resblock mean: out = scale * (sum_j (y_j + r_j) + bias_sum[c]) over dense [rows, C] arrays
// ------------------------------------------------------------------------------------------------
constexpr int kMeanMaxK = 4;
struct MeanArgs {
    const void* y[kMeanMaxK];
    const void* r[kMeanMaxK];
    const float* bias_sum;   // optional [C] fp32: sum of the K convolution biases
    void* out;
    int64_t n;               // rows * C elements
    FastDiv cvec;            // C / VEC
    int32_t K;
    float scale;
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256) afa_mean_kernel(const __grid_constant__ MeanArgs a) {
    const int64_t nv = a.n / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        float acc[VEC];
        const uint32_t cv = (uint32_t)i - a.cvec.div((uint32_t)i) * a.cvec.d;
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = a.bias_sum ? __ldg(a.bias_sum + cv * VEC + e) : 0.f;
        for (int j = 0; j < a.K; ++j) {
            float v[VEC], u[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) u[e] = 0.f;
            if (VEC == 1) {
                v[0] = cl_load(static_cast<const T*>(a.y[j]) + i);
                if (a.r[j]) u[0] = cl_load(static_cast<const T*>(a.r[j]) + i);
            } else {
                IO<T>::load_chunk(static_cast<const T*>(a.y[j]) + i * VEC, v);
                if (a.r[j]) IO<T>::load_chunk(static_cast<const T*>(a.r[j]) + i * VEC, u);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += v[e] + u[e];
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] *= a.scale;
        if (VEC == 1) cl_store(static_cast<T*>(a.out) + i, acc[0]);
        else IO<T>::store_chunk(static_cast<T*>(a.out) + i * VEC, acc);
    }
}

}  // namespace afa
