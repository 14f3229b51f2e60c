// afa_mel.cu -- fused log-mel spectrogram for sm_100a: forward and the adjoint the
// training loss needs.
//
// One launch replaces the torch-op chain of the reference's `mel_spectrogram`
// (BigVGAN/meldataset.py:51-123: reflect pad -> torch.stft(hann) -> sqrt(re^2 + im^2 + 1e-9) -> mel_basis @ spec
// -> log(clamp(., 1e-5))) and of `MultiScaleMelSpectrogramLoss.mel_spectrogram` + its log10
// (BigVGAN/loss.py:110-167, 195-200), which differ only in pad width, magnitude epsilon, clamp and log base.
//
// Per CTA: a run of consecutive STFT frames of one waveform row (several in flight at once for the small windows, see
// Slots).  Each frame is windowed while it is gathered from
// global memory (replicated / reflected borders resolved by index arithmetic, no padded copy), packed as N/2 complex
// points, transformed by a radix-4 (+ one radix-2 stage for odd log2) Stockham FFT in shared memory (twiddles staged
// once per CTA, the window read through L1), unpacked to the N/2+1 one-sided bins as magnitudes, contracted with the
// mel filterbank in its banded form (each triangular filter touches a contiguous run of bins; one thread per filter)
// and written as full 32-byte sectors of the [rows, n_mels, n_frames] output.
// Nothing of length n_frames * n_fft ever reaches HBM.
//
// Backward (d loss / d waveform, what `loss_mel.backward()` asks of the chain at BigVGAN/train_binaural_mel.py:759-787):
// kernel 1 recomputes the frame's spectrum (nothing but the waveform is saved), pushes the output gradient through
// log / clamp, the transposed mel contraction and the magnitude, turns the one-sided spectral gradient into the packed
// half-size transform of the real adjoint, runs the same FFT (conjugated in and out) and writes the windowed frame
// gradient to a caller-provided workspace; kernel 2 gathers, per waveform sample and in a fixed order, the frames (and
// the reflected positions) that touched it -- no atomics, so the gradient is bitwise reproducible.
#include <cuda_runtime.h>
#include <stdint.h>

#include "afa_b200.h"
#include "afa_internal.h"

namespace afa_mel {

// Frames per CTA (1 << fshift): at least 8 = one 32-byte sector of every output row when the launch is large; fewer when
// it would otherwise leave SMs idle (a training batch is ~1000 frames: latency-bound, not bandwidth-bound).  See pick_fshift().

struct MelArgs {
    const float* wav;          // [rows][row_pitch]
    float* out;                // [rows][n_mels][n_frames]
    const float* window;       // [N]
    const float2* twiddle;     // [N/2]: exp(-2 pi i t / N)
    const int32_t* band_start; // [n_mels] first bin of the filter's support
    const int32_t* band_len;   // [n_mels] bins in the support
    const int32_t* band_off;   // [n_mels] offset of the filter's weights in band_w
    const float* band_w;       // packed non-zero runs of the mel basis rows
    int64_t rows;
    int64_t T;
    int64_t row_pitch;
    int64_t n_frames;
    int n_mels;
    int hop;
    int pad;
    int pad_mode;              // AFA_MEL_PAD_REFLECT / AFA_MEL_PAD_ZERO
    float mag_eps;
    float clamp_eps;
    float log_scale;
    int raw;                   // 1: write the mel magnitudes, no clamp / log
    int fshift;                // log2(frames per CTA)
    // backward only
    const float* gout;         // [rows][n_mels][n_frames]
    float* frames;             // workspace [rows][n_frames][N]: windowed frame gradients
    const int32_t* bin_mlo;    // [N/2 + 1] the filters covering bin k lie in [bin_mlo[k], bin_mhi[k])
    const int32_t* bin_mhi;
    const float* gother;       // AFA_MEL_FLAG_L1_SIGN: gout / gother are the two log-mel tensors of an L1 loss and the
    const float* gscale_dev;   //   output gradient is sign(gout - gother) * gcoef * (*gscale_dev)  (null pointer: 1)
    float gcoef;
    int l1_sign;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// gather + window + pack one frame: z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1]  (borders by index arithmetic)
template <int M, int NT>
__device__ __forceinline__ void gather_frame(const MelArgs& p, const float* x, int64_t s0, float2* z, int tid) {
    const float2* win2 = reinterpret_cast<const float2*>(p.window);       // read-only, L1-resident, coalesced: no staging
    for (int n = tid; n < M; n += NT) {
        float v[2];
        const float2 wn = __ldg(win2 + n);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int64_t idx = s0 + 2 * n + h;
            float s;
            if (idx >= 0 && idx < p.T) {
                s = __ldg(x + idx);
            } else if (p.pad_mode == AFA_MEL_PAD_REFLECT) {
                idx = idx < 0 ? -idx : 2 * (p.T - 1) - idx;
                s = __ldg(x + idx);
            } else {
                s = 0.f;
            }
            v[h] = s * (h ? wn.y : wn.x);
        }
        z[n] = make_float2(v[0], v[1]);
    }
}

// M-point complex FFT (M = 2^(LOG2N-1)), Stockham autosort in shared memory, ping-pong between the two buffers:
// radix-4 stages (half the barriers and shared-memory round trips of radix-2) and one closing radix-2 stage when
// log2(M) is odd.  The caller has synchronised after filling buf0; returns the buffer holding the transform (synchronised).
template <int LOG2N, int NT>
__device__ __forceinline__ float2* fft_stockham(float2* buf0, float2* buf1, const float2* tw, int tid) {
    constexpr int LOG2M = LOG2N - 1;
    constexpr int M = 1 << LOG2M;
    float2* src = buf0;
    float2* dst = buf1;
#pragma unroll
    for (int ls = 0; ls + 2 <= LOG2M; ls += 2) {
        const int Ns = 1 << ls;
        for (int j = tid; j < M / 4; j += NT) {
            const int k = j & (Ns - 1);
            const int t1 = k << (LOG2M - 1 - ls);              // exp(-2 pi i k r / (4 Ns)): r = 1, 2 from the table, r = 3 as their product
            const float2 w1 = tw[t1];
            const float2 w2 = tw[2 * t1];
            const float2 w3 = cmul(w1, w2);
            const float2 v0 = src[j];
            const float2 v1 = cmul(src[j + M / 4], w1);
            const float2 v2 = cmul(src[j + M / 2], w2);
            const float2 v3 = cmul(src[j + 3 * (M / 4)], w3);
            const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
            const int j0 = ((j - k) << 2) + k;
            dst[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
            dst[j0 + Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);          // d02 - i d13
            dst[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
            dst[j0 + 3 * Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);      // d02 + i d13
        }
        __syncthreads();
        float2* t = src;
        src = dst;
        dst = t;
    }
    if (LOG2M & 1) {
        constexpr int Ns = M / 2;
        for (int j = tid; j < M / 2; j += NT) {
            const int k = j & (Ns - 1);
            const float2 a = src[j];
            const float2 b = cmul(src[j + M / 2], tw[2 * k]);  // exp(-2 pi i k / M)
            const int j0 = ((j - k) << 1) + k;
            dst[j0] = make_float2(a.x + b.x, a.y + b.y);
            dst[j0 + Ns] = make_float2(a.x - b.x, a.y - b.y);
        }
        __syncthreads();
        return dst;
    }
    return src;
}

// bin k (0 <= k <= M) of the real frame's one-sided spectrum from the packed transform Z
template <int M>
__device__ __forceinline__ float2 unpack_bin(const float2* Z, const float2* tw, int k) {
    if (k == 0 || k == M) {
        const float2 z0 = Z[0];
        return make_float2((k == 0) ? z0.x + z0.y : z0.x - z0.y, 0.f);
    }
    const float2 zk = Z[k];
    const float2 zm = Z[M - k];
    const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    const float2 o = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
    const float2 t = cmul(tw[k], o);
    return make_float2(e.x + t.y, e.y - t.x);
}

// A frame keeps M/4 threads busy (one radix-4 butterfly each), so a CTA of NT threads carries G = NT / (M/4) frames at once:
// 32 frames for the 32-point window, one for 1024 points and more.  All slots run in lockstep between CTA-wide barriers.
template <int LOG2N, int NT>
struct Slots {
    static constexpr int M = 1 << (LOG2N - 1);
    static constexpr int TPF = (M / 4 < NT) ? M / 4 : NT;
    static constexpr int G = NT / TPF;
    static constexpr int MS = G > 1 ? M + 1 : M;     // slot stride of the FFT buffers (float2): odd, so slots land on different banks
};
constexpr int kIdleLane = 1 << 30;                   // lane index of an idle slot: past every loop bound

// mel[m] = sum_q w[q] * val(b0 + q) over the filter's run, one thread per filter (a warp per filter with a shuffle
// tree was measured: 3.3x the instructions for runs of 2 ... 40 bins, 142 -> 245 us at clip length).
template <typename F>
__device__ __forceinline__ float band_sum(const MelArgs& p, int m, F val) {
    const int b0 = __ldg(p.band_start + m);
    const int bl = __ldg(p.band_len + m);
    const float* w = p.band_w + __ldg(p.band_off + m);
    float s = 0.f;
    for (int q = 0; q < bl; ++q) s = fmaf(__ldg(w + q), val(b0 + q), s);
    return s;
}

template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT) afa_logmel_kernel(const MelArgs p) {
    constexpr int N = 1 << LOG2N;
    constexpr int M = N / 2;
    constexpr int TPF = Slots<LOG2N, NT>::TPF;       // threads per frame = radix-4 butterflies per stage
    constexpr int G = Slots<LOG2N, NT>::G;           // frames in flight per CTA (small windows: up to 32)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int g = tid / TPF;
    const int l = tid % TPF;
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    constexpr int MS = Slots<LOG2N, NT>::MS;
    float2* buf0 = tw + M + g * MS;                  // [G][MS]
    float2* buf1 = buf0 + G * MS;                    // [G][MS]
    float* mag = reinterpret_cast<float*>(tw + M + 2 * G * MS) + g * (M + 1);      // [G][M + 1]
    float* melout = reinterpret_cast<float*>(tw + M + 2 * G * MS) + G * (M + 1);  // [n_mels][frames per CTA]

    const int fshift = p.fshift;
    const int fpc = 1 << fshift;
    const int64_t f0 = (int64_t)blockIdx.x << fshift;

    for (int i = tid; i < M; i += NT) tw[i] = p.twiddle[i];          // visible after the first frame's barrier

    const int nf = (int)min((int64_t)fpc, p.n_frames - f0);
    for (int64_t row = blockIdx.y; row < p.rows; row += gridDim.y) {   // gridDim.y == rows unless rows > 65535
        const float* x = p.wav + row * p.row_pitch;
        for (int fb = 0; fb < nf; fb += G) {
            const int fi = fb + g;                                     // this slot's frame; idle slots only keep the barriers
            const int le = fi < nf ? l : kIdleLane;
            const int64_t s0 = (f0 + fi) * (int64_t)p.hop - p.pad;
            gather_frame<M, TPF>(p, x, s0, buf0, le);
            __syncthreads();
            const float2* Z = fft_stockham<LOG2N, TPF>(buf0, buf1, tw, le);

            // one-sided spectrum of the real frame from the packed transform, as magnitudes
            for (int k = le; k <= M; k += TPF) {
                const float2 c = unpack_bin<M>(Z, tw, k);
                mag[k] = sqrtf(c.x * c.x + c.y * c.y + p.mag_eps);
            }
            __syncthreads();

            // banded mel contraction + dynamic-range compression
            for (int m = le; m < p.n_mels; m += TPF) {
                float s = band_sum(p, m, [&](int k) { return mag[k]; });
                if (!p.raw) s = logf(fmaxf(s, p.clamp_eps)) * p.log_scale;
                melout[(m << fshift) + fi] = s;
            }
            __syncthreads();   // mag / buf0 are rewritten by the next pass
        }

        float* o = p.out + row * (int64_t)p.n_mels * p.n_frames;
        for (int i = tid; i < (p.n_mels << fshift); i += NT) {
            const int m = i >> fshift;
            const int fi = i & (fpc - 1);
            if (fi < nf) o[(int64_t)m * p.n_frames + f0 + fi] = melout[i];
        }
        __syncthreads();   // melout is rewritten by the next row
    }
}

// Frames per CTA: a power of two, at most max(8, frames in flight per CTA), traded against the number of CTAs.
int pick_fshift(int64_t n_frames, int64_t rows, int slots) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) sms = n;
        else sms = 148;
    }
    int fs = 3;
    while ((1 << fs) < slots) ++fs;
    // more frames than slots only deepen the serial loop: give them up until every SM has 8 CTAs; fewer frames than slots
    // idle lanes: give those up only when the launch would not even put 2 CTAs on every SM
    while (fs > 0 && ((n_frames + (1 << fs) - 1) >> fs) * rows < ((1 << fs) > slots ? 8LL : 2LL) * sms) --fs;
    return fs;
}

template <int LOG2N>
int launch(MelArgs a, int64_t rows, cudaStream_t stream) {
    constexpr int N = 1 << LOG2N;
    constexpr int M = N / 2;
    constexpr int NT = (M / 4 < 128) ? 128 : (M / 4 > 256 ? 256 : M / 4);  // one radix-4 butterfly per thread and stage
    constexpr int G = Slots<LOG2N, NT>::G;
    a.fshift = pick_fshift(a.n_frames, rows, G);
    constexpr int MS = Slots<LOG2N, NT>::MS;
    const size_t smem = sizeof(float2) * (M + 2 * G * MS) + sizeof(float) * G * (M + 1) + sizeof(float) * ((size_t)a.n_mels << a.fshift);
    const int64_t gx = (a.n_frames + (1 << a.fshift) - 1) >> a.fshift;
    if (gx > 0x7fffffffLL) return afa_internal::set_error(AFA_ERR_TOO_LARGE, "afa_logmel_fwd: %lld frames per row exceed the grid", (long long)a.n_frames);
    const unsigned gy = (unsigned)(rows < 65535 ? rows : 65535);
    auto kern = afa_logmel_kernel<LOG2N, NT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return afa_internal::cuda_error(e, "cudaFuncSetAttribute(afa_logmel_kernel)");
    }
    kern<<<dim3((unsigned)gx, gy), NT, smem, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return afa_internal::cuda_error(e, "afa_logmel_kernel launch");
    afa_internal::count_launch();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// backward, kernel 1: output gradient -> windowed frame gradients (same frame slots as the forward)
// ---------------------------------------------------------------------------------------------------------------
template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT) afa_logmel_bwd_frames_kernel(const MelArgs p) {
    constexpr int N = 1 << LOG2N;
    constexpr int M = N / 2;
    constexpr int TPF = Slots<LOG2N, NT>::TPF;
    constexpr int G = Slots<LOG2N, NT>::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int g = tid / TPF;
    const int l = tid % TPF;
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    constexpr int MS = Slots<LOG2N, NT>::MS;
    float2* buf0 = tw + M + g * MS;                          // [G][MS]
    float2* buf1 = buf0 + G * MS;                            // [G][MS]
    float2* spec = tw + M + 2 * G * MS + g * (M + 1);        // [G][M + 1] (re, im), later the Hermitian half H
    float* fbase = reinterpret_cast<float*>(tw + M + 2 * G * MS + G * (M + 1));
    float* gmel = fbase + g * p.n_mels;                      // [G][n_mels]
    float* gtile = fbase + G * p.n_mels;                     // [n_mels][frames per CTA]

    const int fshift = p.fshift;
    const int fpc = 1 << fshift;
    const int64_t f0 = (int64_t)blockIdx.x << fshift;
    for (int i = tid; i < M; i += NT) tw[i] = p.twiddle[i];
    const float2* win2 = reinterpret_cast<const float2*>(p.window);
    const int nf = (int)min((int64_t)fpc, p.n_frames - f0);

    for (int64_t row = blockIdx.y; row < p.rows; row += gridDim.y) {
        const float* x = p.wav + row * p.row_pitch;
        const float* go = p.gout + row * (int64_t)p.n_mels * p.n_frames;
        const float* g2 = p.l1_sign ? p.gother + row * (int64_t)p.n_mels * p.n_frames : nullptr;
        const float l1c = p.l1_sign ? p.gcoef * (p.gscale_dev ? __ldg(p.gscale_dev) : 1.f) : 0.f;
        __syncthreads();                                     // tw staged; gtile free again
        for (int i = tid; i < (p.n_mels << fshift); i += NT) {
            const int m = i >> fshift;
            const int fi = i & (fpc - 1);
            float v = 0.f;
            if (fi < nf) {
                const int64_t at = (int64_t)m * p.n_frames + f0 + fi;
                v = __ldg(go + at);
                if (p.l1_sign) {                             // d |a - b| / d a = sign(a - b), 0 at 0 (torch.sign)
                    const float d = v - __ldg(g2 + at);
                    v = d > 0.f ? l1c : (d < 0.f ? -l1c : 0.f);
                }
            }
            gtile[i] = v;
        }
        for (int fb = 0; fb < nf; fb += G) {
            const int fi = fb + g;
            const int le = fi < nf ? l : kIdleLane;          // idle slots only keep the barriers
            const int64_t s0 = (f0 + fi) * (int64_t)p.hop - p.pad;
            gather_frame<M, TPF>(p, x, s0, buf0, le);
            __syncthreads();
            const float2* Z = fft_stockham<LOG2N, TPF>(buf0, buf1, tw, le);
            for (int k = le; k <= M; k += TPF) spec[k] = unpack_bin<M>(Z, tw, k);
            __syncthreads();

            // d loss / d mel[m]: the forward's mel value is recomputed, then log / clamp are differentiated
            for (int m = le; m < p.n_mels; m += TPF) {
                float gm = gtile[(m << fshift) + fi];
                if (!p.raw) {
                    const float s = band_sum(p, m, [&](int k) {
                        const float2 c = spec[k];
                        return sqrtf(c.x * c.x + c.y * c.y + p.mag_eps);
                    });
                    gm = s >= p.clamp_eps ? gm * p.log_scale / s : 0.f;      // torch.clamp passes the gradient where x >= min
                }
                gmel[m] = gm;
            }
            __syncthreads();

            // transposed mel contraction, magnitude, and the Hermitian half of the real adjoint:
            // g[n] = Re sum_{k=0..M} G_k e^{+2 pi i k n / N} = sum_{k=0..N-1} H_k e^{...},  H_0 = Re G_0, H_M = Re G_M, H_k = G_k / 2
            for (int k = le; k <= M; k += TPF) {
                const float2 c = spec[k];
                const float mag = sqrtf(c.x * c.x + c.y * c.y + p.mag_eps);
                float gk = 0.f;
                const int mhi = __ldg(p.bin_mhi + k);
                for (int m = __ldg(p.bin_mlo + k); m < mhi; ++m) {
                    const int q = k - __ldg(p.band_start + m);
                    if (q >= 0 && q < __ldg(p.band_len + m)) gk = fmaf(__ldg(p.band_w + __ldg(p.band_off + m) + q), gmel[m], gk);
                }
                const float sc = mag > 0.f ? gk / mag : 0.f;                 // |.| has gradient 0 at 0 (torch.abs), as sqrt(. + eps) never gets there
                const float half = (k == 0 || k == M) ? 1.f : 0.5f;
                spec[k] = make_float2(half * sc * c.x, (k == 0 || k == M) ? 0.f : half * sc * c.y);
            }
            __syncthreads();

            // packed half-size spectrum of the real sequence: Z_k = (H_k + conj H_{M-k}) + i (H_k - conj H_{M-k}) e^{+2 pi i k / N};
            // the inverse transform is taken as conj(FFT(conj Z))
            for (int k = le; k < M; k += TPF) {
                const float2 hk = spec[k];
                const float2 hm = spec[M - k];
                const float2 e = make_float2(hk.x + hm.x, hk.y - hm.y);
                const float2 d = make_float2(hk.x - hm.x, hk.y + hm.y);
                const float2 t = tw[k];
                const float2 o = cmul(d, make_float2(t.x, -t.y));
                buf0[k] = make_float2(e.x - o.y, -(e.y + o.x));
            }
            __syncthreads();
            const float2* z = fft_stockham<LOG2N, TPF>(buf0, buf1, tw, le);
            float2* fr = reinterpret_cast<float2*>(p.frames + ((row * p.n_frames + f0 + fi) << LOG2N));
            for (int n = le; n < M; n += TPF) {
                const float2 v = z[n];
                const float2 wn = __ldg(win2 + n);
                fr[n] = make_float2(v.x * wn.x, -v.y * wn.y);
            }
            __syncthreads();   // buf0 / buf1 / spec / gmel are rewritten by the next frame
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, kernel 2: overlap-add as a gather.  Sample t of the waveform was read at padded position t + pad and,
// under reflect padding, at pad - t (1 <= t <= pad) and pad + 2 (T - 1) - t (1 <= T - 1 - t <= pad).
// ---------------------------------------------------------------------------------------------------------------
struct OlaArgs {
    const float* frames;
    float* gwav;
    int64_t rows, T, gwav_pitch, n_frames;
    int n_fft, hop, pad, pad_mode;
    int accumulate;            // AFA_MEL_FLAG_ACCUMULATE: gwav += instead of gwav =
};

__device__ __forceinline__ float ola_position(const OlaArgs& p, const float* fr, int64_t pos) {
    int64_t f_hi = pos / p.hop;
    if (f_hi > p.n_frames - 1) f_hi = p.n_frames - 1;
    int64_t f_lo = pos - p.n_fft + 1;
    f_lo = f_lo <= 0 ? 0 : (f_lo + p.hop - 1) / p.hop;
    float s = 0.f;
    for (int64_t f = f_lo; f <= f_hi; ++f) s += __ldg(fr + f * p.n_fft + (pos - f * p.hop));
    return s;
}

__global__ void __launch_bounds__(256) afa_logmel_bwd_ola_kernel(const OlaArgs p) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= p.T) return;
    for (int64_t row = blockIdx.y; row < p.rows; row += gridDim.y) {
        const float* fr = p.frames + row * p.n_frames * p.n_fft;
        float s = ola_position(p, fr, t + p.pad);
        if (p.pad_mode == AFA_MEL_PAD_REFLECT) {
            if (t >= 1 && t <= p.pad) s += ola_position(p, fr, p.pad - t);
            const int64_t r = p.T - 1 - t;
            if (r >= 1 && r <= p.pad) s += ola_position(p, fr, p.pad + 2 * (p.T - 1) - t);
        }
        float* o = p.gwav + row * p.gwav_pitch + t;
        *o = p.accumulate ? *o + s : s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sum |a - b| as n_partial per-CTA partial sums in a fixed order (the caller finishes the few hundred values):
// the L1 of two log-mel tensors without materialising a - b, |.| and the reduction tree as separate launches
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) afa_l1_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                              float* __restrict__ partial) {
    __shared__ float warp_sums[8];
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s += fabsf(__ldg(a + i) - __ldg(b + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += warp_sums[w];
        partial[blockIdx.x] = t;
    }
}

template <int LOG2N>
int launch_bwd(MelArgs a, const OlaArgs& o, int64_t rows, cudaStream_t stream) {
    constexpr int N = 1 << LOG2N;
    constexpr int M = N / 2;
    constexpr int NT = (M / 4 < 128) ? 128 : (M / 4 > 256 ? 256 : M / 4);
    constexpr int G = Slots<LOG2N, NT>::G;
    a.fshift = pick_fshift(a.n_frames, rows, G);
    constexpr int MS = Slots<LOG2N, NT>::MS;
    const size_t smem = sizeof(float2) * (M + 2 * G * MS + G * (M + 1)) + sizeof(float) * (((size_t)a.n_mels << a.fshift) + (size_t)G * a.n_mels);
    const int64_t gx = (a.n_frames + (1 << a.fshift) - 1) >> a.fshift;
    const int64_t ox = (a.T + 255) / 256;
    if (gx > 0x7fffffffLL || ox > 0x7fffffffLL) return afa_internal::set_error(AFA_ERR_TOO_LARGE, "afa_logmel_bwd: row too long for the grid");
    const unsigned gy = (unsigned)(rows < 65535 ? rows : 65535);
    auto kern = afa_logmel_bwd_frames_kernel<LOG2N, NT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return afa_internal::cuda_error(e, "cudaFuncSetAttribute(afa_logmel_bwd_frames_kernel)");
    }
    kern<<<dim3((unsigned)gx, gy), NT, smem, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return afa_internal::cuda_error(e, "afa_logmel_bwd_frames_kernel launch");
    afa_internal::count_launch();
    afa_logmel_bwd_ola_kernel<<<dim3((unsigned)ox, gy), 256, 0, stream>>>(o);
    e = cudaGetLastError();
    if (e != cudaSuccess) return afa_internal::cuda_error(e, "afa_logmel_bwd_ola_kernel launch");
    afa_internal::count_launch();
    return 0;
}

}  // namespace afa_mel

extern "C" int afa_l1_partial_sums(const float* a, const float* b, int64_t n, float* partial, int n_partial, void* stream) {
    if (n < 0 || n_partial <= 0 || n_partial > 65535)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_l1_partial_sums: bad sizes (n %lld, n_partial %d)", (long long)n, n_partial);
    if (!partial || (n > 0 && (!a || !b))) return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_l1_partial_sums: null pointer");
    afa_mel::afa_l1_partial_kernel<<<n_partial, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, n, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return afa_internal::cuda_error(e, "afa_l1_partial_kernel launch");
    afa_internal::count_launch();
    return 0;
}

extern "C" int64_t afa_logmel_num_frames(int64_t T, int n_fft, int hop, int pad) {
    if (T <= 0 || n_fft <= 0 || hop <= 0 || pad < 0) return 0;
    const int64_t padded = T + 2 * (int64_t)pad;
    return padded < n_fft ? 0 : 1 + (padded - n_fft) / hop;
}

extern "C" int afa_logmel_fwd(const float* wav, float* out, int64_t rows, int64_t T, int64_t row_pitch,
                              int n_fft, int hop, int pad, int pad_mode,
                              const float* window, const float* twiddle,
                              int n_mels, const int32_t* band_start, const int32_t* band_len, const int32_t* band_off,
                              const float* band_w,
                              float mag_eps, float clamp_eps, float log_scale, int flags, void* stream) {
    using namespace afa_mel;
    if (rows < 0 || T <= 0 || row_pitch < T || hop <= 0 || pad < 0 || n_mels <= 0 || n_mels > 1024)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_fwd: bad shape (rows %lld, T %lld, pitch %lld, hop %d, pad %d, n_mels %d)",
                                       (long long)rows, (long long)T, (long long)row_pitch, hop, pad, n_mels);
    if (pad_mode != AFA_MEL_PAD_REFLECT && pad_mode != AFA_MEL_PAD_ZERO)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_fwd: pad_mode %d", pad_mode);
    if (pad_mode == AFA_MEL_PAD_REFLECT && pad >= T)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_fwd: reflect padding %d needs T > pad (T = %lld), as torch's F.pad does", pad, (long long)T);
    int log2n = 0;
    while ((1 << log2n) < n_fft) ++log2n;
    if ((1 << log2n) != n_fft || log2n < 5 || log2n > 11)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_fwd: n_fft %d is not a power of two in [32, 2048]", n_fft);
    MelArgs a;
    a.wav = wav; a.out = out; a.window = window; a.twiddle = reinterpret_cast<const float2*>(twiddle);
    a.band_start = band_start; a.band_len = band_len; a.band_off = band_off; a.band_w = band_w;
    a.T = T; a.row_pitch = row_pitch;
    a.n_frames = afa_logmel_num_frames(T, n_fft, hop, pad);
    a.n_mels = n_mels; a.hop = hop; a.pad = pad; a.pad_mode = pad_mode;
    a.mag_eps = mag_eps; a.clamp_eps = clamp_eps; a.log_scale = log_scale; a.raw = (flags & AFA_MEL_FLAG_RAW) ? 1 : 0;
    a.rows = rows;
    a.gout = nullptr; a.frames = nullptr; a.bin_mlo = nullptr; a.bin_mhi = nullptr;
    a.gother = nullptr; a.gscale_dev = nullptr; a.gcoef = 0.f; a.l1_sign = 0;
    if (rows == 0 || a.n_frames == 0) return 0;     // nothing to do: empty batches / rows shorter than one frame are legal
    if (!wav || !out || !window || !twiddle || !band_start || !band_len || !band_off || !band_w)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_fwd: null pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (log2n) {
        case 5: return launch<5>(a, rows, s);
        case 6: return launch<6>(a, rows, s);
        case 7: return launch<7>(a, rows, s);
        case 8: return launch<8>(a, rows, s);
        case 9: return launch<9>(a, rows, s);
        case 10: return launch<10>(a, rows, s);
        default: return launch<11>(a, rows, s);
    }
}

extern "C" size_t afa_logmel_bwd_workspace_bytes(int64_t rows, int64_t T, int n_fft, int hop, int pad) {
    const int64_t nf = afa_logmel_num_frames(T, n_fft, hop, pad);
    return rows <= 0 ? 0 : (size_t)rows * (size_t)nf * (size_t)n_fft * sizeof(float);
}

extern "C" int afa_logmel_bwd(const float* wav, const float* gout, float* gwav, int64_t rows, int64_t T, int64_t row_pitch,
                              int64_t gwav_pitch, int n_fft, int hop, int pad, int pad_mode,
                              const float* window, const float* twiddle,
                              int n_mels, const int32_t* band_start, const int32_t* band_len, const int32_t* band_off,
                              const float* band_w, const int32_t* bin_mlo, const int32_t* bin_mhi,
                              float mag_eps, float clamp_eps, float log_scale, int flags,
                              const float* gother, float gcoef, const float* gscale_dev,
                              void* workspace, size_t workspace_bytes, void* stream) {
    using namespace afa_mel;
    if (rows < 0 || T <= 0 || row_pitch < T || gwav_pitch < T || hop <= 0 || pad < 0 || n_mels <= 0 || n_mels > 1024)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_bwd: bad shape (rows %lld, T %lld, pitches %lld / %lld, hop %d, pad %d, n_mels %d)",
                                       (long long)rows, (long long)T, (long long)row_pitch, (long long)gwav_pitch, hop, pad, n_mels);
    if (pad_mode != AFA_MEL_PAD_REFLECT && pad_mode != AFA_MEL_PAD_ZERO)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_bwd: pad_mode %d", pad_mode);
    if (pad_mode == AFA_MEL_PAD_REFLECT && pad >= T)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_bwd: reflect padding %d needs T > pad (T = %lld)", pad, (long long)T);
    int log2n = 0;
    while ((1 << log2n) < n_fft) ++log2n;
    if ((1 << log2n) != n_fft || log2n < 5 || log2n > 11)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_bwd: n_fft %d is not a power of two in [32, 2048]", n_fft);
    if (rows == 0) return 0;
    if (!wav || !gwav) return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_bwd: null pointer");
    MelArgs a;
    a.wav = wav; a.out = nullptr; a.window = window; a.twiddle = reinterpret_cast<const float2*>(twiddle);
    a.band_start = band_start; a.band_len = band_len; a.band_off = band_off; a.band_w = band_w;
    a.rows = rows; a.T = T; a.row_pitch = row_pitch;
    a.n_frames = afa_logmel_num_frames(T, n_fft, hop, pad);
    a.n_mels = n_mels; a.hop = hop; a.pad = pad; a.pad_mode = pad_mode;
    a.mag_eps = mag_eps; a.clamp_eps = clamp_eps; a.log_scale = log_scale; a.raw = (flags & AFA_MEL_FLAG_RAW) ? 1 : 0;
    a.gout = gout; a.frames = reinterpret_cast<float*>(workspace); a.bin_mlo = bin_mlo; a.bin_mhi = bin_mhi;
    a.l1_sign = (flags & AFA_MEL_FLAG_L1_SIGN) ? 1 : 0;
    a.gother = gother; a.gcoef = gcoef; a.gscale_dev = gscale_dev;
    const int accumulate = (flags & AFA_MEL_FLAG_ACCUMULATE) ? 1 : 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (a.n_frames == 0) {                          // no frame read the waveform: the gradient is zero
        if (accumulate) return 0;
        cudaError_t e = cudaMemset2DAsync(gwav, (size_t)gwav_pitch * sizeof(float), 0, (size_t)T * sizeof(float), (size_t)rows, s);
        return e == cudaSuccess ? 0 : afa_internal::cuda_error(e, "cudaMemset2DAsync(gwav)");
    }
    if (!gout || !window || !twiddle || !band_start || !band_len || !band_off || !band_w || !bin_mlo || !bin_mhi || !workspace)
        return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_bwd: null pointer");
    if (a.l1_sign && !gother) return afa_internal::set_error(AFA_ERR_BAD_ARG, "afa_logmel_bwd: AFA_MEL_FLAG_L1_SIGN needs the second log-mel tensor");
    const size_t need = afa_logmel_bwd_workspace_bytes(rows, T, n_fft, hop, pad);
    if (workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 7))
        return afa_internal::set_error(AFA_ERR_WORKSPACE, "afa_logmel_bwd: workspace of %zu bytes (8-byte aligned) needed, %zu given", need, workspace_bytes);
    OlaArgs o;
    o.frames = a.frames; o.gwav = gwav; o.rows = rows; o.T = T; o.gwav_pitch = gwav_pitch; o.n_frames = a.n_frames;
    o.n_fft = n_fft; o.hop = hop; o.pad = pad; o.pad_mode = pad_mode; o.accumulate = accumulate;
    switch (log2n) {
        case 5: return launch_bwd<5>(a, o, rows, s);
        case 6: return launch_bwd<6>(a, o, rows, s);
        case 7: return launch_bwd<7>(a, o, rows, s);
        case 8: return launch_bwd<8>(a, o, rows, s);
        case 9: return launch_bwd<9>(a, o, rows, s);
        case 10: return launch_bwd<10>(a, o, rows, s);
        default: return launch_bwd<11>(a, o, rows, s);
    }
}
