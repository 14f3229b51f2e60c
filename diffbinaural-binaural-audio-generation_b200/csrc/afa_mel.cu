// afa_mel.cu -- fused log-mel spectrogram for sm_100a (forward).
//
// One launch replaces the torch-op chain of the reference's `mel_spectrogram`
// (BigVGAN/meldataset.py:51-123: reflect pad -> torch.stft(hann) -> sqrt(re^2 + im^2 + 1e-9) -> mel_basis @ spec
// -> log(clamp(., 1e-5))) and of `MultiScaleMelSpectrogramLoss.mel_spectrogram` + its log10
// (BigVGAN/loss.py:110-167, 195-200), which differ only in pad width, magnitude epsilon, clamp and log base.
//
// Per CTA: kFrames consecutive STFT frames of one waveform row.  Each frame is windowed while it is gathered from
// global memory (replicated / reflected borders resolved by index arithmetic, no padded copy), packed as N/2 complex
// points, transformed by a radix-2 Stockham FFT in shared memory (twiddles and window staged once per CTA), unpacked
// to the N/2+1 one-sided bins as magnitudes, contracted with the mel filterbank in its banded form (each triangular
// filter touches a contiguous run of bins) and written as full 32-byte sectors of the [rows, n_mels, n_frames] output.
// Nothing of length n_frames * n_fft ever reaches HBM.
#include <cuda_runtime.h>
#include <stdint.h>

#include "afa_b200.h"
#include "afa_internal.h"

namespace afa_mel {

constexpr int kFrames = 8;   // frames per CTA = one 32-byte sector of every output row

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
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

template <int LOG2N, int NT>
__global__ void __launch_bounds__(NT) afa_logmel_kernel(const MelArgs p) {
    constexpr int N = 1 << LOG2N;
    constexpr int M = N / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);
    float2* buf1 = buf0 + M;
    float2* tw = buf1 + M;
    float* win = reinterpret_cast<float*>(tw + M);
    float* mag = win + N;                 // [M + 1]
    float* melout = mag + (M + 1);        // [n_mels][kFrames]

    const int tid = threadIdx.x;
    const int64_t f0 = (int64_t)blockIdx.x * kFrames;

    for (int i = tid; i < M; i += NT) tw[i] = p.twiddle[i];
    for (int i = tid; i < N; i += NT) win[i] = p.window[i];
    __syncthreads();

    const int nf = (int)min((int64_t)kFrames, p.n_frames - f0);
    for (int64_t row = blockIdx.y; row < p.rows; row += gridDim.y) {   // gridDim.y == rows unless rows > 65535
        const float* x = p.wav + row * p.row_pitch;
        for (int fi = 0; fi < nf; ++fi) {
            const int64_t s0 = (f0 + fi) * (int64_t)p.hop - p.pad;
            // gather + window + pack: z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1]
            for (int n = tid; n < M; n += NT) {
                float v[2];
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
                    v[h] = s * win[2 * n + h];
                }
                buf0[n] = make_float2(v[0], v[1]);
            }
            __syncthreads();

            // M-point complex FFT, radix-2 Stockham autosort, ping-pong between buf0 and buf1
            float2* src = buf0;
            float2* dst = buf1;
#pragma unroll
            for (int ls = 0; ls < LOG2N - 1; ++ls) {
                const int Ns = 1 << ls;
                for (int j = tid; j < M / 2; j += NT) {
                    const int k = j & (Ns - 1);
                    const float2 w = tw[k << (LOG2N - 1 - ls)];      // exp(-2 pi i k / (2 Ns))
                    const float2 a = src[j];
                    const float2 b = cmul(src[j + M / 2], w);
                    const int j0 = ((j - k) << 1) + k;
                    dst[j0] = make_float2(a.x + b.x, a.y + b.y);
                    dst[j0 + Ns] = make_float2(a.x - b.x, a.y - b.y);
                }
                __syncthreads();
                float2* t = src;
                src = dst;
                dst = t;
            }
            const float2* Z = src;

            // one-sided spectrum of the real frame from the packed transform, as magnitudes
            for (int k = tid; k <= M; k += NT) {
                float re, im;
                if (k == 0 || k == M) {
                    const float2 z0 = Z[0];
                    re = (k == 0) ? z0.x + z0.y : z0.x - z0.y;
                    im = 0.f;
                } else {
                    const float2 zk = Z[k];
                    const float2 zm = Z[M - k];
                    const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                    const float2 o = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
                    const float2 t = cmul(tw[k], o);
                    re = e.x + t.y;
                    im = e.y - t.x;
                }
                mag[k] = sqrtf(re * re + im * im + p.mag_eps);
            }
            __syncthreads();

            // banded mel contraction + dynamic-range compression
            for (int m = tid; m < p.n_mels; m += NT) {
                const int b0 = __ldg(p.band_start + m);
                const int bl = __ldg(p.band_len + m);
                const float* w = p.band_w + __ldg(p.band_off + m);
                float s = 0.f;
                for (int q = 0; q < bl; ++q) s = fmaf(__ldg(w + q), mag[b0 + q], s);
                if (!p.raw) s = logf(fmaxf(s, p.clamp_eps)) * p.log_scale;
                melout[m * kFrames + fi] = s;
            }
            __syncthreads();   // mag / buf0 are rewritten by the next frame
        }

        float* o = p.out + row * (int64_t)p.n_mels * p.n_frames;
        for (int i = tid; i < p.n_mels * kFrames; i += NT) {
            const int m = i / kFrames;
            const int fi = i - m * kFrames;
            if (fi < nf) o[(int64_t)m * p.n_frames + f0 + fi] = melout[i];
        }
        __syncthreads();   // melout is rewritten by the next row
    }
}

template <int LOG2N>
int launch(const MelArgs& a, int64_t rows, cudaStream_t stream) {
    constexpr int N = 1 << LOG2N;
    constexpr int M = N / 2;
    constexpr int NT = (M / 2 < 32) ? 32 : (M / 2 > 256 ? 256 : M / 2);
    const size_t smem = sizeof(float2) * 3 * M + sizeof(float) * (N + M + 1) + sizeof(float) * (size_t)a.n_mels * kFrames;
    const int64_t gx = (a.n_frames + kFrames - 1) / kFrames;
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

}  // namespace afa_mel

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
