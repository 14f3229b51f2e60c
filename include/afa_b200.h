/*
 * afa_b200.h -- C ABI of libafa_sm100.so: the fused anti-aliased activation (Activation1d) for
 * NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the one hot path this repository accelerates.  Each entry point
 * replaces the torch-op sequence behind a reference interface (paths relative to the reference tree):
 *
 *   afa_activation1d_fwd   <->  Activation1d.forward            BigVGAN/alias_free_activation/act.py:25-30
 *                               = UpSample1d.forward            BigVGAN/alias_free_activation/resample.py:29-38
 *                               + Snake/SnakeBeta.forward       BigVGAN/activations.py:51-62, 113-126
 *                               + LowPassFilter1d.forward       BigVGAN/alias_free_activation/filter.py:94-101
 *   afa_activation1d_bwd   <->  autograd backward of the above, triggered at
 *                               BigVGAN/train_binaural_mel.py:787 (loss_gen_all.backward())
 *
 * The reference reaches this boundary through `alias_free_activation.cuda.activation1d.Activation1d`
 * (BigVGAN/bigvgan.py:94-102, 194-202, 272-280); the Python module of that name in this repository
 * binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All device pointers are BORROWED: the library
 *     never allocates, frees or retains device memory.
 *   - x / y / gy / gx are dense row-major [batch, channels, T] device arrays of `dtype`
 *     (AFA_DTYPE_F32 or AFA_DTYPE_BF16; bf16 I/O computes in fp32).
 *   - alpha / beta are the RAW per-channel parameters (float32 [channels], device) exactly as stored
 *     in `act.alpha` / `act.beta`; AFA_FLAG_LOGSCALE applies exp() (activations.py:119-123);
 *     AFA_FLAG_SNAKE means beta aliases alpha (activations.py:57-60) and `beta` is ignored.
 *   - taps_up12 / taps_down12 are HOST pointers to the 12 filter taps held in the module buffers
 *     `upsample.filter` / `downsample.lowpass.filter` (resample.py:23-26, filter.py:90-91).
 *   - `stream` is a cudaStream_t (NULL = default stream).  Launches are asynchronous; nothing
 *     synchronises the device, so calls are CUDA-graph capturable.
 *   - return 0 on success; a negative AFA_ERR_* for argument errors; a positive cudaError_t for
 *     CUDA failures.  afa_last_error() returns a thread-local description.  Nothing throws or exits.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef AFA_B200_H_
#define AFA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFA_VERSION 152            /* 0.1.5: + afa_activation1d_fwd_pitched, split parameter-gradient finalize, channels-last tensor-core forward; 0.1.4: tensor-core (tcgen05) forward for bf16 activations */

#define AFA_DTYPE_F32 0
#define AFA_DTYPE_BF16 1

#define AFA_FLAG_LOGSCALE 1        /* act.alpha_logscale                        activations.py:42,98  */
#define AFA_FLAG_SNAKE 2           /* Snake (beta := alpha) instead of SnakeBeta  activations.py:51-62  */

#define AFA_ERR_BAD_ARG (-1)
#define AFA_ERR_BAD_DTYPE (-2)
#define AFA_ERR_TOO_LARGE (-3)
#define AFA_ERR_WORKSPACE (-4)
#define AFA_ERR_ALIGNMENT (-5)

int afa_version(void);
const char *afa_last_error(void);

/* y = down2x(snake(up2x(x))).  x, y: [batch, channels, T] of dtype; y may not alias x. */
int afa_activation1d_fwd(const void *x, void *y,
                         const float *alpha, const float *beta,
                         const float *taps_up12, const float *taps_down12,
                         int64_t batch, int64_t channels, int64_t T,
                         int dtype, int flags, void *stream);

/*
 * The same forward for tensors whose ROWS are `row_pitch` elements apart (x_row_pitch, y_row_pitch >= T), e.g. a time slice
 * x[:, :, :T] of a longer buffer: what the reference's torch ops accept through strides (act.py:25-30 never asks for a dense
 * tensor).  Served by the tensor-core kernel, whose tensor maps carry the pitch: bf16, T % 8 == 0, T >= 64, both pitches
 * multiples of 8 elements, x and y 16-byte aligned; anything else returns AFA_ERR_ALIGNMENT (the caller copies to a dense
 * tensor and calls afa_activation1d_fwd).  Pitches equal to T forward to afa_activation1d_fwd.
 */
int afa_activation1d_fwd_pitched(const void *x, int64_t x_row_pitch, void *y, int64_t y_row_pitch,
                                 const float *alpha, const float *beta,
                                 const float *taps_up12, const float *taps_down12,
                                 int64_t batch, int64_t channels, int64_t T,
                                 int dtype, int flags, void *stream);

/* Scratch the backward needs (per-segment parameter-gradient partials, slice sums, counters), in bytes. */
size_t afa_bwd_workspace_bytes(int64_t batch, int64_t channels, int64_t T, int dtype);

/*
 * gx = d<y,gy>/dx ; galpha / gbeta = gradients w.r.t. the RAW parameters (float32 [channels],
 * log-scale chain rule and Snake aliasing applied; gbeta may be NULL with AFA_FLAG_SNAKE).
 * The intermediate 2x-rate signal is recomputed from x; only x is needed from the forward.
 * Reductions are staged and deterministic: fixed-order sums per segment, per slice of a channel's
 * segments, per channel (an arrival counter only elects the CTA that adds the slice sums, in slice
 * order; no floating-point atomics).  The workspace need not be initialised.
 */
int afa_activation1d_bwd(const void *x, const void *gy, void *gx,
                         float *galpha, float *gbeta,
                         const float *alpha, const float *beta,
                         const float *taps_up12, const float *taps_down12,
                         int64_t batch, int64_t channels, int64_t T,
                         int dtype, int flags,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * AMP-block entry points on CHANNELS-LAST activations: dense [batch, T, channels] arrays, channels
 * contiguous, with a per-tensor batch stride in elements (>= T*channels; rows may be padded in time).
 * They are what an inference engine needs to keep the generator channels-last between cuDNN's NHWC
 * tensor-core convolutions (no layout conversion kernels) and to drop the separate bias / residual /
 * mean kernels of the reference's AMPBlock.  Forward only.
 *
 *   afa_amp_activation1d_fwd_cl  <->  `xt = c(xt)` bias add + `x = xt + x` + the next `a(x)`
 *                                      BigVGAN/bigvgan.py:132-141 (AMPBlock1.forward), :233-236 (AMPBlock2)
 *   afa_resblock_mean            <->  `xs += resblocks[...](x)` ... `x = xs / self.num_kernels`
 *                                      BigVGAN/bigvgan.py:368-376
 *   afa_tail_fwd_cl              <->  activation_post -> conv_post -> clamp | tanh   BigVGAN/bigvgan.py:379-385
 *                                      (+ `* MAX_WAV_VALUE`, astype("int16"), stereo interleave
 *                                       BigVGAN/inference_e2e.py:193-201)
 * ------------------------------------------------------------------------------------------------ */

/*
 * y = down2x(snake(up2x(x + res + bias[c]))) ;  xsum = x + res  (the new residual stream; `bias` -- the sum of
 * the convolution biases not yet applied to x and res -- stays pending with the caller, who passes it again,
 * plus the next convolution's, to the next call and finally to afa_resblock_mean).
 * bias: float32 [channels] device or NULL; res and xsum: both NULL or both arrays shaped like x.  y_tpad >= T (0 = T):
 * rows [T, y_tpad) of every batch entry of y are written as zeros (the zero padding a dilated convolution
 * run as a (k x 1) convolution over the [y_tpad/d, d] polyphase view reads).  Outputs must not alias inputs.
 */
int afa_amp_activation1d_fwd_cl(const void *x, int64_t x_bstride,
                                const void *res, int64_t res_bstride,
                                const float *bias,
                                void *xsum, int64_t xsum_bstride,
                                void *y, int64_t y_bstride, int64_t y_tpad,
                                const float *alpha, const float *beta,
                                const float *taps_up12, const float *taps_down12,
                                int64_t batch, int64_t channels, int64_t T,
                                int dtype, int flags, void *stream);

/*
 * out = scale * (sum_{j < num_kernels} (xt[j] + xres[j]) + bias_sum[c]) over dense [rows, channels] arrays
 * (rows = batch*T); an entry xres[j] may be NULL (xt[j] already contains its residual stream).  xt[j]: output of the last convolution of resblock j WITHOUT its bias; xres[j]: that
 * resblock's residual stream; bias_sum: float32 [channels] = sum of those biases (or NULL); scale = 1/num_kernels.
 * xt / xres are HOST arrays of device pointers.
 */
int afa_resblock_mean(const void *const *xt, const void *const *xres, int num_kernels,
                      const float *bias_sum, float scale, void *out,
                      int64_t rows, int64_t channels, int dtype, void *stream);

/*
 * wave[b][t] = final(conv_post(activation_post(x)))  with conv_post: channels -> 1, kernel 7, zero padding 3,
 * weight w_post float32 [channels][7] (device), optional bias_post (1 float, device), final = clamp(-1, 1)
 * (use_tanh = 0) or tanh.  channels <= 32.  Outputs (at least one): wave float32 [batch][T]; pcm int16 with
 * element (b, t) at ((b / pcm_interleave) * T_out + t) * pcm_interleave + b % pcm_interleave, value
 * (int16)(wave * pcm_scale) truncated toward zero like numpy's astype("int16") (pcm_scale = 32767).
 * Zero-frame restoration (BigVGAN/inference_e2e.py:38-111, reconstruct_audio_with_silence): with frame_map
 * (int32 [batch][T / hop], device) sample t of batch entry b is written at frame_map[b][t / hop] * hop + t % hop
 * of an output row of T_out samples (wave [batch][T_out], pcm [batch / il][T_out][il]); the caller zero-fills the
 * outputs first (the silence).  frame_map values must lie in [0, T_out / hop); a sample whose frame index is negative or whose
 * position falls at or beyond T_out is dropped (the reference clamps every copy to the original length,
 * inference_e2e.py:94-109).  frame_map = NULL: identity, T_out = T (or 0).
 */
int afa_tail_fwd_cl(const void *x, int64_t x_bstride,
                    const float *alpha, const float *beta,
                    const float *taps_up12, const float *taps_down12,
                    const float *w_post, const float *bias_post, int use_tanh,
                    float *wave, int16_t *pcm, int pcm_interleave, float pcm_scale,
                    const int32_t *frame_map, int hop, int64_t T_out,
                    int64_t batch, int64_t channels, int64_t T,
                    int dtype, int flags, void *stream);

/*
 * Activation1d as the PROLOGUE of the AMPBlock convolution, for the narrow stages of the generator:
 *   y = conv1d(down2x(snake(up2x(x + res + bias[c]))), w, no bias, 'same' padding, dilation) + addend ; xsum = x + res
 *   <->  `xt = a(x); xt = c(xt)` [+ `x = xt + x`]  BigVGAN/bigvgan.py:134-141 (AMPBlock1), :234-236 (AMPBlock2)
 * addend (optional, [batch, T, channels], 16-byte aligned): the block's residual stream, added in fp32 to the
 * convolution's accumulators before the single rounding of y -- then y IS the new residual stream and the next
 * activation needs no residual prologue.
 * Same conventions as afa_amp_activation1d_fwd_cl (the convolution's own bias stays pending with the caller).
 * w_kcc: bf16 device array [kernel_size][channels_out][channels_in] (= conv.weight.permute(2, 0, 1)), 16-byte
 * aligned; channels_out == channels_in == channels.  bf16 activations only (the convolution runs on the
 * tensor cores with fp32 accumulation).  afa_amp_act_conv_supported() says whether a configuration is compiled.
 */
int afa_amp_act_conv_supported(int64_t channels, int kernel_size, int dilation, int dtype);
int afa_amp_act_conv_fwd_cl(const void *x, int64_t x_bstride,
                            const void *res, int64_t res_bstride,
                            const float *bias,
                            void *xsum, int64_t xsum_bstride,
                            const void *addend, int64_t addend_bstride,
                            void *y, int64_t y_bstride,
                            const float *alpha, const float *beta,
                            const float *taps_up12, const float *taps_down12,
                            const void *w_kcc, int kernel_size, int dilation,
                            int64_t batch, int64_t channels, int64_t T,
                            int dtype, int flags, void *stream);

/*
 * Fused log-mel spectrogram: one launch for
 *   reflect/zero pad -> STFT (n_fft-point, hop, window) -> sqrt(re^2 + im^2 + mag_eps) -> mel_basis @ magnitudes
 *   -> log(max(., clamp_eps)) * log_scale
 *   <->  mel_spectrogram                                      BigVGAN/meldataset.py:51-123
 *        (pad = (n_fft - hop) / 2, center=False, mag_eps = 1e-9, clamp_eps = 1e-5, log_scale = 1; called at
 *         BigVGAN/train_binaural_mel.py:386, 640, 711 and, via get_mel_spectrogram, BigVGAN/inference_binaural.py:131-132)
 *   <->  MultiScaleMelSpectrogramLoss.mel_spectrogram + log10  BigVGAN/loss.py:110-167, 195-200
 *        (center=True => pad = n_fft / 2, mag_eps = 0, clamp_eps = 1e-5, log_scale = 1 / ln 10)
 * wav:  float32 device array [rows][row_pitch], T valid samples per row.
 * out:  float32 device array [rows][n_mels][n_frames], n_frames = afa_logmel_num_frames(T, n_fft, hop, pad).
 * n_fft: power of two in [32, 2048].  window: float32 device [n_fft] (a shorter window is zero-padded to n_fft
 *   and centred by the caller, as torch.stft does).  twiddle: float32 device [n_fft / 2][2] = (cos, -sin)(2 pi t / n_fft).
 * The mel basis is passed in banded form, built by the caller from the dense [n_mels][n_fft / 2 + 1] matrix the
 * reference caches (meldataset.py:89-93): row m is non-zero on bins [band_start[m], band_start[m] + band_len[m])
 * and its weights there are band_w[band_off[m] ...] (all device arrays; int32 / float32).
 * pad_mode: AFA_MEL_PAD_REFLECT (2-D input in the reference) or AFA_MEL_PAD_ZERO (its 1-D input branch, :96-97).
 * flags: AFA_MEL_FLAG_RAW writes the mel magnitudes without clamp / log.
 */
#define AFA_MEL_PAD_REFLECT 0
#define AFA_MEL_PAD_ZERO 1
#define AFA_MEL_FLAG_RAW 1
#define AFA_MEL_FLAG_L1_SIGN 2       /* afa_logmel_bwd: the output gradient is sign(gout - gother) * gcoef * (*gscale_dev) */
#define AFA_MEL_FLAG_ACCUMULATE 4    /* afa_logmel_bwd: gwav += instead of gwav = (summing the scales of the loss) */
int64_t afa_logmel_num_frames(int64_t T, int n_fft, int hop, int pad);
int afa_logmel_fwd(const float *wav, float *out, int64_t rows, int64_t T, int64_t row_pitch,
                   int n_fft, int hop, int pad, int pad_mode,
                   const float *window, const float *twiddle,
                   int n_mels, const int32_t *band_start, const int32_t *band_len, const int32_t *band_off,
                   const float *band_w,
                   float mag_eps, float clamp_eps, float log_scale, int flags, void *stream);

/*
 * Its adjoint: d loss / d wav from gout = d loss / d out ([rows][n_mels][n_frames], float32), i.e. what autograd
 * computes through the same op chain when BigVGAN/train_binaural_mel.py:787 calls loss_gen_all.backward() with
 * loss_mel = fn_mel_loss_multiscale(y, y_g_hat) (:759) or the single-scale L1 on mel_spectrogram(y_g_hat) (:711-720, :762).
 * Only the waveform is needed from the forward (the spectrum is recomputed).  gwav ([rows][gwav_pitch], T valid
 * samples per row) is overwritten (added to under AFA_MEL_FLAG_ACCUMULATE).  The clamp passes the gradient where mel >= clamp_eps and
 * the magnitude has gradient 0 at 0 (torch.clamp / torch.abs).  bin_mlo / bin_mhi (int32 device [n_fft / 2 + 1]):
 * the filters whose support contains bin k all lie in [bin_mlo[k], bin_mhi[k]).
 * gother / gcoef / gscale_dev: see AFA_MEL_FLAG_L1_SIGN below (NULL, 0, NULL otherwise).
 * workspace: device scratch of afa_logmel_bwd_workspace_bytes() bytes (8-byte aligned) holding the windowed frame
 * gradients between the two kernels; the overlap-add is a fixed-order gather, so gwav is bitwise reproducible.
 */
size_t afa_logmel_bwd_workspace_bytes(int64_t rows, int64_t T, int n_fft, int hop, int pad);
int afa_logmel_bwd(const float *wav, const float *gout, float *gwav, int64_t rows, int64_t T, int64_t row_pitch,
                   int64_t gwav_pitch, int n_fft, int hop, int pad, int pad_mode,
                   const float *window, const float *twiddle,
                   int n_mels, const int32_t *band_start, const int32_t *band_len, const int32_t *band_off,
                   const float *band_w, const int32_t *bin_mlo, const int32_t *bin_mhi,
                   float mag_eps, float clamp_eps, float log_scale, int flags,
                   const float *gother, float gcoef, const float *gscale_dev,
                   void *workspace, size_t workspace_bytes, void *stream);

/*
 * The L1 of the loss (loss.py:203-207: loss_fn = nn.L1Loss() on the two log-mel tensors) without its torch launches.
 * Forward: afa_l1_partial_sums writes n_partial per-CTA sums of |a - b| (fixed order; the caller adds them and divides
 * by n).  Backward: with AFA_MEL_FLAG_L1_SIGN, afa_logmel_bwd takes the two saved log-mel tensors as gout / gother and
 * uses sign(gout - gother) * gcoef * (*gscale_dev) as the output gradient (gscale_dev: device scalar holding the
 * upstream d / d loss, NULL = 1), so sign / scale never exist as tensors; AFA_MEL_FLAG_ACCUMULATE adds this scale's
 * waveform gradient to gwav.
 */
int afa_l1_partial_sums(const float *a, const float *b, int64_t n, float *partial, int n_partial, void *stream);

/*
 * Zero-frame compaction of a batch of mel spectrograms on the device:  <->  detect_and_exclude_zero_frames,
 * BigVGAN/inference_e2e.py:38-74 (called per file and channel on the host at :146-147).
 * mel / packed: float32 [rows, n_mels, T] device (rows = clips x channels); a frame is dropped when the sum of |mel| over its
 * bands is <= zero_threshold (1e-10 in the reference; the sum runs in the reference's float32 order, so the mask is
 * bit-identical).  packed[r][m][p] = mel[r][m][f] for the p-th kept frame f of row r (columns >= n_kept[r] are zero-filled),
 * frame_map[r][p] = f (-1 beyond n_kept[r]: afa_tail_fwd_cl drops such hops), n_kept: int32 [rows].  One launch for the
 * whole batch; the generator then runs on packed[:, :, :n_kept] and afa_tail_fwd_cl scatters the hops back through
 * frame_map (reconstruct_audio_with_silence, inference_e2e.py:77-111).
 */
int afa_compact_zero_frames(const float *mel, float *packed, int32_t *frame_map, int32_t *n_kept,
                            int64_t rows, int n_mels, int64_t T, float zero_threshold, void *stream);

/*
 * Tuning / introspection (used by bench.py and the tests; not part of the reference interface).
 * afa_set_tuning: which=0 forward, 1 backward; chunks = 16-byte chunks per thread segment
 * (odd, one of the compiled values), threads = CTA size.  0 keeps the built-in choice.
 * which=2: channels-last forward, segment length = 12 * chunks + 2 samples.
 * which=3: tensor path of the fused activation+convolution: chunks = 1 tcgen05 (default), 0 legacy mma.sync.
 * which=4: its input staging: chunks = 1 bulk-copy the tile's input rows into shared memory (default), 0 global loads.
 * which=5: tensor-core forward (bf16 tensors, T % 8 == 0 -- or T % 8 == 4 with an even batch * channels --, 16-byte aligned x and y: both FIR filters as banded-Toeplitz
 *          tcgen05 products, csrc/afa_tc_kernels.cuh): chunks = 0 never, 1 built-in choice (default), 2 whenever eligible;
 *          threads = blocks of 16 outputs per TMEM lane and CTA (a multiple of 4 up to 4096; 0 = built-in choice).
 * which=6: its rows per CTA as log2 (3..7; -1 = built-in choice).
 * which=7: the channels-last tensor-core forward (afa_amp_activation1d_fwd_cl on bf16 tensors without res / xsum,
 *          channels % 8 == 0 (units of 32 channels, the last one at least three quarters full), T % 4 == 0, 16-byte aligned
 *          x, y and batch strides, at most 8 zero rows behind T;
 *          csrc/afa_tc_cl_kernels.cuh): chunks = 0 never, 1 built-in choice (default), 2 whenever eligible; threads = blocks of
 *          16 outputs per CTA (a multiple of 4 up to 4096; 0 = built-in choice).  which=5 with chunks = 0 turns it off as well.
 * which=8: tail kernel (afa_tail_fwd_cl): walk length = 12 * chunks + 2 samples per warp segment (0 = built-in 98).
 * which=9: programmatic dependent launch of the two tensor-core kernels (their set-up overlaps the tail of the kernel in front
 *          of them on the stream; they touch no global memory before griddepcontrol.wait): chunks = 1 on (default), 0 off.
 * afa_kernel_info: writes {regs, static+dynamic smem bytes, threads, elems per segment,
 * max resident CTAs/SM, launches so far} for the kernel that (which, dtype, T) selects (which = 5: the tensor-core forward,
 * 7: its channels-last variant).
 */
int afa_set_tuning(int which, int chunks, int threads);
int afa_kernel_info(int which, int dtype, int64_t T, int32_t out[6]);
/* Same, for the kernel variant a launch of shape [batch, channels, T] selects (segment length depends on size). */
int afa_kernel_info_shape(int which, int dtype, int64_t batch, int64_t channels, int64_t T, int32_t out[6]);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
int64_t afa_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* AFA_B200_H_ */
