/*
 * wft.h -- C ABI of the B200-native Whisper audio front end (libwft_b200.so).
 *
 * Drop-in boundary for ONE hot path of i4Ds/whisper-finetune (paths relative to the reference root):
 *
 *   PCM -> zero pad -> Hann(400)/hop-160 STFT -> |X|^2 -> mel(80|128) -> log10 / clamp / max-8 -> (x+4)/4
 *       -> partial-segment cut -> min-value pad_or_trim -> SpecAugment time + frequency masks -> [B, n_mels, T]
 *
 * i.e. what AudioDataset.__getitem__/_calculate_mel compute per clip on the CPU
 * (src/whisper_finetune/data/data_loader.py:273-292, :344-346) and collate_fn stacks (:362-367).
 *
 * Conventions
 *   - Plain pointers and sizes only; every pointer marked "device" must be CUDA device memory of the
 *     current device; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - All calls are asynchronous on `stream`; none of them synchronises the host.
 *   - Return value 0 = success, negative = error (WFT_ERR_*).  The message for the last error of the calling
 *     thread is available from wft_last_error().  Nothing aborts, nothing throws across the boundary.
 *   - Inputs are never modified; outputs and workspaces are caller-owned (the reference's functional
 *     semantics: masked_fill / F.pad / index_select all allocate, data/utils.py:380-404).
 *   - There is NO CPU implementation behind this header: without a CUDA device every compute entry point
 *     fails with WFT_ERR_CUDA.
 */
#ifndef WFT_H_
#define WFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFT_ABI_VERSION 9

/* Front-end constants (whisper.audio: SAMPLE_RATE, N_FFT, HOP_LENGTH, CHUNK_LENGTH, N_SAMPLES, N_FRAMES;
 * imported by the reference at data_loader.py:13 and data/utils.py:10). */
#define WFT_SAMPLE_RATE 16000
#define WFT_N_FFT 400
#define WFT_HOP_LENGTH 160
#define WFT_N_SAMPLES 480000
#define WFT_N_FRAMES 3000

enum wft_pcm_dtype { WFT_PCM_F32 = 0, WFT_PCM_I16 = 1 };

/* How the few counters in the workspace get back to zero between launches (the workspace holds WFT_WS_PHASES copies of
 * them -- one tile counter and 8 bytes of statistics per clip -- plus per-tile scratch).
 *   WFT_WS_MEMSET   the library enqueues a cudaMemsetAsync in front of every launch; the workspace may hold anything.
 *   WFT_WS_PHASE_A / WFT_WS_PHASE_B   self-cleaning: a launch uses one copy and zeroes the other for the launch after it (no
 *       memset node, so back-to-back launches chain kernel to kernel).  Contract: the caller zeroes the whole workspace ONCE
 *       before its first use, passes A, B, A, B, ... on consecutive launches that use it, and never shares it between streams.
 *   WFT_WS_RING + k, k = 0 .. WFT_WS_PHASES-1   the copies are used round robin; the call with k == 0 zeroes all of them with one
 *       cudaMemsetAsync (the workspace may hold anything before it).  Contract: the caller passes k = 0, 1, ..., WFT_WS_PHASES-1,
 *       0, 1, ... on consecutive launches that use the workspace and never shares it between streams.  The only mode in which
 *       launches may overlap (WFT_LAUNCH_OVERLAP): two launches in flight never share a copy, and the memset at k == 0 is the
 *       one point per WFT_WS_PHASES launches where everything in front of it has to be complete. */
#define WFT_WS_PHASES 16
enum wft_workspace_mode { WFT_WS_MEMSET = 0, WFT_WS_PHASE_A = 1, WFT_WS_PHASE_B = 2, WFT_WS_RING = 16 /* + k */ };

/* launch_flags.
 * WFT_LAUNCH_PDL: programmatic dependent launch -- when the previous operation on the stream is a kernel, this grid is scheduled
 *   while that kernel drains (its table prologue overlaps the tail) and waits on the device before it touches anything.  Worth
 *   a few per cent on back-to-back front-end launches of one stream; it does NOT pay when other streams are busy on the same GPU
 *   (the early CTAs hold SM slots while they wait), so it is a per-call choice.
 * WFT_LAUNCH_OVERLAP (implies PDL; needs WFT_WS_RING): the caller declares this call INDEPENDENT of the wft_frontend_forward call
 *   in front of it on the stream -- it reads nothing that call writes (its `out`) and writes nothing that call reads or writes
 *   (`pcm`, `lengths`, `n_valid_frames`, `mask_params`, `out`) -- and that nothing but wft_frontend_forward calls was enqueued
 *   on the stream in between.  The grid then does not wait for the previous call's grids to complete: its CTAs start working
 *   in the SM slots the previous batch frees as it runs out of tiles, so consecutive batches run without a tail (B = 64:
 *   -12 % per launch).  Everything enqueued AFTER the call is ordered behind all of it as usual. */
#define WFT_LAUNCH_PDL 1
#define WFT_LAUNCH_OVERLAP 2

enum wft_status {
  WFT_OK = 0,
  WFT_ERR_INVALID = -1, /* bad argument (shape, dtype, n_mels not in {80,128}, ...) -> ValueError in Python */
  WFT_ERR_CUDA = -2,    /* CUDA runtime failure (no device, launch error, ...)       -> RuntimeError        */
};

/* Per-call description of a batched front-end pass (wft_frontend_forward). */
typedef struct wft_frontend_args {
  /* ---- input PCM ---- */
  const void* pcm;          /* device; clip b starts at element b*clip_stride                                   */
  int32_t pcm_dtype;        /* enum wft_pcm_dtype; int16 is scaled by 1/32768 like whisper.audio.load_audio      */
  int32_t batch;            /* B >= 1                                                                           */
  int64_t clip_stride;      /* elements between consecutive clips                                               */
  int32_t n_samples;        /* samples stored per clip (<= clip_stride)                                         */
  int32_t padding;          /* zeros appended after n_samples: the `padding` arg of log_mel_spectrogram;
                               together with `lengths` this is also the zero pad of data_loader.py:346          */
  const int32_t* lengths;   /* device [B] or NULL: valid prefix of each clip (rest treated as 0), <= n_samples  */
  /* ---- features ---- */
  int32_t n_mels;           /* 80 or 128 (whisper.audio asserts the same set)                                   */
  int32_t n_frames_out;     /* T of the output; 0 -> (n_samples+padding)/160 (no pad_or_trim)                   */
  const int32_t* n_valid_frames; /* device [B] or NULL: keep frames [0, n_valid) (data_loader.py:279-280), then
                               min-value pad up to n_frames_out (data/utils.py:380-404); <0 = keep all           */
  /* ---- SpecAugment (data_loader.py:286-287), all four NULL = no masking ---- */
  const int32_t* mask_params; /* device [B,4] = (t0, t1, f0, f1): frames [t0,t1) and rows [f0,f1) := mask_value  */
  float mask_value;         /* 0.0f in the reference (torchaudio default)                                        */
  /* ---- output ---- */
  float* out;               /* device [B, n_mels, n_frames_out] contiguous                                      */
  void* workspace;          /* device, >= wft_frontend_workspace_bytes(); contents are scratch                  */
  size_t workspace_bytes;
  int32_t workspace_mode;   /* enum wft_workspace_mode; 0 (a zero-initialised struct) = WFT_WS_MEMSET                */
  int32_t launch_flags;     /* WFT_LAUNCH_*; 0 = plain launch                                                    */
  /* ---- optional: draw the SpecAugment intervals inside this call (mask_params must be NULL) ----
   * draw_masks != 0: the library launches wft_specaug_draw's kernel into a corner of the workspace right in front of the
   * fused kernel -- (seed, clip_offset, time / freq mask parameter, p) mean what they mean for wft_specaug_draw, n_mels and
   * n_frames_out are this call's -- so that one augmented batch is ONE call from the host. */
  int32_t draw_masks;
  int32_t draw_time_mask_param;
  int32_t draw_freq_mask_param;
  float draw_p;
  uint64_t draw_seed;
  uint64_t draw_clip_offset;
} wft_frontend_args;

/* ABI version of the loaded library (== WFT_ABI_VERSION of the header it was built from). */
int wft_abi_version(void);

/* Message of the last error raised on the calling thread ("" if none). */
const char* wft_last_error(void);

/* Scratch bytes needed by wft_frontend_forward for (batch, n_samples+padding, n_frames_out). */
int wft_frontend_workspace_bytes(int32_t batch, int32_t n_samples_total, int32_t n_frames_out, size_t* bytes);

/* The fused front end: replaces, for a whole batch and in one launch,
 *   np.pad(audio, (0, N_SAMPLES - len))                    data_loader.py:346
 *   whisper.audio.log_mel_spectrogram(audio, n_mels)       data_loader.py:278   (per-clip max-8 floor)
 *   mel[:, :int(start * 100)]                              data_loader.py:279-280
 *   pad_or_trim(mel, N_FRAMES)                             data_loader.py:281-282, data/utils.py:380-404
 *   time_masking(mel); freq_masking(mel)                   data_loader.py:286-287
 *   pad_sequence(x, batch_first=True)                      data_loader.py:362-367 (collate_fn)
 */
int wft_frontend_forward(const wft_frontend_args* args, void* stream);

/* Stand-alone pad_or_trim on float32 (data/utils.py:380-404): `in` is [outer, len_in, inner], `out` is
 * [outer, len_out, inner]; trims, or right-pads with the minimum over the WHOLE input (computed on device,
 * no host sync).  `scratch` is a device buffer of >= 16 bytes.  len_in == 0 with len_out > 0 is an error
 * (the reference's torch.min raises on an empty tensor). */
int wft_pad_or_trim_f32(const float* in, int64_t outer, int64_t len_in, int64_t inner, int64_t len_out,
                        float* out, void* scratch, void* stream);

/* Stand-alone SpecAugment masks (torchaudio mask_along_axis semantics as used at data_loader.py:286-287):
 * out[b, r, t] = (t0<=t<t1 || f0<=r<f1) ? mask_value : in[b, r, t]; `in == out` is allowed (in place).
 * mask_params is device int32 [B,4].  Also serves [B, C, T] activations (model/model_utils.py:382-437). */
int wft_specaug_apply_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames,
                          const int32_t* mask_params, float mask_value, void* stream);

/* Counter-based draw of the mask intervals on the device: Philox4x32-10 keyed by `seed`, counter = global
 * clip index (clip_offset + b) so that a rank's shard draws the same masks at any world size.  Interval
 * arithmetic is torchaudio's (float32): width = u*param, start = trunc(u'*(size-width)), end = start+trunc(width).
 * `p` is the gate of data_loader.py:294-301 (p>=1 always, p<=0 never, else u_gate < p).
 * mask_params_out is device int32 [B,4]. */
int wft_specaug_draw(uint64_t seed, uint64_t clip_offset, int32_t batch, int32_t n_mels, int32_t n_frames,
                     int32_t time_mask_param, int32_t freq_mask_param, float p, int32_t* mask_params_out,
                     void* stream);

/* ---- "next" rows of the path, stand-alone ------------------------------------------------------------
 * SpecAugment time-warp, the step between pad_or_trim and the masks (data_loader.py:285, data/utils.py:41-143):
 * out[b] = in[b] resampled along time through the 3-knot cubic Hermite map defined by (warp_p, warp_d), bilinear,
 * zeros outside, grid_sample(align_corners=True) semantics.  warp_params is device int32 [B,2]; in != out. */
int wft_time_warp_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames,
                      const int32_t* warp_params, void* stream);

/* Counter-based draw of (warp_p in [W, T-W), warp_d in [-W, W)) per clip (the reference's torch.randint ranges,
 * data/utils.py:107-111), Philox keyed like wft_specaug_draw; (-1, 0) = "no warp" when the p gate rejects the clip (both
 * wft_time_warp_f32 and wft_augment_f32 copy such a clip unchanged). */
int wft_time_warp_draw(uint64_t seed, uint64_t clip_offset, int32_t batch, int32_t n_frames, int32_t time_warp_w, float p,
                       int32_t* warp_params_out, void* stream);

/* Fused augmentation epilogue: everything AudioDataset._calculate_mel does to the finished features when the SpecAugment
 * gate passes (data_loader.py:284-290: time_warping -> time_masking -> freq_masking -> extreme_freq_masking), in ONE read and
 * ONE write of the features instead of one pass per transform:
 *   out[b, r, t] = (t0 <= t < t1 || f0 <= r < f1 || r < lo || r >= n_rows - hi) ? mask_value : warp_b(in[b])[r, t]
 * warp_params  device int32 [B,2] = (warp_p, warp_d) as in wft_time_warp_f32, or NULL = no warp; a clip whose warp_p is
 *              outside (0, n_frames - 1) -- wft_time_warp_draw writes (-1, 0) when the p gate rejects it -- is copied.
 * mask_params  device int32 [B,4] = (t0, t1, f0, f1) or NULL.
 * extremes     device int32 [B,2] = (lo, hi): rows masked from the bottom / top (ExtremesFrequencyMasking,
 *              data/utils.py:146-190) or NULL.
 * spline_f32   0: the warp's cubic Hermite source map is evaluated in float64 and rounded once; 1: the reference's own
 *              float32 evaluation order (data/utils.py:65-93) is restated, which lands on the reference's coordinate
 *              wherever torch.pow returned the correctly rounded power.
 * in == out is allowed only without a warp. */
int wft_augment_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames, const int32_t* warp_params,
                    const int32_t* mask_params, const int32_t* extremes, float mask_value, int32_t spline_f32, void* stream);

/* The same epilogue with the clip parameters drawn INSIDE the kernel: intervals = wft_specaug_draw(seed, clip_offset, batch,
 * n_rows, n_frames, time_mask_param, freq_mask_param, p), warp points = wft_time_warp_draw(seed, clip_offset, batch, n_frames,
 * time_warp_w, p) -- bit for bit the same draws, without the two launches in front (an augmented production batch is then
 * front-end grid -> fix-up grid -> this grid).  time_warp_w == 0 = masks only (in == out allowed). */
int wft_augment_drawn_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames, uint64_t seed,
                          uint64_t clip_offset, int32_t time_mask_param, int32_t freq_mask_param, int32_t time_warp_w, float p,
                          const int32_t* extremes, float mask_value, int32_t spline_f32, void* stream);

/* A production batch as ONE call and TWO grids (every reference config time-warps: configs/config_turbo_best.yaml:97,
 * config_large_v3_best_muon_ddp4.yaml:120; data_loader.py:273-292 in order): the front-end grid of wft_frontend_forward writes
 * the un-augmented features into `fe->out` (scratch of the output's shape), and the augmentation epilogue right behind it
 * finishes every cell as it loads it -- the max-8 floor, the clamp value of tiles that were never computed, the min-value
 * pad beyond the kept frames: what the fix-up grid of wft_frontend_forward would have rewritten in place -- and writes
 * warp -> time mask -> frequency mask -> extremes mask of it to `aug->out`.  Bit-identical to wft_frontend_forward followed by
 * wft_augment_f32 / wft_augment_drawn_f32, one grid and one host call fewer.  `fe->out` holds unspecified values afterwards.
 * `fe` as for wft_frontend_forward with mask_params == NULL and draw_masks == 0 (the masks belong to `aug`); workspace modes
 * and launch flags mean the same (WFT_LAUNCH_OVERLAP: independent of the front-end / front-end + augmentation call in front). */
typedef struct wft_augment_args {
  float* out;                 /* device [B, n_mels, n_frames_out] contiguous, != fe->out                                   */
  const int32_t* warp_params; /* device [B,2] or NULL, as for wft_augment_f32                                             */
  const int32_t* mask_params; /* device [B,4] or NULL                                                                     */
  const int32_t* extremes;    /* device [B,2] or NULL                                                                     */
  float mask_value;
  int32_t spline_f32;
  int32_t draw;               /* != 0: warp points and intervals drawn inside the kernel (wft_augment_drawn_f32's draws);
                                 warp_params and mask_params must be NULL                                                 */
  int32_t draw_time_mask_param;
  int32_t draw_freq_mask_param;
  int32_t draw_time_warp_w;
  float draw_p;
  uint64_t draw_seed;
  uint64_t draw_clip_offset;
} wft_augment_args;
int wft_frontend_augment_forward(const wft_frontend_args* fe, const wft_augment_args* aug, void* stream);

/* Deep SpecAugment on encoder activations (model/model_utils.py:382-437: permute -> TimeMasking -> FrequencyMasking ->
 * permute on every hooked layer-norm output), without the permutes: x is [batch, seq, dim] contiguous with 16-bit (fp16 /
 * bf16) or 32-bit elements, out[b, s, d] = (t0 <= s < t1 || f0 <= d < f1) ? fill : x[b, s, d].  ONE mask for the whole batch,
 * like torchaudio's mask_along_axis on a 3-D input; the intervals are host integers because the reference draws them on
 * the host (torch.rand).  fill_bits is the fill value's bit pattern in the element type (0 for the reference's 0.0).
 * in == out is allowed.  The backward of this op is the same call on the gradient with fill 0. */
int wft_mask_bsd(const void* in, void* out, int32_t elem_bytes, int64_t batch, int32_t seq, int32_t dim, int32_t t0,
                 int32_t t1, int32_t f0, int32_t f1, uint32_t fill_bits, void* stream);

/* Introspection used by bench.py / tests: number of kernel launches issued by this library on the calling
 * thread since the last reset, and the persistent grid the fused kernel would use on the current device. */
int64_t wft_launch_count(int reset);
int wft_frontend_grid(int32_t n_mels, int32_t pcm_dtype, int32_t* ctas, int32_t* threads, int32_t* smem_bytes);

/* Test hook: cap the persistent grids of wft_frontend_forward (front-end and fix-up kernel) at `max_ctas` CTAs (0 = no cap, the
 * default).  Results must not depend on the grid. */
int wft_debug_set_max_ctas(int32_t max_ctas);
/* Test hook: != 0 makes the augmentation epilogue use its generic kernel (taps straight from global memory) even where the staged
 * one (source windows by bulk copy through shared memory; n_frames % 4 == 0) applies.  Both must agree bit for bit. */
int wft_debug_set_augment_generic(int32_t on);
/* Development hook (tools/cosched_probe.py): pad the front-end CTA's dynamic shared memory by `bytes`, i.e. lower its CTAs per
 * SM (0 = the default, 6 per SM on B200). */
int wft_debug_set_extra_smem(int32_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* WFT_H_ */
