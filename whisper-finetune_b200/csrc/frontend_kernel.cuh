// Fused Whisper audio front end for sm_100a: PCM -> log-mel (+ cut / min-pad / SpecAugment masks): the front-end grid and the
// lean fix-up grid that runs behind it.
//
// Replaces, for a whole batch, the per-clip CPU path of the reference
//   np.pad -> whisper.audio.log_mel_spectrogram -> mel[:, :T'] -> pad_or_trim -> time/freq masks -> collate
//   (src/whisper_finetune/data/data_loader.py:346, :278, :279-282, :286-287, :362-367; data/utils.py:380-404).
//
// Work decomposition
//   tile      = 16 consecutive frames of one clip = 8 frame PAIRS; 160 threads = 8 pairs x 20 slots, pair index fastest
//               (pair_coord_a / pair_coord_b); 6 CTAs per SM.
//   pair      = frames (2q, 2q+1) packed as re/im of ONE 400-point complex FFT (two real frames per transform).
//   400-point = 20 x 20 Cooley-Tukey, one register-resident 20-point DFT (dft20.cuh, packed f32x2) per thread per stage:
//               stage A: thread (q, n2) transforms over n1, multiplies by W400^(n2 k1), scatters to the exchange;
//               stage B: thread (q, k1) transforms over n2, keeps Z[k1 + 20 k2] for k2 < 10 in registers and receives the
//                        upper half of the mirror row 20 - k1 by warp shuffle (both rows sit in one warp);
//               power  : thread (q, j) pairs Z[j + 20 m] with its mirror Z[400 - j - 20 m] and separates the two real
//                        spectra: 4|Xa|^2 = |Z[k] + conj Z[400-k]|^2, 4|Xb|^2 = |Z[k] - conj Z[400-k]|^2.
//   mel phase = thread <-> one mel row x 8 consecutive frames per pass: the power tile is stored [bin][frame], so one tap of
//               the sparse triangular filter is LDS.128 + 2 FFMA2 per 4 frames with the weight held in a register; rows
//               are dealt to the warps in groups of 16 by tap count (wft_tables.inc: one or two passes per warp) and each
//               pass's tap loop is fully unrolled; log10 via MUFU.LG2; the FINAL feature (L + 4) / 4 with the SpecAugment
//               masks applied is written once to `out` as 32-byte stores (STG.256).  The intervals of the masks may be
//               drawn in here (Philox keyed by the global clip index), so an augmented batch needs no launch in front.
//   per-clip max / min = ordered-int red.max into the workspace (fire and forget), plus the tile's own minimum in a
//               per-tile slot.  The few cells that can only be finished once the WHOLE clip is known -- the max-8 floor
//               where it binds, the min-value pad of the frames beyond the kept part, silent (all-zero PCM) tiles that
//               were never computed -- are left to `fixup_kernel`, a small second grid launched right behind this one:
//               it looks at every tile's minimum against its clip's maximum and rewrites only the tiles that still
//               differ from their final value (none at all for ordinary audio).  The kernel boundary is the only
//               synchronisation between the two: no completion counters, no waiting CTAs, no fences in here.
//   scheduling = persistent CTAs pulling tiles from an atomic counter in clip-major order, one tile ahead; the next tile's
//               PCM travels global->shared by TMA bulk copies (mbarrier completion) under the current tile's mel phase.
//               A CTA that runs out of tiles exits at once, and the grid behind it on the stream (programmatic dependent
//               launch) takes its SM slot; a launch flagged WFT_LAUNCH_OVERLAP does not even wait for the grid in front of
//               it to complete (independent batches), so consecutive batches run back to back without a tail.
//
// Shared memory per CTA: one 28.8 KB region time-multiplexed as
//   [audio tile at the top] -> stage A->B exchange -> [power tile at the bottom | next audio tile at the top],
// plus ~8.4 KB of window / twiddle / mel-weight tables, drawn intervals and control words.  The thread order and every stride
// in here were chosen against measured shared-memory wavefront costs (tools/micro/smem_wavefronts.cu,
// profiles/r01_smem_wavefront_probe.md): the SM (L1 data pipe at 0.80 of its peak, FMA pipe 48 %, issue 2.3 of 4), not HBM, is what bounds
// this kernel (profiles/r02_ncu_summary.md, profiles/r02_ab_experiments.md).  The hot loop is one DFT20 copy per stage and
// ~2.8 k SASS instructions so that it stays resident in the SM's instruction cache.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dft20.cuh"
#include "wft_tables.inc"

namespace wft {

constexpr int kHop = 160;
constexpr int kNfft = 400;
constexpr int kTileFrames = 16;
constexpr int kPairs = kTileFrames / 2;                          // 8
constexpr int kPairThreads = 20;
constexpr int kThreads = kPairs * kPairThreads;                  // 160
constexpr int kWarps = kThreads / 32;                            // 5
constexpr int kTileSamples = kTileFrames * kHop + (kNfft - kHop);  // 2800
constexpr int kSkewBlock = 320;                                  // samples per frame pair
// extra elements after every block of 320 samples: keeps stage A's stride-20 gathers on distinct banks AND every block on
// a 16-byte boundary for the bulk copies: 20 floats (1360 B per block) / 24 int16 (688 B per block)
template <typename PcmT>
struct Skew {
  static constexpr int value = sizeof(PcmT) == 4 ? 20 : 24;
};
constexpr int kAudioElems = kTileSamples + 20 * ((kTileSamples - 1) / kSkewBlock);     // 2960 float32 slots (int16 needs less)
constexpr int kRowStride = 44;                                   // floats per exchange row (20 complex + pad)
constexpr int kPairStride = 20 * kRowStride + 20;                // 900: pair stride == 4 (mod 32) banks (pair_coord_a explains)
constexpr int kRegionFloats = kPairs * kPairStride;              // 7200 floats = 28800 B
constexpr int kPStride = WFT_MEL_P_STRIDE;                       // power tile [bin][frame]: 20 floats per bin (16 frames + pad)
constexpr int kPFloats = 200 * kPStride;                         // bins 0..199 (bin 200 carries no mel weight)
constexpr int kAudioBase = kRegionFloats - kAudioElems;          // float32 audio tile sits at the TOP of the region
constexpr int kWinFloats = WFT_WINDOW_TABLE_LEN;                 // 400
constexpr int kTwFloats = WFT_TWIDDLE_TABLE_LEN;                 // 880
constexpr int kTwRow = WFT_TWIDDLE_ROW;                          // 44
constexpr int kMelWFloats = ((WFT_MEL80_W_LEN > WFT_MEL128_W_LEN ? WFT_MEL80_W_LEN : WFT_MEL128_W_LEN) + 3) & ~3;
constexpr int kCtlInts = 64;
constexpr int kDrawCache = 64;    // clips whose drawn SpecAugment intervals a CTA keeps in shared memory (batches up to this size)
constexpr int kSmemBytes = (kRegionFloats + kWinFloats + kTwFloats + kMelWFloats) * 4 + kCtlInts * 4 + kDrawCache * 16;

static_assert(kPFloats <= kAudioBase, "power tile and prefetched audio tile must not overlap");
static_assert((kAudioBase & 3) == 0 && ((kSkewBlock + Skew<float>::value) * 4) % 16 == 0 &&
                  ((kSkewBlock + Skew<int16_t>::value) * 2) % 16 == 0 && (kSkewBlock & 7) == 0 && ((kTileSamples % kSkewBlock) & 7) == 0,
              "every bulk copy (f32 or i16) must start and end on 16 bytes");
static_assert(6 * (kSmemBytes + 1024) <= 228 * 1024, "6 CTAs per SM: 6 x (dynamic + 1 KB reserved) must fit 228 KB");

__device__ __align__(16) const float g_window_table[kWinFloats] = WFT_WINDOW_TABLE_INIT;
__device__ __align__(16) const float g_twiddle_table[kTwFloats] = WFT_TWIDDLE_TABLE_INIT;
__device__ __align__(16) const float g_mel80_w[WFT_MEL80_W_LEN] = WFT_MEL80_W_INIT;
__device__ __align__(16) const float g_mel128_w[WFT_MEL128_W_LEN] = WFT_MEL128_W_INIT;
static_assert(kWinFloats % 4 == 0 && kTwFloats % 4 == 0 && WFT_MEL80_W_LEN % 4 == 0 && WFT_MEL128_W_LEN % 4 == 0,
              "the tables are copied to shared memory as 16-byte vectors");
__device__ const uint32_t g_mel80_thread[kThreads] = WFT_MEL80_THREAD_INIT;
__device__ const uint32_t g_mel128_thread[kThreads] = WFT_MEL128_THREAD_INIT;

// mel plan (gen_tables.py): a thread owns one mel row x 8 frames per PASS (lane l: row slot l % 16, frames 8 * (l / 16) ..);
// warp w runs one or two passes, pass k with T taps; its weights start at WBASE(w) (tap j of slot i at (taps before + j) * 16 + i)
constexpr int kMelSlots = 16;
template <int NM>
__host__ __device__ constexpr int mel_warp_passes(int w) {
  constexpr int a[kWarps] = WFT_MEL80_WARP_PASSES, b[kWarps] = WFT_MEL128_WARP_PASSES;
  return NM == 80 ? a[w] : b[w];
}
template <int NM>
__host__ __device__ constexpr int mel_warp_taps(int w, int pass) {
  constexpr int a0[kWarps] = WFT_MEL80_WARP_T0, a1[kWarps] = WFT_MEL80_WARP_T1;
  constexpr int b0[kWarps] = WFT_MEL128_WARP_T0, b1[kWarps] = WFT_MEL128_WARP_T1;
  return NM == 80 ? (pass ? a1[w] : a0[w]) : (pass ? b1[w] : b0[w]);
}
template <int NM>
__host__ __device__ constexpr int mel_warp_wbase(int w) {
  constexpr int a[kWarps] = WFT_MEL80_WARP_WBASE, b[kWarps] = WFT_MEL128_WARP_WBASE;
  return NM == 80 ? a[w] : b[w];
}

struct ClipStat {      // zeroed before the launch (0 is below every encoding)
  uint32_t max_enc;   // ordered-int encoding of max L2 = log2(mel) over ALL frames of the clip
  uint32_t min_inv;   // ~encoding of min L2 over the KEPT frames (pad value of pad_or_trim)
};

struct FrontendParams {
  const void* pcm;
  int64_t clip_stride;
  const int32_t* lengths;
  const int32_t* n_valid;
  const int32_t* masks;
  float* out;
  uint32_t* tile_counter;
  ClipStat* stats;
  float* tile_min;      // [total_tiles] min L2 over the live cells of every computed tile (what fixup_kernel decides on)
  int32_t n_samples;
  int32_t n_total;      // n_samples + padding
  int32_t batch;
  int32_t n_frames;     // n_total / 160
  int32_t n_frames_out;
  int32_t tiles_per_clip;
  int32_t total_tiles;
  float mask_value;
  uint32_t tpc_magic;   // floor(2^32 / tiles_per_clip): tile -> clip by multiply-high (+ one correction step)
  int32_t vec_ok;       // `out` is 32-byte aligned and n_frames_out % 8 == 0: the mel phase may use 32-byte stores
  uint4* clean;         // self-cleaning workspace: the header + statistics of the OTHER phase, zeroed by CTA 0 for the launch
  int32_t clean_vec;    // after this one (16-byte vectors; 0 = zeroed by the host)
  int32_t draw;         // != 0: masks == nullptr and the intervals are drawn in here (draw_mask_intervals)
  int32_t draw_tparam, draw_fparam;
  float draw_p;
  uint64_t draw_seed, draw_clip_offset;
  int32_t overlap;      // != 0 (WFT_LAUNCH_OVERLAP): independent of the grids in front of it on the stream -- does not wait for
                        // them to complete, so its CTAs work in the SM slots their tails free
  int32_t chunk;        // consecutive tiles a CTA takes per claim (>= 1): neighbours share the clip, so the per-clip loads of
                        // describe_tile hit its memo and the tile counter sees 1 / chunk of the atomics
};

__device__ __forceinline__ uint32_t enc_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t e) {
  const uint32_t b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  return __uint_as_float(b);
}

// ---- SpecAugment intervals drawn on the device (wft_specaug_draw / draw_masks): Philox4x32-10 keyed by the seed, counter =
// global clip index, so the draw is a pure function of (seed, clip) -- any thread that needs a clip's intervals computes them
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t w) { return static_cast<float>(w >> 8) * 5.9604644775390625e-08f; }

// torchaudio interval: width = u*param ; start = trunc(u' * (size - width)) ; end = start + trunc(width)
__device__ __forceinline__ void interval(float u_w, float u_s, int param, int size, int& a, int& b) {
  if (param < 1) { a = 0; b = 0; return; }
  const float value = __fmul_rn(u_w, static_cast<float>(param));
  const float minv = __fmul_rn(u_s, __fsub_rn(static_cast<float>(size), value));
  a = static_cast<int>(minv);
  b = a + static_cast<int>(value);
}

// (t0, t1, f0, f1) of one clip; `p` is the gate of data_loader.py:294-301 (block 1 of the clip's counter), block 0 the intervals
__device__ __noinline__ int4 draw_mask_intervals(uint64_t seed, uint64_t clip_index, int n_mels, int n_frames, int tparam,
                                                 int fparam, float p) {
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const uint32_t lo = static_cast<uint32_t>(clip_index), hi = static_cast<uint32_t>(clip_index >> 32);
  uint32_t r[4];
  philox4x32_10(lo, hi, 0u, 0u, k0, k1, r);
  bool apply = p >= 1.0f;
  if (!apply && p > 0.0f) {
    uint32_t g[4];
    philox4x32_10(lo, hi, 1u, 0u, k0, k1, g);
    apply = u01(g[0]) < p;
  }
  int4 mk = make_int4(0, 0, 0, 0);
  if (apply) {
    interval(u01(r[0]), u01(r[1]), tparam, n_frames, mk.x, mk.y);
    interval(u01(r[2]), u01(r[3]), fparam, n_mels, mk.z, mk.w);
  }
  return mk;
}

struct TileCoord {
  int clip, t0;
};
__device__ __forceinline__ TileCoord tile_coord(const FrontendParams& p, int tile) {
  TileCoord c;
  c.clip = tile / p.tiles_per_clip;
  c.t0 = (tile - c.clip * p.tiles_per_clip) * kTileFrames;
  return c;
}

// ---- stage 0: 5360 samples of a tile -> shared memory, in the PCM's own type, skewed ---------------------------
// sample m of the tile lives at element m + skew * (m / 320), skew = 20 (float32) or 24 (int16).  int16 PCM stays int16 in shared memory and is
// converted when stage A gathers it (the 2^-15 scale is folded into the window table, exactly).

// p[j] of the reflect-padded, zero-extended clip; j is relative to sample 0 of the un-padded clip.
template <typename PcmT>
__device__ __forceinline__ PcmT load_sample(const PcmT* __restrict__ x, int j, int len, int n_total) {
  const int r = j < 0 ? -j : (j >= n_total ? 2 * (n_total - 1) - j : j);
  return (r >= 0 && r < len) ? x[r] : PcmT(0);
}

// every sample the tile touches is a plain in-range sample and the source is 16-byte aligned (bulk-copy requirement)
template <typename PcmT>
__device__ __forceinline__ bool tile_is_interior(const PcmT* x, int g0, int len) {
  return g0 >= 0 && g0 + kTileSamples <= len && (reinterpret_cast<uintptr_t>(x + g0) & 15) == 0;
}

// ---- TMA (bulk async copy) plumbing: one elected thread moves a whole tile global -> shared, an mbarrier counts the
// bytes, every thread waits on the barrier's phase before it gathers from the tile ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// PCM is read exactly once: its lines are marked evict-first in L2, so that the 98 MB of features a B = 64 launch writes are
// what stays in the 126 MB L2 for whoever reads them next (the fix-up grid, the augmentation epilogue, the encoder)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_evict_first_policy()) : "memory");
}

// interior tile: the 2800 samples travel as 9 bulk copies (8 blocks of 320 samples + a 240-sample tail), block b landing
// at element (320 + skew) * b (the skew keeps stage A's stride-20 gathers conflict free).  ONE warp issues them, lane b sends
// block b and lane 0 announces the bytes of the whole tile (the mbarrier expects one arrival per phase): the issue sequence
// (~40 instructions: fence, shared-window addresses, cache policy, the copy) then costs the CTA once.  A single THREAD issuing
// all nine held its warp back by ~2000 cycles per tile (tools/timeline.cu); lane 0 of every warp sending two blocks each
// fixed that, but all five warps then executed the sequence -- 300 of a tile's 4600 warp-instructions.
constexpr int kPrefetchWarp = 2;   // (not warp 0: it describes the next tile and publishes the statistics)
template <typename PcmT>
__device__ __forceinline__ void prefetch_audio_tile(PcmT* __restrict__ sm_audio, const PcmT* __restrict__ src, uint64_t* bar, int lane) {
  constexpr int kBlocks = (kTileSamples + kSkewBlock - 1) / kSkewBlock;  // 9
  constexpr uint32_t kFull = kSkewBlock * sizeof(PcmT), kTail = (kTileSamples - (kBlocks - 1) * kSkewBlock) * sizeof(PcmT);
  static_assert(kBlocks <= 32, "one lane per block");
  if (lane < kBlocks) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy use of the region comes first
  if (lane == 0) mbar_expect_tx(bar, kTileSamples * sizeof(PcmT));
  __syncwarp();
  if (lane < kBlocks)
    tma_bulk_g2s(sm_audio + lane * (kSkewBlock + Skew<PcmT>::value), src + lane * kSkewBlock, lane == kBlocks - 1 ? kTail : kFull, bar);
}

// generic synchronous staging (clip edges, ragged lengths, unaligned sources)
template <typename PcmT>
__device__ __noinline__ void stage_audio_edge(PcmT* __restrict__ sm_audio, const PcmT* __restrict__ x, int g0,
                                              int len, int n_total, int tid) {
  // every load of the thread in flight before its first store: as a loop of dependent round trips (18 x global latency) an
  // edge-staged tile -- the first and the last tile of every clip -- took nearly twice as long as an interior one
  constexpr int kPer = (kTileSamples + kThreads - 1) / kThreads;   // 18
  PcmT v[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int m = tid + k * kThreads;
    v[k] = m < kTileSamples ? load_sample(x, g0 + m, len, n_total) : PcmT(0);
  }
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int m = tid + k * kThreads;
    if (m < kTileSamples) sm_audio[m + Skew<PcmT>::value * (m / kSkewBlock)] = v[k];
  }
}

__device__ __forceinline__ float pcm_as_float(float v) { return v; }
__device__ __forceinline__ float pcm_as_float(int16_t v) { return static_cast<float>(v); }

__device__ __forceinline__ float fast_log2(float x) {  // x is a normal float here: plain MUFU.LG2
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// whole-warp float max / min in one instruction (CREDUX.{MAX,MIN}.F32, sm_100a)
__device__ __forceinline__ float warp_max(float x) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float warp_min(float x) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x));
  return r;
}

// (pair, slot) of this thread, re-derived from %tid wherever a phase needs it: the volatile read keeps the compiler from
// carrying ~10 phase-local shared-memory addresses in registers (or on the stack) across the whole tile loop
// opaque copy of a register value: stops the compiler from keeping addresses derived from it alive (on the stack) for a
// whole tile when one ADD re-derives them
__device__ __forceinline__ int launder(int x) {
  asm volatile("" : "+r"(x));
  return x;
}
//
// Thread <-> (pair q, slot r) with the PAIR index fastest: a warp holds 8 pairs x 4 slots, so everything that depends on the
// slot only -- the window column, the twiddle row -- is a quarter-warp-uniform 16-byte load (2 wavefronts instead of 4.8
// when 20 consecutive lanes walked 20 different rows), and with a pair stride of 900 floats (== 4 banks mod 32) the eight
// pairs of a quarter warp land on eight different 16-byte bank groups: exchange stores 2, row loads / mirror hand-off 4
// wavefronts per instruction, i.e. their ideal (tools/micro/smem_wavefronts.cu measures every pattern used here).
// Stages A and B need not agree on the slot a thread plays (nothing but shared memory crosses the stage boundary):
//   stage A (slot = n2): slots in thread order;
//   stage B / power (slot = k1 = j): two constraints pick the rows of a warp.  (1) Row j needs the upper half of row
//   20 - j (the mirror bins): with both rows in one warp, 16 lanes apart, the hand-off is 20 SHFL per thread instead of a
//   round trip through shared memory and two CTA barriers.  (2) The power store writes 16 consecutive floats of row j per
//   quarter warp at a row stride of 20 floats, conflict free only when the two rows of a HALF warp are 4 (mod 8) apart.
//   Warps 0..3 hold rows {a, a+4 | 20-a, 16-a}, a = 1..4; warp 4 holds what is left, {0, 10 | 9, 11}: rows 0 and 10 mirror
//   themselves, 9 and 11 sit 8 lanes apart, and its power stores keep a 2-way conflict.
struct PairCoord {
  int q, r;
};
__device__ __forceinline__ PairCoord pair_coord_a() {
  int t;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
  PairCoord c;
  c.q = t & (kPairs - 1);
  c.r = t >> 3;
  return c;
}
__device__ __forceinline__ PairCoord pair_coord_b() {
  int t;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
  PairCoord c;
  c.q = t & (kPairs - 1);
  // rows of warp w = slots 4w .. 4w+3, one byte each
  const uint32_t rows = (t >> 5) == 0 ? 0x0f130501u : (t >> 5) == 1 ? 0x0e120602u : (t >> 5) == 2 ? 0x0d110703u
                      : (t >> 5) == 3 ? 0x0c100804u : 0x0b090a00u;
  c.r = static_cast<int>((rows >> (((t >> 3) & 3) * 8)) & 0xffu);
  return c;
}
// lane that plays the mirror row (20 - r) % 20 of the same pair (see pair_coord_b)
__device__ __forceinline__ int mirror_lane() {
  int t;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
  const int lane = t & 31;
  return t < 128 ? (lane ^ 16) : (lane < 16 ? lane : (lane ^ 8));
}
static_assert(kPairs == 8 && kThreads == 160, "pair_coord_a / pair_coord_b are written for 8 pairs x 20 slots");

constexpr float kLog10Of2 = 0.301029995663981195f;
constexpr float kFeatScale = 0.25f * kLog10Of2;   // (log10 x + 4) / 4 == log2 x * kFeatScale + 1 (0.25 * c is exact)

// Everything downstream of the mel sum works on L2 = MUFU.LG2(max(mel, 1e-10)): the clip statistics are max / min of L2,
// the feature is ONE monotone FMA of it, and the floor / pad values are the same FMA applied to the statistics, so
//   max(feature, floor)  and  pad = min over the kept, floored features
// hold bit for bit whatever the rounding of the FMA is.
__device__ __forceinline__ float feature_of_l2(float l2) { return fmaf(l2, kFeatScale, 1.0f); }
// feature value of the clip's dynamic-range floor, (max log10 - 8 + 4) / 4 (whisper.audio: log_spec.max() - 8.0)
__device__ __forceinline__ float floor_feature(float max_l2) {
  return fmaf(__fsub_rn(__fmul_rn(max_l2, kLog10Of2), 8.0f), 0.25f, 1.0f);
}
// L2 of the 1e-10 clamp exactly as the mel phase computes it for an all-zero frame
__device__ __forceinline__ float silent_l2() { return fast_log2(1e-10f); }

// three-input max / min (FMNMX3, sm_100+)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// (only on paths the compiler cannot if-convert: ptxas 12.9 turned this asm into ONE 4-byte STG when it sat in a short
// predicated branch of mel_edge -- that path uses two 16-byte stores)
__device__ __forceinline__ void st_global_256(float* dst, const float (&v)[8]) {   // one full 32-byte sector per thread
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// ---- mel projection, fast tiles -------------------------------------------------------------------------------------------
// A thread owns ONE mel row for NF consecutive frames (gen_tables.py deals the rows to the warps by tap count, so T, NF and
// the weight stride WS are warp constants and this is fully unrolled, branch free code).  pk = &P[start_bin][first frame] in
// the [bin][frame] power tile, wp = the thread's weight column (tap j at wp[j * WS]).  Taps run in ascending bin order;
// padded taps carry a zero weight on an always-finite bin, so the sum is exactly the dense row product.
// Fast tile: all 16 frames are real, stored, all kept or all cut, none (or all) inside the time mask; sc / of fold the
// row / tile mask into the feature FMA (masked: 0 * L2 + mask_value).
template <int T, int NF, int WS>
__device__ __forceinline__ void mel_fast(const float* __restrict__ pk, const float* __restrict__ wp, float sc, float of,
                                         float* __restrict__ dst, float& mx, float& mn) {
  constexpr int CH = NF >= 8 ? 8 : 4;   // frames per pass: 8 (one 32-byte store) or 4 (16 bytes)
  float w[T];
#pragma unroll
  for (int j = 0; j < T; ++j) w[j] = wp[j * WS];
#pragma unroll
  for (int c = 0; c < NF; c += CH) {
    cpx acc[CH / 2];
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const cpx ww = splat(w[j]);
#pragma unroll
      for (int qd = 0; qd < CH / 4; ++qd) {
        const float4 v = *reinterpret_cast<const float4*>(pk + j * kPStride + c + 4 * qd);
        if (j == 0) {
          acc[2 * qd] = cmul(ww, make_float2(v.x, v.y));
          acc[2 * qd + 1] = cmul(ww, make_float2(v.z, v.w));
        } else {
          acc[2 * qd] = cfma(ww, make_float2(v.x, v.y), acc[2 * qd]);
          acc[2 * qd + 1] = cfma(ww, make_float2(v.z, v.w), acc[2 * qd + 1]);
        }
      }
    }
    float v[CH];
#pragma unroll
    for (int i = 0; i < CH; i += 2) {
      const float la = fast_log2(fmaxf(acc[i >> 1].x, 1e-10f));
      const float lb = fast_log2(fmaxf(acc[i >> 1].y, 1e-10f));
      mx = max3(mx, la, lb);
      mn = min3(mn, la, lb);
      v[i] = fmaf(la, sc, of);
      v[i + 1] = fmaf(lb, sc, of);
    }
    if constexpr (CH == 8) {
      st_global_256(dst + c, v);
    } else {
      *reinterpret_cast<float4*>(dst + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// frames [lo, hi) as bits of the 16-frame tile starting at t0 (bit f = frame t0 + f)
__device__ __forceinline__ uint32_t frame_window(int lo, int hi, int t0) {
  lo = min(max(lo - t0, 0), kTileFrames);
  hi = min(max(hi - t0, 0), kTileFrames);
  return hi > lo ? (1u << hi) - (1u << lo) : 0u;
}

// tile descriptor (16 ints in sm_ctl, written by thread 0 in describe_tile, read by every phase that needs a field)
enum { kDescId = 0, kDescClip = 1, kDescT0 = 2, kDescKind = 3, kDescPcmLo = 4, kDescPcmHi = 5, kDescFlags = 6, kDescKeep = 7,
       kDescMask = 8 /* t0, t1, f0, f1 */, kDescOutLo = 12, kDescOutHi = 13 };
enum { kFlagFast = 1,        // the mel phase may take the branch-free path
       kFlagAllKept = 2,     // (fast tiles) every frame survives the partial-segment cut -> counts for the pad minimum
       kFlagAllMasked = 4 }; // (fast tiles) every frame lies inside the SpecAugment time mask

// ---- mel projection, edge tiles (generic): the clip's last tile, the two tiles a time mask or the partial-segment cut
// passes through, unaligned outputs -- three of a full clip's 188 tiles.  Same thread <-> (row, 8 frames) map, the same 16-byte
// loads of the power tile and the same summation order as mel_fast (tap 0 is a product, every later tap one FMA), so a cell's
// value does not depend on the path that produced it; what differs is that the tap count is a run-time value (one copy of the
// code for every warp and pass) and that every frame carries its own live / kept / stored / time-masked bit. ------------------
template <int NM>
__device__ __noinline__ void mel_edge(const float* __restrict__ sm_region, const float* __restrict__ sm_melw,
                                      const int* __restrict__ desc, uint32_t mel_desc, float* __restrict__ out, int n_frames,
                                      int n_frames_out, float mask_value, float* __restrict__ red) {
  const int warp = static_cast<int>(threadIdx.x) >> 5, lane = static_cast<int>(threadIdx.x) & 31;
  int passes = 1, T0 = 0, T1 = 0, wbase = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w)
    if (warp == w) {
      passes = mel_warp_passes<NM>(w);
      T0 = mel_warp_taps<NM>(w, 0);
      T1 = mel_warp_taps<NM>(w, 1);
      wbase = mel_warp_wbase<NM>(w);
    }
  const int f0 = (lane >> 4) * 8;
  const int clip = desc[kDescClip], t0 = desc[kDescT0], keep = desc[kDescKeep];
  // this thread's 8 frames: bit i = frame t0 + f0 + i
  const uint32_t live = (frame_window(0, n_frames, t0) >> f0) & 0xffu, kept = (frame_window(0, keep, t0) >> f0) & 0xffu,
                 store = (frame_window(0, n_frames < n_frames_out ? n_frames : n_frames_out, t0) >> f0) & 0xffu,
                 tmask = (frame_window(desc[kDescMask], desc[kDescMask + 1], t0) >> f0) & 0xffu;
  float mx = -INFINITY, mn_kept = INFINITY, mn_live = INFINITY;
#pragma unroll 1
  for (int ps = 0; ps < passes; ++ps) {
    const int T = ps ? T1 : T0;
    const int row = static_cast<int>((mel_desc >> (ps ? 15 : 0)) & 0x7fu);
    const float* pk = sm_region + ((mel_desc >> (ps ? 22 : 7)) & 0xffu) * kPStride + f0;
    const float* wp = sm_melw + wbase + (ps ? T0 * kMelSlots : 0) + (lane & (kMelSlots - 1));
    const bool rowmask = row >= desc[kDescMask + 2] && row < desc[kDescMask + 3];
    float* dst = out + (static_cast<size_t>(clip) * NM + row) * n_frames_out + t0 + f0;
    cpx acc[4];
    {
      const cpx ww = splat(wp[0]);
      const float4 va = *reinterpret_cast<const float4*>(pk), vb = *reinterpret_cast<const float4*>(pk + 4);
      acc[0] = cmul(ww, make_float2(va.x, va.y));
      acc[1] = cmul(ww, make_float2(va.z, va.w));
      acc[2] = cmul(ww, make_float2(vb.x, vb.y));
      acc[3] = cmul(ww, make_float2(vb.z, vb.w));
    }
#pragma unroll 1
    for (int j = 1; j < T; ++j) {
      const cpx ww = splat(wp[j * kMelSlots]);
      const float4 va = *reinterpret_cast<const float4*>(pk + j * kPStride), vb = *reinterpret_cast<const float4*>(pk + j * kPStride + 4);
      acc[0] = cfma(ww, make_float2(va.x, va.y), acc[0]);
      acc[1] = cfma(ww, make_float2(va.z, va.w), acc[1]);
      acc[2] = cfma(ww, make_float2(vb.x, vb.y), acc[2]);
      acc[3] = cfma(ww, make_float2(vb.z, vb.w), acc[3]);
    }
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = (i & 1) ? acc[i >> 1].y : acc[i >> 1].x;
      const float l2 = fast_log2(fmaxf(a, 1e-10f));
      if ((live >> i) & 1u) {
        mx = fmaxf(mx, l2);
        mn_live = fminf(mn_live, l2);
      }
      if ((kept >> i) & 1u) mn_kept = fminf(mn_kept, l2);
      v[i] = (rowmask || ((tmask >> i) & 1u)) ? mask_value : feature_of_l2(l2);
    }
    if (store == 0xffu && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      // two 16-byte stores: inside this predicated branch ptxas 12.9 turned the st.global.v8.f32 of st_global_256 into a single
      // 4-byte STG (seen in the SASS and on the device: only frame 0 of the clip's last tile was written)
      reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if ((store >> i) & 1u) dst[i] = v[i];
    }
  }
  mx = warp_max(mx);
  mn_kept = warp_min(mn_kept);
  mn_live = warp_min(mn_live);
  if (lane == 0) {
    red[3 * warp] = mx;
    red[3 * warp + 1] = mn_kept;
    red[3 * warp + 2] = mn_live;
  }
}

// ---- fix-up grid: what can only be finished once the whole clip is known ---------------------------------------------
// The front-end kernel wrote every computed cell as its final feature with the masks applied.  What may still be missing
// once the clip's max / min are complete: the floor max(feature, floor_feature(max)) where it binds, the min-value pad of
// the frames beyond the kept part (data/utils.py:380-404), and the constant rows of silent / pad-only tiles that were never
// computed.  `fixup_kernel` runs right behind the front-end grid: a warp owns 32 consecutive tiles at a time, lane <-> tile
// decides from the tile's recorded minimum whether anything is missing, and the warp rewrites the flagged tiles (in place,
// from L2).  Ordinary full-length audio flags nothing and the grid is a few microseconds of scanning.
struct FixupParams {
  float* out;
  const ClipStat* stats;
  const float* tile_min;
  const int32_t* lengths;
  const int32_t* n_valid;
  const int32_t* masks;
  int32_t n_samples, n_total, n_frames, n_frames_out, tiles_per_clip, total_tiles;
  float mask_value;
  int32_t draw, draw_tparam, draw_fparam;
  float draw_p;
  uint64_t draw_seed, draw_clip_offset;
};

__device__ __forceinline__ int kept_frames(const int32_t* n_valid, int clip, int n_frames) {
  int keep = n_frames;
  if (n_valid != nullptr) {
    const int nv = __ldg(n_valid + clip);
    if (nv >= 0 && nv < keep) keep = nv;
  }
  return keep;
}

constexpr int kFixThreads = 128;
constexpr int kFixWarps = kFixThreads / 32;
constexpr int kFixTiles = 32;       // lean instance: tiles per warp and pass of the scan, one per lane
constexpr int kFixTilesHeavy = 8;   // ragged batches rewrite a third or more of all tiles: smaller groups = four times the warps

// what a lane found out about ITS tile during the scan; broadcast to the warp when the tile is rewritten
struct FixTile {
  int clip, t0, keep;
  float floorn, padv;
  int mt0, mt1, mf0, mf1;
};

// one warp finishes one tile; `constant` = nothing was written yet (silent or pad-only tile): every kept cell is the clamp value
template <int NM, int kBatch, bool kHoist>
__device__ __forceinline__ void fixup_tile(const FixupParams& p, const FixTile& ft, bool constant, int lane) {
  const float vsilent = feature_of_l2(silent_l2());
  const float floorn = ft.floorn, padv = ft.padv, mv = p.mask_value;
  const int keep = ft.keep, t0 = ft.t0, mt0 = ft.mt0, mt1 = ft.mt1, mf0 = ft.mf0, mf1 = ft.mf1;
  const int pitch = p.n_frames_out;
  float* base = p.out + static_cast<size_t>(ft.clip) * NM * pitch;
  if (kHoist && (pitch & 3) == 0) {
    constexpr int kGroups = kTileFrames / 4;        // float4 groups per row (4)
    constexpr int kRowsPerPass = 32 / kGroups;      // 8 rows per warp-wide access
    // kBatch = rows in flight per lane (2 in the lean instance, which has to stay within 32 registers)
    static_assert(NM % (kRowsPerPass * kBatch) == 0, "row loop");
    const int f = t0 + ((lane & (kGroups - 1)) << 2);
    if (f >= pitch) return;
    // A lane's four columns are the same for every row: which of them are pad / inside the time mask is decided once, and so
    // is the whole value of a cell that needs no load (a tile that was never computed, columns beyond the kept frames) --
    // ncu on a ragged batch: the per-cell selects were 770 warp-instructions per rewritten tile, 19 M per launch at IPC 0.9.
    uint32_t padbits = 0, maskbits = 0;
    float cv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int fc = f + c;
      if (fc >= keep) padbits |= 1u << c;
      if (fc >= mt0 && fc < mt1) maskbits |= 1u << c;
      cv[c] = (fc >= mt0 && fc < mt1) ? mv : (fc >= keep ? padv : fmaxf(vsilent, floorn));
    }
    const float4 cv4 = make_float4(cv[0], cv[1], cv[2], cv[3]), mv4 = make_float4(mv, mv, mv, mv);
    const bool load = !constant && f < keep;
#pragma unroll 1
    for (int r0 = lane / kGroups; r0 < NM; r0 += kRowsPerPass * kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int it = 0; it < kBatch; ++it) {
        const int row = r0 + kRowsPerPass * it;
        const bool rowmask = row >= mf0 && row < mf1;
        if (load && !rowmask) v[it] = __ldcg(reinterpret_cast<const float4*>(base + static_cast<size_t>(row) * pitch + f));
      }
#pragma unroll
      for (int it = 0; it < kBatch; ++it) {
        const int row = r0 + kRowsPerPass * it;
        const bool rowmask = row >= mf0 && row < mf1;
        float4 o = rowmask ? mv4 : cv4;
        if (load && !rowmask) {
          float e[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) e[c] = ((maskbits >> c) & 1u) ? mv : (((padbits >> c) & 1u) ? padv : fmaxf(e[c], floorn));
          o = make_float4(e[0], e[1], e[2], e[3]);
        }
        *reinterpret_cast<float4*>(base + static_cast<size_t>(row) * pitch + f) = o;
      }
    }
  } else if ((pitch & 3) == 0) {   // the lean instance (32 registers): everything per cell
    constexpr int kGroups = kTileFrames / 4;        // float4 groups per row (4)
    constexpr int kRowsPerPass = 32 / kGroups;      // 8 rows per warp-wide access
    // kBatch = rows in flight per lane (2 in the lean instance, which has to stay within 32 registers)
    static_assert(NM % (kRowsPerPass * kBatch) == 0, "row loop");
    const int f = t0 + ((lane & (kGroups - 1)) << 2);
    if (f >= pitch) return;
    const bool load = !constant && f < keep;
#pragma unroll 1
    for (int r0 = lane / kGroups; r0 < NM; r0 += kRowsPerPass * kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int it = 0; it < kBatch; ++it) {
        const int row = r0 + kRowsPerPass * it;
        v[it] = make_float4(vsilent, vsilent, vsilent, vsilent);
        if (load && row < NM) v[it] = __ldcg(reinterpret_cast<const float4*>(base + static_cast<size_t>(row) * pitch + f));
      }
#pragma unroll
      for (int it = 0; it < kBatch; ++it) {
        const int row = r0 + kRowsPerPass * it;
        if (row < NM) {
          const bool rowmask = row >= mf0 && row < mf1;
          float e[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int fc = f + c;
            float r = (fc < keep) ? fmaxf(e[c], floorn) : padv;
            if (rowmask || (fc >= mt0 && fc < mt1)) r = mv;
            e[c] = r;
          }
          *reinterpret_cast<float4*>(base + static_cast<size_t>(row) * pitch + f) = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
    }
  } else {
    for (int idx = lane; idx < NM * kTileFrames; idx += 32) {
      const int row = idx / kTileFrames;
      const int f = t0 + (idx % kTileFrames);
      if (f >= pitch) continue;
      float* ptr = base + static_cast<size_t>(row) * pitch + f;
      float r = padv;
      if (f < keep) r = fmaxf(constant ? vsilent : __ldcg(ptr), floorn);
      if ((row >= mf0 && row < mf1) || (f >= mt0 && f < mt1)) r = mv;
      *ptr = r;
    }
  }
}

__device__ __forceinline__ bool tile_is_silent(int t0, int len, int n_total);

// kLean: the instance launched behind batches that hardly ever need a fix-up (full-length clips) -- <= 32 registers, one CTA
// per SM, so that it fits next to six front-end CTAs.  The other instance serves ragged batches, where a third or more of
// all tiles are constant fills: more registers, more CTAs.
template <int NM, bool kLean>
__global__ void __launch_bounds__(kFixThreads, kLean ? 16 : 4) fixup_kernel(const FixupParams p) {
  // the kernel behind this one may be scheduled now (it decides itself what it has to wait for); this grid needs the
  // front-end grid complete: its features, the clip statistics and the per-tile minima
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = static_cast<int>(threadIdx.x) & 31, warp = static_cast<int>(threadIdx.x) >> 5;
  // a WARP owns groups of 32 consecutive tiles (lane <-> tile), grid-stride: no shared memory, no barrier.  Every lane
  // gathers everything its own tile's rewrite needs (clip statistics, kept frames, mask intervals) -- 32 tiles' worth of
  // dependent loads in parallel -- and the warp then rewrites the flagged tiles one after the other from broadcast values.
  constexpr int kGroup = kLean ? kFixTiles : kFixTilesHeavy;   // tiles per warp and pass (lanes beyond the group idle in the scan)
  for (int base = (static_cast<int>(blockIdx.x) * kFixWarps + warp) * kGroup; base < p.total_tiles;
       base += static_cast<int>(gridDim.x) * kFixWarps * kGroup) {
    const int tile = base + lane;
    bool needs = false, constant = false;
    FixTile ft{};
    if (lane < kGroup && tile < p.total_tiles) {
      ft.clip = tile / p.tiles_per_clip;
      ft.t0 = (tile - ft.clip * p.tiles_per_clip) * kTileFrames;
      if (ft.t0 < p.n_frames_out) {
        ft.keep = kept_frames(p.n_valid, ft.clip, p.n_frames);
        const uint32_t max_enc = __ldcg(&p.stats[ft.clip].max_enc);
        ft.floorn = floor_feature(dec_ordered(max_enc));
        if (ft.t0 >= p.n_frames) {            // pad-only tile (n_frames_out > n_frames)
          needs = constant = true;
        } else {
          int len = p.n_samples;
          if (p.lengths != nullptr) {
            const int l = __ldg(p.lengths + ft.clip);
            len = l < 0 ? 0 : (l < len ? l : len);
          }
          if (tile_is_silent(ft.t0, len, p.n_total)) {
            needs = constant = true;
          } else {
            const int hi = ft.t0 + kTileFrames < p.n_frames_out ? ft.t0 + kTileFrames : p.n_frames_out;
            const bool has_pad = (ft.t0 > ft.keep ? ft.t0 : ft.keep) < hi;
            const bool floor_binds = feature_of_l2(__ldcg(p.tile_min + tile)) < ft.floorn;
            needs = has_pad || floor_binds;
          }
        }
        if (needs) {
          ft.padv = fmaxf(feature_of_l2(dec_ordered(~__ldcg(&p.stats[ft.clip].min_inv))), ft.floorn);
          if (p.masks != nullptr || p.draw) {
            const int4 mk = p.masks != nullptr ? __ldg(reinterpret_cast<const int4*>(p.masks) + ft.clip)
                                               : draw_mask_intervals(p.draw_seed, p.draw_clip_offset + static_cast<uint64_t>(ft.clip),
                                                                     NM, p.n_frames_out, p.draw_tparam, p.draw_fparam, p.draw_p);
            ft.mt0 = mk.x; ft.mt1 = mk.y; ft.mf0 = mk.z; ft.mf1 = mk.w;
          }
        }
      }
    }
    uint32_t todo = __ballot_sync(0xffffffffu, needs);
    const uint32_t cst = __ballot_sync(0xffffffffu, constant);
    while (todo != 0u) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1u;
      FixTile b;
      b.clip = __shfl_sync(0xffffffffu, ft.clip, l); b.t0 = __shfl_sync(0xffffffffu, ft.t0, l);
      b.keep = __shfl_sync(0xffffffffu, ft.keep, l);
      b.floorn = __shfl_sync(0xffffffffu, ft.floorn, l); b.padv = __shfl_sync(0xffffffffu, ft.padv, l);
      b.mt0 = __shfl_sync(0xffffffffu, ft.mt0, l); b.mt1 = __shfl_sync(0xffffffffu, ft.mt1, l);
      b.mf0 = __shfl_sync(0xffffffffu, ft.mf0, l); b.mf1 = __shfl_sync(0xffffffffu, ft.mf1, l);
      fixup_tile<NM, 2, !kLean>(p, b, ((cst >> l) & 1u) != 0u, lane);   // (8 rows in flight in the heavy instance: 113 registers, same time)
    }
  }
}

// sm_ctl slots
// [kCtlDesc .. +15] and [kCtlDesc + kCtlSlot .. +15] = two tile descriptors (kDesc*): the tile being worked on and the one
// after it, swapping roles every iteration.  Every phase re-reads the few fields it needs from here instead of carrying
// them in registers across the FFT stages.
enum { kCtlDesc = 0, kCtlSlot = 16,
       kCtlRed = 32,    // 3 floats per warp (max, kept min, live min)
       kCtlMbar = 48,   // 8 bytes
       kCtlMemo = 52,   // describe_tile's per-clip memo: clip, len, keep, -, mask[4]
       kCtlChunkNext = 60,   // next tile of the chunk this CTA claimed, and how many of its tiles are still to be described
       kCtlChunkLeft = 61 };
static_assert(kCtlChunkLeft < kCtlInts && kCtlRed + 3 * kWarps <= kCtlMbar, "sm_ctl layout");

// how a tile's PCM reaches shared memory
enum { kTileEdge = 0,      // reflection / zero extension / unaligned source: scalar staging
       kTileInterior = 1,  // plain in-range samples: TMA bulk copies, one tile ahead
       kTileSilent = 2,    // every sample it touches is zero padding: no FFT, the fix-up later writes the constant rows;
                           // this is the FIRST such tile of its clip and records the clamp value in the clip statistics
       kTileSilentRest = 3 };  // a later tile of the same silent tail: the statistics are already in, only counted
__device__ __forceinline__ bool tile_is_silent(int t0, int len, int n_total) {
  const int g0 = t0 * kHop - kNfft / 2;
  const int last = g0 + kTileSamples - 1;  // beyond n_total the samples are mirrored around n_total - 1
  return len == 0 || (g0 >= len && (last < n_total || 2 * (n_total - 1) - last >= len));
}

struct TileGeom {   // the few launch constants describe_tile needs
  const void* pcm;
  const int32_t* lengths;
  const int32_t* n_valid;
  const int32_t* masks;
  int64_t clip_stride;
  int32_t n_samples, n_total, n_frames, n_frames_out, tiles_per_clip, total_tiles;
  uint32_t tpc_magic;   // floor(2^32 / tiles_per_clip)
  int32_t vec_ok;
  int32_t draw, draw_tparam, draw_fparam;
  float draw_p;
  uint64_t draw_seed, draw_clip_offset;
};
__device__ __forceinline__ TileGeom tile_geom(const FrontendParams& q) {
  TileGeom g;
  g.pcm = q.pcm; g.lengths = q.lengths; g.n_valid = q.n_valid; g.masks = q.masks; g.clip_stride = q.clip_stride;
  g.n_samples = q.n_samples; g.n_total = q.n_total; g.n_frames = q.n_frames; g.n_frames_out = q.n_frames_out;
  g.tiles_per_clip = q.tiles_per_clip; g.total_tiles = q.total_tiles; g.tpc_magic = q.tpc_magic; g.vec_ok = q.vec_ok;
  g.draw = q.draw; g.draw_tparam = q.draw_tparam; g.draw_fparam = q.draw_fparam; g.draw_p = q.draw_p;
  g.draw_seed = q.draw_seed; g.draw_clip_offset = q.draw_clip_offset;
  return g;
}

// thread 0: describe tile `t` for everybody (one multiply-high instead of 160 divisions; lengths[], n_valid[] and the
// mask intervals are loaded once per clip and memoised in shared memory -- consecutive tiles mostly share the clip)
// (forced inline: as a real call this sat on thread 0's critical path before a CTA barrier and cost 2.5 % overall)
template <int NM, typename PcmT>
__device__ __forceinline__ void describe_tile(const TileGeom p, int t, int* __restrict__ slot, int* __restrict__ memo,
                                              const int4* __restrict__ draw_cache) {
  int clip = 0, t0 = 0, kind = kTileEdge, flags = 0, keep = 0;
  long long off = 0, out_off = 0;
  int4 mk = make_int4(0, 0, 0, 0);
  if (t < p.total_tiles) {
    clip = static_cast<int>(__umulhi(static_cast<uint32_t>(t), p.tpc_magic));
    if (t - clip * p.tiles_per_clip >= p.tiles_per_clip) ++clip;   // the magic number undershoots by at most one
    t0 = (t - clip * p.tiles_per_clip) * kTileFrames;
    if (clip != memo[0]) {
      int len = p.n_samples;
      if (p.lengths != nullptr) {
        const int l = __ldg(p.lengths + clip);
        len = l < 0 ? 0 : (l < len ? l : len);
      }
      int kp = p.n_frames;
      if (p.n_valid != nullptr) {
        const int nv = __ldg(p.n_valid + clip);
        if (nv >= 0 && nv < kp) kp = nv;
      }
      int4 m = make_int4(0, 0, 0, 0);
      if (p.masks != nullptr) m = __ldg(reinterpret_cast<const int4*>(p.masks) + clip);
      else if (p.draw) m = draw_cache != nullptr ? draw_cache[clip]
                                                 : draw_mask_intervals(p.draw_seed, p.draw_clip_offset + static_cast<uint64_t>(clip), NM,
                                                                       p.n_frames_out, p.draw_tparam, p.draw_fparam, p.draw_p);
      memo[0] = clip; memo[1] = len; memo[2] = kp;
      *reinterpret_cast<int4*>(memo + 4) = m;
    }
    keep = memo[2];
    mk = *reinterpret_cast<const int4*>(memo + 4);
    out_off = static_cast<long long>(clip) * NM * p.n_frames_out + t0;
    if (t0 < p.n_frames) {
      const int len = memo[1];
      const int g0 = t0 * kHop - kNfft / 2;
      off = static_cast<long long>(clip) * p.clip_stride + g0;
      if (tile_is_silent(t0, len, p.n_total)) {
        // silence is a suffix of the clip: only its first tile does any bookkeeping
        kind = (t0 == 0 || !tile_is_silent(t0 - kTileFrames, len, p.n_total)) ? kTileSilent : kTileSilentRest;
      } else if (tile_is_interior(reinterpret_cast<const PcmT*>(p.pcm) + off - g0, g0, len)) {
        kind = kTileInterior;
      }
      const int t1 = t0 + kTileFrames;
      const bool all_kept = t1 <= keep;
      const bool no_tmask = mk.y <= mk.x || mk.y <= t0 || mk.x >= t1;
      const bool all_tmask = mk.x <= t0 && mk.y >= t1;
      if (p.vec_ok && (all_kept || keep <= t0) && t1 <= p.n_frames && t1 <= p.n_frames_out && (no_tmask || all_tmask))
        flags = kFlagFast | (all_kept ? kFlagAllKept : 0) | (all_tmask ? kFlagAllMasked : 0);
    }
  }
  *reinterpret_cast<int4*>(slot) = make_int4(t, clip, t0, kind);
  *reinterpret_cast<int4*>(slot + 4) = make_int4(static_cast<int>(off & 0xffffffffll), static_cast<int>(off >> 32), flags, keep);
  *reinterpret_cast<int4*>(slot + 8) = mk;
  *reinterpret_cast<int2*>(slot + 12) = make_int2(static_cast<int>(out_off & 0xffffffffll), static_cast<int>(out_off >> 32));
}

// development build only (-DWFT_TIMELINE): lane 0 of every warp stamps %clock at 13 points of its first kTlIters tile
// iterations; tools/timeline.cu turns the stamps into per-phase work / barrier-wait times
#ifdef WFT_TIMELINE
constexpr int kTlIters = 32, kTlPoints = 16, kTlMaxCtas = 148 * 6;
__device__ uint32_t g_timeline[kTlMaxCtas * kTlIters * kWarps * kTlPoints];
#define WFT_TL(k)                                                                                                      \
  do {                                                                                                                 \
    if (lane == 0 && tl_iter < kTlIters && blockIdx.x < kTlMaxCtas) {                                                  \
      uint32_t c_;                                                                                                     \
      asm volatile("mov.u32 %0, %%clock;" : "=r"(c_));                                                                 \
      g_timeline[((blockIdx.x * kTlIters + tl_iter) * kWarps + warp) * kTlPoints + (k)] = c_;                          \
    }                                                                                                                  \
  } while (0)
#else
#define WFT_TL(k) do {} while (0)
#endif

template <int NM, typename PcmT>
__global__ void __launch_bounds__(kThreads, 6) frontend_kernel(const FrontendParams p) {
  extern __shared__ __align__(16) float smem[];
  float* sm_region = smem;
  PcmT* sm_audio = reinterpret_cast<PcmT*>(smem + kAudioBase);
  float* sm_win = smem + kRegionFloats;
  float* sm_tw = sm_win + kWinFloats;
  float* sm_melw = sm_tw + kTwFloats;
  int* sm_ctl = reinterpret_cast<int*>(sm_melw + kMelWFloats);
  int4* sm_draw = reinterpret_cast<int4*>(sm_ctl + kCtlInts);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  {
    // int16 PCM: the 1/32768 scale (whisper.audio.load_audio) is folded into the window, exactly (power of two)
    const float wscale = sizeof(PcmT) == 2 ? (1.0f / 32768.0f) : 1.0f;
    // 7.4 KB of tables as 16-byte vectors, every load of a thread in flight before its first store (this prologue is
    // ~1/40 of a B = 64 launch when it is a chain of dependent 4-byte round trips)
    constexpr int kMelW = NM == 128 ? WFT_MEL128_W_LEN : WFT_MEL80_W_LEN;
    constexpr int kVecWin = kWinFloats / 4, kVecTw = kTwFloats / 4, kVecMel = kMelW / 4;
    constexpr int kVecAll = kVecWin + kVecTw + kVecMel;
    constexpr int kPerThread = (kVecAll + kThreads - 1) / kThreads;
    const float4* gwin = reinterpret_cast<const float4*>(g_window_table);
    const float4* gtw = reinterpret_cast<const float4*>(g_twiddle_table);
    const float4* gmel = reinterpret_cast<const float4*>(NM == 128 ? g_mel128_w : g_mel80_w);
    float4 v[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int k = tid + i * kThreads;
      if (k < kVecWin) v[i] = __ldg(gwin + k);
      else if (k < kVecWin + kVecTw) v[i] = __ldg(gtw + (k - kVecWin));
      else if (k < kVecAll) v[i] = __ldg(gmel + (k - kVecWin - kVecTw));
    }
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int k = tid + i * kThreads;
      if (k < kVecWin) v[i] = make_float4(v[i].x * wscale, v[i].y * wscale, v[i].z * wscale, v[i].w * wscale);
      // sm_win, sm_tw and sm_melw are contiguous in that order
      if (k < kVecAll) reinterpret_cast<float4*>(sm_win)[k] = v[i];
    }
  }
  // Intervals drawn in here (draw_masks): a small batch is drawn once per CTA, one clip per thread, into shared memory -- with
  // 13 tiles per CTA nearly every tile belongs to another clip, and ~150 instructions of Philox on thread 0 per tile held up
  // every barrier behind it (3 % of the launch).  Larger batches draw on demand (consecutive tiles mostly share the clip).
  const int4* draw_cache = (p.draw && p.batch <= kDrawCache) ? sm_draw : nullptr;
  if (draw_cache != nullptr && tid < p.batch)
    sm_draw[tid] = draw_mask_intervals(p.draw_seed, p.draw_clip_offset + static_cast<uint64_t>(tid), NM, p.n_frames_out,
                                       p.draw_tparam, p.draw_fparam, p.draw_p);
  uint64_t* audio_bar = reinterpret_cast<uint64_t*>(sm_ctl + kCtlMbar);  // completion of the audio tile's bulk copies
  // Programmatic dependent launch: everything above (7.4 KB of tables into shared memory) may overlap the tail of the
  // previous kernel on the stream; its results and the workspace may only be touched from here on.
  // The grid behind this one on the stream (the fix-up grid) may be scheduled as soon as SM slots free up.  An independent
  // launch (p.overlap) never looks at what the grids in front of it produced and starts working at once.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!p.overlap) asm volatile("griddepcontrol.wait;" ::: "memory");
  // self-cleaning modes: the other phase's counters belong to the call BEFORE this one (complete: see above) and to the
  // one AFTER it (not started: it waits for this grid): zeroing them here needs no fence and replaces a memset per call
  if (blockIdx.x == 0)
    for (int i = tid; i < p.clean_vec; i += kThreads) p.clean[i] = make_uint4(0u, 0u, 0u, 0u);
  if (draw_cache != nullptr) __syncthreads();   // the first describe_tile below reads the drawn intervals
  if (tid == 0) {
    sm_ctl[kCtlMemo] = -1;
    mbar_init(audio_bar, 1);   // one arrive.expect_tx per tile (prefetch_audio_tile)
    const int first = static_cast<int>(atomicAdd(p.tile_counter, static_cast<uint32_t>(p.chunk)));
    sm_ctl[kCtlChunkNext] = first + 1;
    sm_ctl[kCtlChunkLeft] = p.chunk - 1;
    describe_tile<NM, PcmT>(tile_geom(p), first, sm_ctl + kCtlDesc, sm_ctl + kCtlMemo, draw_cache);
  }
  __syncthreads();
  // a tile of kind kTileInterior always arrives by TMA: the first one is sent here, every later one under the tile before it
  if (warp == kPrefetchWarp && sm_ctl[kCtlDesc + kDescKind] == kTileInterior) {
    const long long off = (static_cast<long long>(sm_ctl[kCtlDesc + kDescPcmHi]) << 32) | static_cast<unsigned int>(sm_ctl[kCtlDesc + kDescPcmLo]);
    prefetch_audio_tile<PcmT>(sm_audio, reinterpret_cast<const PcmT*>(p.pcm) + off, audio_bar, lane);
  }
  // loop state in ONE register: bit 4 (kCtlSlot) = descriptor slot of the CURRENT tile, bit 0 = parity of the audio mbarrier
  int lstate = 0;

  // mel phase role of this thread (fixed for the whole launch): row | start_bin << 8 | first_frame << 16 | weight_offset << 20
  // (kept packed in ONE register across the FFT stages; unpacked again in every mel phase)
  const uint32_t mel_desc = (NM == 128 ? g_mel128_thread : g_mel80_thread)[tid];

#ifdef WFT_TIMELINE
  int tl_iter = -1;
#endif
  for (;;) {
#ifdef WFT_TIMELINE
    ++tl_iter;
#endif
    // current tile at DESC; thread 0 describes the next tile at NDESC (behind this tile's first barrier)
#define DESC (sm_ctl + (launder(lstate) & kCtlSlot))
#define NDESC (sm_ctl + ((launder(lstate) & kCtlSlot) ^ kCtlSlot))
#define DESC4 (*reinterpret_cast<const int4*>(DESC))
    // thread 0: the tile after the current one = the next tile of the claimed chunk, or the first tile of the chunk just claimed
#define DESCRIBE_NEXT()                                                                              \
  do {                                                                                               \
    const int left_ = sm_ctl[kCtlChunkLeft];                                                         \
    const int t_ = left_ == 0 ? nxt_claim : sm_ctl[kCtlChunkNext];                                   \
    sm_ctl[kCtlChunkLeft] = left_ == 0 ? p.chunk - 1 : left_ - 1;                                    \
    sm_ctl[kCtlChunkNext] = t_ + 1;                                                                  \
    describe_tile<NM, PcmT>(tile_geom(p), t_, NDESC, sm_ctl + kCtlMemo, draw_cache);                 \
  } while (0)
    const int4 d_top = DESC4;
    if (d_top.x >= p.total_tiles) break;
    WFT_TL(0);
    // claim the tile AFTER this one now; the answer is consumed two barriers later (latency hidden by stage A)
    int nxt_claim = 0;
    if (tid == 0 && sm_ctl[kCtlChunkLeft] == 0) nxt_claim = static_cast<int>(atomicAdd(p.tile_counter, static_cast<uint32_t>(p.chunk)));

    if (d_top.z < p.n_frames && d_top.w != kTileSilent && d_top.w != kTileSilentRest) {
      // stage 0 ---------------------------------------------------------------------------------------------
      if (d_top.w != kTileInterior) {   // clip edge, ragged end or unaligned source: scalar staging, now
        const int clip = d_top.y, t0 = d_top.z;
        const PcmT* x = reinterpret_cast<const PcmT*>(p.pcm) + static_cast<size_t>(clip) * p.clip_stride;
        int len = p.n_samples;
        if (p.lengths != nullptr) {
          const int l = __ldg(p.lengths + clip);
          len = l < 0 ? 0 : (l < len ? l : len);
        }
        stage_audio_edge<PcmT>(sm_audio, x, t0 * kHop - kNfft / 2, len, p.n_total, tid);
        __syncthreads();
      } else {                          // the tile was sent by TMA (normally one tile ago): wait for its bytes
        mbar_wait(audio_bar, static_cast<uint32_t>(lstate) & 1u);
        lstate ^= 1;
      }
      WFT_TL(1);

      // stage A: thread (q, n2 = r): x[n1] = w[20 n1 + n2] * (pa + i pb)[20 n1 + n2] ------------------------------
      {
        cpx x[20];
        {
          float u[28];
          const auto [q, r] = pair_coord_a();
          const PcmT* a = sm_audio + (kSkewBlock + Skew<PcmT>::value) * q + r;
#pragma unroll
          for (int j = 0; j < 28; ++j) u[j] = pcm_as_float(a[20 * j + (j >= 16 ? Skew<PcmT>::value : 0)]);
          const float4* w4 = reinterpret_cast<const float4*>(sm_win + r * 20);
#pragma unroll
          for (int a4 = 0; a4 < 5; ++a4) {
            const float4 w = w4[a4];
            const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int n1 = 4 * a4 + e;
              x[n1] = make_float2(ww[e] * u[n1], ww[e] * u[n1 + 8]);
            }
          }
        }
        WFT_TL(2);
        __syncthreads();  // audio is dead from here on: the region becomes the exchange buffer
        WFT_TL(3);
        dft20(x);
        const auto [q, r] = pair_coord_a();
        const float4* t4 = reinterpret_cast<const float4*>(sm_tw + r * kTwRow);
        float2* e2 = reinterpret_cast<float2*>(sm_region + q * kPairStride) + r;
#pragma unroll
        for (int h = 0; h < 10; ++h) {
          const float4 t = t4[h];
          const int k0 = 2 * h, k1 = 2 * h + 1;
          e2[k0 * (kRowStride / 2)] = make_float2(x[k0].x * t.x - x[k0].y * t.y, fmaf(x[k0].x, t.y, x[k0].y * t.x));
          e2[k1 * (kRowStride / 2)] = make_float2(x[k1].x * t.z - x[k1].y * t.w, fmaf(x[k1].x, t.w, x[k1].y * t.z));
        }
      }
      WFT_TL(15);
      if (tid == 0) DESCRIBE_NEXT();
      WFT_TL(4);
      __syncthreads();
      WFT_TL(5);

      // stage B: thread (q, k1 = r): Z[k1 + 20 k2] = DFT20 over n2.  The lower half (k2 < 10, bins k1 + 20 k2 <= 199)
      // stays in registers; the upper half is what the mirror thread (q, 20 - k1) needs and travels by warp shuffle.
      // power: thread (q, j = r): bins j + 20 m (m = 0..9) against their mirrors Z[400 - j - 20 m] = row (20-j)%20,
      // position 19 - m = mz[9 - m].  Row 0 mirrors ITSELF one position further (20 - m, with 20 == 0): it reads its own
      // upper half back from the shuffle and picks mz[10 - m] (m > 0) or z[0]; only warp 4 holds row 0 and pays for the
      // selection.
      {
        cpx z[10], mz[10];
        {
          cpx y[20];
          {
            const auto [q, r] = pair_coord_b();
            const float4* row4 = reinterpret_cast<const float4*>(sm_region + q * kPairStride + r * kRowStride);
#pragma unroll
            for (int a = 0; a < 10; ++a) {
              const float4 v = row4[a];
              y[2 * a] = make_float2(v.x, v.y);
              y[2 * a + 1] = make_float2(v.z, v.w);
            }
          }
          dft20(y);
#pragma unroll
          for (int m = 0; m < 10; ++m) z[m] = y[m];
          const int src = mirror_lane();
#pragma unroll
          for (int m = 0; m < 10; ++m)
            mz[m] = make_float2(__shfl_sync(0xffffffffu, y[10 + m].x, src), __shfl_sync(0xffffffffu, y[10 + m].y, src));
        }
        WFT_TL(6);
        __syncthreads();  // exchange is dead: the region becomes power tile (bottom) + next audio tile (top)
        WFT_TL(7);
        // prefetch the NEXT tile's PCM into the top of the region (its descriptor stays in sm_ctl until the loop ends)
        if (warp == kPrefetchWarp && NDESC[kDescKind] == kTileInterior) {
          const long long off = (static_cast<long long>(NDESC[kDescPcmHi]) << 32) | static_cast<unsigned int>(NDESC[kDescPcmLo]);
          WFT_TL(13);
          prefetch_audio_tile<PcmT>(sm_audio, reinterpret_cast<const PcmT*>(p.pcm) + off, audio_bar, lane);
          WFT_TL(14);
        }

        // power tile [bin][frame]: the pair's two frames are neighbours, one 8-byte store per bin
        const auto [q, r] = pair_coord_b();
        float* pw = sm_region + r * kPStride + 2 * q;
        if (tid < 128) {
#pragma unroll
          for (int m = 0; m < 10; ++m) {
            // Z[k] = z[m], Z[400-k] = mz[9-m]
            const cpx sa = cfma(mz[9 - m], make_float2(1.0f, -1.0f), z[m]);   // Z[k] + conj Z[400-k]  -> frame 2q
            const cpx sb = cfma(mz[9 - m], make_float2(-1.0f, 1.0f), z[m]);   // Z[k] - conj Z[400-k]  -> frame 2q + 1
            *reinterpret_cast<float2*>(pw + 20 * m * kPStride) =
                make_float2(fmaf(sa.x, sa.x, sa.y * sa.y), fmaf(sb.x, sb.x, sb.y * sb.y));
          }
        } else {   // warp 4: rows 0, 10, 9, 11
          const bool row0 = r == 0;
#pragma unroll
          for (int m = 0; m < 10; ++m) {
            const cpx alt = m == 0 ? z[0] : mz[10 - m];
            const cpx mir = make_float2(row0 ? alt.x : mz[9 - m].x, row0 ? alt.y : mz[9 - m].y);
            const cpx sa = cfma(mir, make_float2(1.0f, -1.0f), z[m]);
            const cpx sb = cfma(mir, make_float2(-1.0f, 1.0f), z[m]);
            *reinterpret_cast<float2*>(pw + 20 * m * kPStride) =
                make_float2(fmaf(sa.x, sa.x, sa.y * sa.y), fmaf(sb.x, sb.x, sb.y * sb.y));
          }
        }
      }
      WFT_TL(8);
      __syncthreads();
      WFT_TL(9);

      // mel phase: thread <-> (mel row, NF frames) -------------------------------------------------------------
      {
        const int* desc = DESC;
        const int flags = desc[kDescFlags];
        float* red = reinterpret_cast<float*>(sm_ctl + kCtlRed);
        if (flags & kFlagFast) {
          const int4 mk = *reinterpret_cast<const int4*>(desc + kDescMask);
          const int2 oo = *reinterpret_cast<const int2*>(desc + kDescOutLo);
          const long long out_off = (static_cast<long long>(oo.y) << 32) | static_cast<unsigned int>(oo.x);
          const int half8 = (lane >> 4) * 8;
          float mx = -INFINITY, mn = INFINITY;
          // one pass = this thread's row x 8 frames; warp-uniform dispatch, one fully unrolled body per (warp, pass) of the plan
#define MEL_PASS(W, PS)                                                                                              \
  do {                                                                                                               \
    constexpr int T_ = mel_warp_taps<NM>(W, PS);                                                                     \
    const int row_ = static_cast<int>((mel_desc >> ((PS) ? 15 : 0)) & 0x7fu);                                        \
    const float* pk_ = sm_region + ((mel_desc >> ((PS) ? 22 : 7)) & 0xffu) * kPStride + half8;                       \
    const float* wp_ = sm_melw + mel_warp_wbase<NM>(W) + ((PS) ? mel_warp_taps<NM>(W, 0) * kMelSlots : 0) + (lane & (kMelSlots - 1)); \
    const bool masked_ = (row_ >= mk.z && row_ < mk.w) || (flags & kFlagAllMasked) != 0;                             \
    const float sc_ = masked_ ? 0.0f : kFeatScale;            /* masked cell: 0 * L2 + mask_value */                  \
    const float of_ = masked_ ? p.mask_value : 1.0f;                                                                 \
    mel_fast<T_, 8, kMelSlots>(pk_, wp_, sc_, of_, p.out + out_off + (row_ * p.n_frames_out + half8), mx, mn);       \
  } while (0)
#define MEL_WARP(W)                                           \
  do {                                                        \
    MEL_PASS(W, 0);                                           \
    if constexpr (mel_warp_passes<NM>(W) > 1) MEL_PASS(W, 1); \
  } while (0)
          if (warp == 0) MEL_WARP(0);
          else if (warp == 1) MEL_WARP(1);
          else if (warp == 2) MEL_WARP(2);
          else if (warp == 3) MEL_WARP(3);
          else MEL_WARP(4);
#undef MEL_WARP
#undef MEL_PASS
          mx = warp_max(mx);
          mn = warp_min(mn);
          if (lane == 0) {
            red[3 * warp] = mx;
            red[3 * warp + 1] = (flags & kFlagAllKept) ? mn : INFINITY;
            red[3 * warp + 2] = mn;
          }
        } else {
          mel_edge<NM>(sm_region, sm_melw, desc, mel_desc, p.out, p.n_frames, p.n_frames_out, p.mask_value, red);
        }
      }
    } else {
      // nothing to compute: a pad-only tile (n_frames_out > n_frames) or a silent tile (all-zero PCM: every mel value is
      // the 1e-10 clamp, so only its statistics are recorded here and the fix-up later writes the constant rows)
      __syncthreads();  // the previous tile's last readers of the other descriptor slot are done
      if (tid == 0) DESCRIBE_NEXT();
      __syncthreads();
      if (warp == kPrefetchWarp && NDESC[kDescKind] == kTileInterior) {
        const long long off = (static_cast<long long>(NDESC[kDescPcmHi]) << 32) | static_cast<unsigned int>(NDESC[kDescPcmLo]);
        prefetch_audio_tile<PcmT>(sm_audio, reinterpret_cast<const PcmT*>(p.pcm) + off, audio_bar, lane);
      }
      if (lane == 0) {
        const int clip = d_top.y, t0 = d_top.z, kind = d_top.w;   // (short path: the loop-top read is still in registers)
        const bool silent = (kind == kTileSilent || kind == kTileSilentRest) && t0 < p.n_frames;
        const bool silent_head = kind == kTileSilent;   // only the first silent tile of a clip touches the clip statistics
        float* red = reinterpret_cast<float*>(sm_ctl + kCtlRed) + 3 * warp;
        const float lc = silent_l2();
        red[0] = (silent && silent_head) ? lc : -INFINITY;                                                 // live frames
        red[1] = (silent && silent_head && t0 < kept_frames(p.n_valid, clip, p.n_frames)) ? lc : INFINITY;  // kept frames
        red[2] = silent ? lc : INFINITY;
      }
    }

    WFT_TL(10);
    __syncthreads();  // tile finished: power tile free, per-warp max / min visible
    WFT_TL(11);

    // publish the tile's statistics, fire and forget: lanes 0 / 1 of warp 0 reduce the clip's max / kept minimum into the
    // workspace (red.max on ordered-int encodings), lane 2 records the tile's own minimum for the fix-up grid.  Nobody in
    // this grid reads them back; the kernel boundary makes them visible to fixup_kernel.
    if (warp == 0) {
      const float* red = reinterpret_cast<const float*>(sm_ctl + kCtlRed) + 3 * (lane < kWarps ? lane : 0);
      float mx = lane < kWarps ? red[0] : -INFINITY;
      float mn_kept = lane < kWarps ? red[1] : INFINITY;
      float mn_live = lane < kWarps ? red[2] : INFINITY;
      mx = warp_max(mx);
      mn_kept = warp_min(mn_kept);
      mn_live = warp_min(mn_live);
      if (lane < 3) {
        const int4 d_pub = DESC4;
        ClipStat* cs = p.stats + d_pub.y;
        if (lane == 0 && mx > -INFINITY) atomicMax(&cs->max_enc, enc_ordered(mx));
        if (lane == 1 && mn_kept < INFINITY) atomicMax(&cs->min_inv, ~enc_ordered(mn_kept));
        if (lane == 2) p.tile_min[d_pub.x] = mn_live;
      }
    }
    WFT_TL(12);
    lstate ^= kCtlSlot;   // the next tile becomes current; its slot is rewritten only behind the tile-after-next's first barrier
  }
#undef DESC
#undef NDESC
#undef DESC4
#undef DESCRIBE_NEXT
}

}  // namespace wft
