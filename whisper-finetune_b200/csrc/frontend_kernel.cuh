// Fused Whisper audio front end for sm_100a: PCM -> log-mel (+ cut / min-pad / SpecAugment masks), one launch.
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
//   mel phase = thread <-> one mel row x 16 (or 8) consecutive frames: the power tile is stored [bin][frame], so one tap of
//               the sparse triangular filter is LDS.128 + 2 FFMA2 per 4 frames with the weight held in a register; rows
//               are banded over the warps by tap count (wft_tables.inc) and each band's tap loop is fully unrolled;
//               log10 via MUFU.LG2; the FINAL feature (L + 4) / 4 with the SpecAugment masks applied is written once to
//               `out` as 32-byte stores (STG.256).
//   per-clip max / min = ordered-int atomicMax into the workspace, completion counted per tile.  Once a clip is complete
//               each CTA re-visits ITS OWN tiles of that clip only if something is still missing: the max-8 floor binds
//               somewhere in the tile, the tile holds min-value pad frames, or it is a silent (all-zero PCM) tile that
//               was never computed.  That fix-up runs on L2-resident lines, so HBM sees each output byte once.
//   scheduling = persistent CTAs pulling tiles from an atomic counter in clip-major order, one tile ahead; the next tile's
//               PCM travels global->shared by TMA bulk copies (mbarrier completion) under the current tile's mel phase;
//               a CTA never waits while tiles are still unclaimed (pending tiles are ringed / parked), so the kernel is
//               deadlock-free for any grid size.
//
// Shared memory per CTA: one 28.8 KB region time-multiplexed as
//   [audio tile at the top] -> stage A->B exchange -> [power tile at the bottom | next audio tile at the top],
// plus ~7.8 KB of window / twiddle / mel-weight tables and control words.  The thread order and every stride in here
// were chosen against measured shared-memory wavefront costs (tools/micro/smem_wavefronts.cu,
// profiles/r01_smem_wavefront_probe.md): the data pipe, not HBM, is what bounds this kernel.  The hot loop is one DFT20
// copy per stage and ~2.4 k SASS instructions so that it stays resident in the SM's instruction cache.
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
constexpr int kMaxPending = 8;   // <= 8: ring slots in sm_ctl
constexpr int kCtlInts = 84;
constexpr int kSmemBytes = (kRegionFloats + kWinFloats + kTwFloats + kMelWFloats) * 4 + kCtlInts * 4;

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

// mel plan (gen_tables.py): warp w runs tap class mel_warp_class(w) = (taps, quads of 4 frames per thread, weight stride)
template <int NM>
__host__ __device__ constexpr int mel_n_classes() { return NM == 80 ? WFT_MEL80_NCLASS : WFT_MEL128_NCLASS; }
template <int NM>
__host__ __device__ constexpr int mel_class_taps(int c) {
  constexpr int a[WFT_MEL_MAX_CLASSES] = WFT_MEL80_CLASS_T, b[WFT_MEL_MAX_CLASSES] = WFT_MEL128_CLASS_T;
  return NM == 80 ? a[c] : b[c];
}
template <int NM>
__host__ __device__ constexpr int mel_class_quads(int c) {
  constexpr int a[WFT_MEL_MAX_CLASSES] = WFT_MEL80_CLASS_NQ, b[WFT_MEL_MAX_CLASSES] = WFT_MEL128_CLASS_NQ;
  return NM == 80 ? a[c] : b[c];
}
template <int NM>
__host__ __device__ constexpr int mel_class_wstride(int c) {
  constexpr int a[WFT_MEL_MAX_CLASSES] = WFT_MEL80_CLASS_WS, b[WFT_MEL_MAX_CLASSES] = WFT_MEL128_CLASS_WS;
  return NM == 80 ? a[c] : b[c];
}
template <int NM>
__host__ __device__ constexpr int mel_warp_class(int w) {
  constexpr int a[kWarps] = WFT_MEL80_WARP_CLASS, b[kWarps] = WFT_MEL128_WARP_CLASS;
  return NM == 80 ? a[w] : b[w];
}
template <int NM>
__host__ __device__ constexpr bool mel_any_wide() {   // does any class hold 16 frames (4 quads) per thread?
  for (int c = 0; c < mel_n_classes<NM>(); ++c)
    if (mel_class_quads<NM>(c) == 4) return true;
  return false;
}

struct ClipStat {
  uint32_t max_enc;   // ordered-int encoding of max log10(mel) over ALL frames of the clip
  uint32_t min_inv;   // ~encoding of min log10(mel) over the KEPT frames (pad value of pad_or_trim)
  uint32_t done;      // tiles of this clip whose stat atomics have been performed (and values written)
  uint32_t pad_;
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
  int32_t* next;        // [total_tiles] per-CTA chains of parked fix-ups
  int32_t n_samples;
  int32_t n_total;      // n_samples + padding
  int32_t batch;
  int32_t n_frames;     // n_total / 160
  int32_t n_frames_out;
  int32_t tiles_per_clip;
  int32_t total_tiles;
  float mask_value;
  uint32_t zero;        // always 0; gives the completion counter a data dependency the compiler cannot fold
};

__device__ __forceinline__ uint32_t enc_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t e) {
  const uint32_t b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  return __uint_as_float(b);
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
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// interior tile: the 2800 samples travel as 9 bulk copies (8 blocks of 320 samples + a 240-sample tail), block b landing
// at element (320 + skew) * b (the skew keeps stage A's stride-20 gathers conflict free).  Called by ONE thread, as a real
// call (measured: inlining it, or spreading the 9 copies over 9 lanes, made the kernel 2 % slower).
template <typename PcmT>
__device__ __noinline__ void prefetch_audio(PcmT* __restrict__ sm_audio, const PcmT* __restrict__ src, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy use of the region comes first
  mbar_expect_tx(bar, kTileSamples * sizeof(PcmT));
  constexpr int kBlocks = (kTileSamples + kSkewBlock - 1) / kSkewBlock;  // 9
#pragma unroll 1
  for (int b = 0; b < kBlocks; ++b) {
    const int n = (b + 1) * kSkewBlock <= kTileSamples ? kSkewBlock : kTileSamples - b * kSkewBlock;
    tma_bulk_g2s(sm_audio + b * (kSkewBlock + Skew<PcmT>::value), src + b * kSkewBlock, n * sizeof(PcmT), bar);
  }
}

// generic synchronous staging (clip edges, ragged lengths, unaligned sources)
template <typename PcmT>
__device__ __noinline__ void stage_audio_edge(PcmT* __restrict__ sm_audio, const PcmT* __restrict__ x, int g0,
                                              int len, int n_total, int tid) {
  for (int m = tid; m < kTileSamples; m += kThreads)
    sm_audio[m + Skew<PcmT>::value * (m / kSkewBlock)] = load_sample(x, g0 + m, len, n_total);
}

__device__ __forceinline__ float pcm_as_float(float v) { return v; }
__device__ __forceinline__ float pcm_as_float(int16_t v) { return static_cast<float>(v); }

// relaxed gpu-scope read of a clip's completion counter (ordering comes from data dependencies, see publish)
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
// one 16-byte snapshot {max_enc, min_inv, done, -} of a clip's statistics (a single L2 sector access)
__device__ __forceinline__ uint4 ld_stat(const ClipStat* st) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(st) : "memory");
  return v;
}
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

constexpr int kSilentBit = 1 << 30;   // flag carried by the tile id inside the pending ring / parked chain
constexpr int kTileIdMask = kSilentBit - 1;

// log10 of the 1e-10 clamp exactly as the mel phase computes it for an all-zero frame
__device__ __forceinline__ float silent_log_mel() { return fast_log2(1e-10f) * 0.301029995663981195f; }

// ---- mel projection ---------------------------------------------------------------------------------------------------
// A thread owns ONE mel row for 16 or 8 consecutive frames and walks them 8 at a time.  pk = &P[start_bin][first frame]
// in the [bin][frame] power tile, w = the thread's weight column (tap j at w[j * WS]).  Taps run in ascending bin order;
// padded taps carry a zero weight on an always-finite bin, so the sum is exactly the dense row product.
template <int T, int WS>
__device__ __forceinline__ void mel_taps(const float* __restrict__ pk, const float* __restrict__ w, cpx (&acc)[4]) {
#pragma unroll
  for (int j = 0; j < T; ++j) {
    const cpx ww = splat(w[j * WS]);
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {
      const float4 v = *reinterpret_cast<const float4*>(pk + j * kPStride + 4 * qd);
      if (j == 0) {
        acc[2 * qd] = cmul(ww, make_float2(v.x, v.y));
        acc[2 * qd + 1] = cmul(ww, make_float2(v.z, v.w));
      } else {
        acc[2 * qd] = cfma(ww, make_float2(v.x, v.y), acc[2 * qd]);
        acc[2 * qd + 1] = cfma(ww, make_float2(v.z, v.w), acc[2 * qd + 1]);
      }
    }
  }
}
// warp-uniform dispatch on the warp's tap class: one fully unrolled tap loop per class
template <int NM, int C>
__device__ __forceinline__ void mel_dispatch(int cls, const float* __restrict__ pk, const float* __restrict__ w, cpx (&acc)[4]) {
  if constexpr (C < mel_n_classes<NM>()) {
    if (C == mel_n_classes<NM>() - 1 || cls == C)
      mel_taps<mel_class_taps<NM>(C), mel_class_wstride<NM>(C)>(pk, w, acc);
    else
      mel_dispatch<NM, C + 1>(cls, pk, w, acc);
  }
}

// frame windows of a tile as 16-bit masks (bit f = frame t0 + f); only edge tiles look at them
struct MelEdge {
  uint32_t live;    // real frame of the clip: counts for the max and for the floor test
  uint32_t kept;    // survives the partial-segment cut: counts for the pad minimum
  uint32_t store;   // has a cell in `out`
  uint32_t tmask;   // inside the SpecAugment time mask
};
__device__ __forceinline__ uint32_t frame_window(int lo, int hi, int t0) {   // frames [lo, hi) as bits of the tile at t0
  lo = min(max(lo - t0, 0), kTileFrames);
  hi = min(max(hi - t0, 0), kTileFrames);
  return hi > lo ? (1u << hi) - (1u << lo) : 0u;
}
__device__ __forceinline__ void st_global_256(float* dst, const float (&v)[8]) {   // one full 32-byte sector per thread
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// 8 consecutive frames of one mel row: log10, statistics, final feature (L + 4) / 4 (row mask folded into sc / of), store.
// kFast: every frame is live, kept, stored and outside the time mask, `dst` is 32-byte aligned.
// Otherwise bit i of the (already shifted) windows in `e` describes frame i of these 8.
template <bool kFast>
__device__ __forceinline__ void mel_post8(const cpx* __restrict__ acc, float sc, float of, float mask_value,
                                          float* __restrict__ dst, const MelEdge& e, float& mx, float& mn_kept, float& mn_live) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = (i & 1) ? acc[i >> 1].y : acc[i >> 1].x;
    const float L = fast_log2(fmaxf(a, 1e-10f)) * 0.301029995663981195f;
    v[i] = fmaf(L, sc, of);   // (L + 4) / 4, exactly
    if (kFast) {
      mx = fmaxf(mx, L);
      mn_live = fminf(mn_live, L);
    } else {
      if ((e.live >> i) & 1u) {
        mx = fmaxf(mx, L);
        mn_live = fminf(mn_live, L);
      }
      if ((e.kept >> i) & 1u) mn_kept = fminf(mn_kept, L);
      if ((e.tmask >> i) & 1u) v[i] = mask_value;
      if ((e.store >> i) & 1u) dst[i] = v[i];
    }
  }
  if (kFast) st_global_256(dst, v);
}

// ---- deferred fix-up of one tile (only tiles that need it, see `tile_needs_fixup`) -----------------------------------
// The mel phase wrote v = (L + 4) / 4 with masks applied.  What may still be missing once the clip's max / min are
// known: the floor max(L, max - 8)  ==  max(v, (max - 8 + 4) / 4)  (monotone map, exact), and the min-value pad of
// the frames beyond the kept part (data/utils.py:380-404).  Masked cells keep the mask value.
struct FixupArgs {
  float* out;
  const ClipStat* stats;
  const int32_t* n_valid;
  const int32_t* masks;
  int32_t n_frames, n_frames_out, tiles_per_clip;
  float mask_value;
};

__device__ __forceinline__ int kept_frames(const int32_t* n_valid, int clip, int n_frames) {
  int keep = n_frames;
  if (n_valid != nullptr) {
    const int nv = __ldg(n_valid + clip);
    if (nv >= 0 && nv < keep) keep = nv;
  }
  return keep;
}

// does tile (clip, t0) still differ from its final value?  tile_min = min log10(mel) over the tile's live cells
__device__ __forceinline__ bool tile_needs_fixup(float tile_min, uint32_t max_enc, int t0, int keep, int n_frames_out) {
  const bool floor_binds = tile_min < dec_ordered(max_enc) - 8.0f;
  const int hi = t0 + kTileFrames < n_frames_out ? t0 + kTileFrames : n_frames_out;
  const bool has_pad = (t0 > keep ? t0 : keep) < hi;
  return floor_binds || has_pad;
}

template <int NM>
__device__ __noinline__ void fixup_tile(const FixupArgs p, int tagged_tile, int clip, int tid) {
  const bool silent = (tagged_tile & kSilentBit) != 0;   // nothing was written yet: every cell is the clamp constant
  const int tile = tagged_tile & kTileIdMask;
  const float vsilent = fmaf(silent_log_mel(), 0.25f, 1.0f);
  const ClipStat* st = p.stats + clip;
  const float lmax = dec_ordered(__ldcg(&st->max_enc));
  const float lmin = dec_ordered(~__ldcg(&st->min_inv));
  const float floorv = lmax - 8.0f;
  const float floorn = fmaf(floorv, 0.25f, 1.0f);
  const float padv = fmaf(fmaxf(lmin, floorv), 0.25f, 1.0f);
  const int keep = kept_frames(p.n_valid, clip, p.n_frames);
  int mt0 = 0, mt1 = 0, mf0 = 0, mf1 = 0;
  if (p.masks != nullptr) {
    const int4 mk = __ldg(reinterpret_cast<const int4*>(p.masks) + clip);
    mt0 = mk.x; mt1 = mk.y; mf0 = mk.z; mf1 = mk.w;
  }
  const float mv = p.mask_value;
  const int pitch = p.n_frames_out;
  float* base = p.out + static_cast<size_t>(clip) * NM * pitch;
  const int t0 = (tile - clip * p.tiles_per_clip) * kTileFrames;
  {
    if ((pitch & 3) == 0) {
      constexpr int kGroups = kTileFrames / 4;                      // float4 groups per row (4)
      constexpr int kVec = NM * kGroups;                            // float4 groups per tile
      constexpr int kIters = (kVec + kThreads - 1) / kThreads;      // 4 (128 mel) / 2 (80 mel)
      const int f = t0 + ((tid & (kGroups - 1)) << 2);              // kThreads % kGroups == 0: same column group every iter
      if (f >= pitch) return;
      float4 v[kIters];
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int row = (tid + kThreads * it) / kGroups;
        v[it] = make_float4(vsilent, vsilent, vsilent, vsilent);
        if (!silent && row < NM && f < keep)
          v[it] = __ldcg(reinterpret_cast<const float4*>(base + static_cast<size_t>(row) * pitch + f));
      }
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int row = (tid + kThreads * it) / kGroups;
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
    } else {
      for (int idx = tid; idx < NM * kTileFrames; idx += kThreads) {
        const int row = idx / kTileFrames;
        const int f = t0 + (idx % kTileFrames);
        if (f >= pitch) continue;
        float* ptr = base + static_cast<size_t>(row) * pitch + f;
        float r = padv;
        if (f < keep) r = fmaxf(silent ? vsilent : __ldcg(ptr), floorn);
        if ((row >= mf0 && row < mf1) || (f >= mt0 && f < mt1)) r = mv;
        *ptr = r;
      }
    }
  }
}

__device__ __forceinline__ FixupArgs make_fixup_args(const FrontendParams& p) {
  FixupArgs fx;
  fx.out = p.out; fx.stats = p.stats; fx.n_valid = p.n_valid; fx.masks = p.masks;
  fx.n_frames = p.n_frames; fx.n_frames_out = p.n_frames_out; fx.tiles_per_clip = p.tiles_per_clip;
  fx.mask_value = p.mask_value;
  return fx;
}

// sm_ctl slots
// [kCtlDesc .. +5] and [kCtlDesc + kCtlSlot .. +5] = two tile descriptors {id, clip, first frame, kind, PCM element offset
// (lo, hi)}: the tile being worked on and the one after it, swapping roles every iteration.  Every phase re-reads the few
// fields it needs from here instead of carrying them in registers across the FFT stages.
enum { kCtlDesc = 0, kCtlSlot = 8, kCtlReady = 76, kCtlDrain = 77, kCtlDrainClip = 78, kCtlNRing = 79, kCtlChain = 80, kCtlList = 16, kCtlListClip = 24, kCtlRing = 32,
       kCtlRingClip = 40, kCtlRingMin = 48, kCtlRed = 56, kCtlMbar = 72, kCtlMemo = 74 };   // kCtlRed: 3 floats per warp (max, kept min, live min)

// how a tile's PCM reaches shared memory
enum { kTileEdge = 0,      // reflection / zero extension / unaligned source: scalar staging
       kTileInterior = 1,  // plain in-range samples: TMA bulk copies, one tile ahead
       kTileSilent = 2,    // every sample it touches is zero padding: no FFT, the fix-up later writes the constant rows;
                           // this is the FIRST such tile of its clip and records the clamp value in the clip statistics
       kTileSilentRest = 3 };  // a later tile of the same silent tail: the statistics are already in, only counted
// thread 0: describe tile `t` for everybody (one division and one lengths[] load per tile instead of 160)
__device__ __forceinline__ bool tile_is_silent(int t0, int len, int n_total) {
  const int g0 = t0 * kHop - kNfft / 2;
  const int last = g0 + kTileSamples - 1;  // beyond n_total the samples are mirrored around n_total - 1
  return len == 0 || (g0 >= len && (last < n_total || 2 * (n_total - 1) - last >= len));
}

struct TileGeom {   // the few launch constants describe_tile needs (passed by value: it is a real call, thread 0 only)
  const void* pcm;
  const int32_t* lengths;
  int64_t clip_stride;
  int32_t n_samples, n_total, n_frames, tiles_per_clip, total_tiles;
};
__device__ __forceinline__ TileGeom tile_geom(const FrontendParams& q) {
  TileGeom g;
  g.pcm = q.pcm; g.lengths = q.lengths; g.clip_stride = q.clip_stride; g.n_samples = q.n_samples; g.n_total = q.n_total;
  g.n_frames = q.n_frames; g.tiles_per_clip = q.tiles_per_clip; g.total_tiles = q.total_tiles;
  return g;
}

// (forced inline: as a real call this sat on thread 0's critical path before a CTA barrier and cost 2.5 % overall)
template <typename PcmT>
__device__ __forceinline__ void describe_tile(const TileGeom p, int t, int* __restrict__ slot, int* __restrict__ memo) {
  int& memo_clip = memo[0];   // lengths[] of the last clip described (kept in shared memory, not in registers)
  int& memo_len = memo[1];
  int clip = 0, t0 = 0, interior = kTileEdge;
  long long off = 0;
  if (t < p.total_tiles) {
    clip = t / p.tiles_per_clip;
    t0 = (t - clip * p.tiles_per_clip) * kTileFrames;
    if (t0 < p.n_frames) {
      if (clip != memo_clip) {  // consecutive tiles mostly belong to the same clip: one lengths[] load per clip
        int len = p.n_samples;
        if (p.lengths != nullptr) {
          const int l = __ldg(p.lengths + clip);
          len = l < 0 ? 0 : (l < len ? l : len);
        }
        memo_clip = clip;
        memo_len = len;
      }
      const int len = memo_len;
      const int g0 = t0 * kHop - kNfft / 2;
      off = static_cast<long long>(clip) * p.clip_stride + g0;
      if (tile_is_silent(t0, len, p.n_total)) {
        // silence is a suffix of the clip: only its first tile does any bookkeeping
        interior = (t0 == 0 || !tile_is_silent(t0 - kTileFrames, len, p.n_total)) ? kTileSilent : kTileSilentRest;
      } else if (tile_is_interior(reinterpret_cast<const PcmT*>(p.pcm) + off - g0, g0, len)) {
        interior = kTileInterior;
      }
    }
  }
  slot[0] = t; slot[1] = clip; slot[2] = t0; slot[3] = interior;
  slot[4] = static_cast<int>(off & 0xffffffffll); slot[5] = static_cast<int>(off >> 32);
}

template <int NM, typename PcmT>
__global__ void __launch_bounds__(kThreads, 6) frontend_kernel(const FrontendParams p) {
  extern __shared__ __align__(16) float smem[];
  float* sm_region = smem;
  PcmT* sm_audio = reinterpret_cast<PcmT*>(smem + kAudioBase);
  float* sm_win = smem + kRegionFloats;
  float* sm_tw = sm_win + kWinFloats;
  float* sm_melw = sm_tw + kTwFloats;
  int* sm_ctl = reinterpret_cast<int*>(sm_melw + kMelWFloats);
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
  uint64_t* audio_bar = reinterpret_cast<uint64_t*>(sm_ctl + kCtlMbar);  // completion of the audio tile's bulk copies
  constexpr int kTmaThread = kThreads - 32;                               // the thread that issues the bulk copies
  if (tid == 0) {
    sm_ctl[kCtlMemo] = -1;
    sm_ctl[kCtlNRing] = 0;
    sm_ctl[kCtlChain] = -1;
    mbar_init(audio_bar, 1);
    describe_tile<PcmT>(tile_geom(p), static_cast<int>(atomicAdd(p.tile_counter, 1u)), sm_ctl + kCtlDesc, sm_ctl + kCtlMemo);
  }
  __syncthreads();
  // a tile of kind kTileInterior always arrives by TMA: the first one is sent here, every later one under the tile before it
  if (tid == kTmaThread && sm_ctl[kCtlDesc + 3] == kTileInterior) {
    const long long off = (static_cast<long long>(sm_ctl[kCtlDesc + 5]) << 32) | static_cast<unsigned int>(sm_ctl[kCtlDesc + 4]);
    prefetch_audio<PcmT>(sm_audio, reinterpret_cast<const PcmT*>(p.pcm) + off, audio_bar);
  }
  // loop state in ONE register: bit 3 (kCtlSlot) = descriptor slot of the CURRENT tile, bit 0 = parity of the audio mbarrier
  int lstate = 0;

  // mel phase role of this thread (fixed for the whole launch): row | start_bin << 8 | first_quad << 16 | weight_offset << 18
  // (kept packed in ONE register across the FFT stages; unpacked again in every mel phase)
  const uint32_t mel_desc = (NM == 128 ? g_mel128_thread : g_mel80_thread)[tid];
  const uint32_t tiles_per_clip_u = static_cast<uint32_t>(p.tiles_per_clip);

  // warp-0 scheduler state lives in sm_ctl (not in registers): ring of tiles whose fix-up is pending, its fill count
  // kCtlNRing and the head kCtlChain of the parked chain

  for (;;) {
    // current tile: {id, clip, t0, kind} at DESC; thread 0 describes the next tile at NDESC (behind this tile's first barrier)
    // (one 16-byte read {id, clip, t0, kind} per phase that needs them)
#define DESC (sm_ctl + (launder(lstate) & kCtlSlot))
#define NDESC (sm_ctl + ((launder(lstate) & kCtlSlot) ^ kCtlSlot))
#define DESC4 (*reinterpret_cast<const int4*>(DESC))
    const int4 d_top = DESC4;
    if (d_top.x >= p.total_tiles) break;
    // claim the tile AFTER this one now; the answer is consumed two barriers later (latency hidden by stage A)
    int nxt_claim = 0;
    if (tid == 0) nxt_claim = static_cast<int>(atomicAdd(p.tile_counter, 1u));
    uint32_t seen_max = 0u, seen_done = 0u;  // warp 0: snapshot of the clip of ring[lane] (max_enc, done), sampled early

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
        __syncthreads();  // audio is dead from here on: the region becomes the exchange buffer
        // (letting half 0 run ahead here with bar.arrive / bar.sync was measured 2-3 % slower)
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
      if (tid == 0) describe_tile<PcmT>(tile_geom(p), nxt_claim, NDESC, sm_ctl + kCtlMemo);
      __syncthreads();

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
        __syncthreads();  // exchange is dead: the region becomes power tile (bottom) + next audio tile (top)
        // prefetch the NEXT tile's PCM into the top of the region (its descriptor stays in sm_ctl until the loop ends)
        if (NDESC[3] == kTileInterior) {
          const long long off = (static_cast<long long>(NDESC[5]) << 32) | static_cast<unsigned int>(NDESC[4]);
          if (tid == kTmaThread) prefetch_audio<PcmT>(sm_audio, reinterpret_cast<const PcmT*>(p.pcm) + off, audio_bar);
        }
        if (warp == 0 && lane < sm_ctl[kCtlNRing]) {
          const uint4 st = ld_stat(p.stats + sm_ctl[kCtlRingClip + lane]);
          seen_max = st.x;
          seen_done = st.z;
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
      __syncthreads();

      // mel phase: thread <-> (mel row, 16 or 8 frames) -------------------------------------------------------------
      {
        const int4 d_mel = DESC4;
        const int clip = d_mel.y, t0 = d_mel.z;
        const int mel_row = static_cast<int>(mel_desc & 0xffu);
        const int mel_q0 = static_cast<int>((mel_desc >> 16) & 3u);
        const float* mel_p = sm_region + ((mel_desc >> 8) & 0xffu) * kPStride + 4 * mel_q0;
        const float* mel_w = sm_melw + (mel_desc >> 18);
        int mel_cls = 0;
        bool mel_wide = false;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
          if (warp == w) {
            mel_cls = mel_warp_class<NM>(w);
            mel_wide = mel_class_quads<NM>(mel_warp_class<NM>(w)) == 4;
          }
        // 32-byte stores need an aligned `out` and a row pitch that is a multiple of 8 frames
        const bool out_vec_ok = (reinterpret_cast<uintptr_t>(p.out) & 31) == 0 && (p.n_frames_out & 7) == 0;
        const int keep = kept_frames(p.n_valid, clip, p.n_frames);
        int4 mk = make_int4(0, 0, 0, 0);
        if (p.masks != nullptr) mk = __ldg(reinterpret_cast<const int4*>(p.masks) + clip);
        const bool rowmask = mel_row >= mk.z && mel_row < mk.w;
        const float sc = rowmask ? 0.0f : 0.25f;            // masked row: 0 * L + mask_value
        const float of = rowmask ? p.mask_value : 1.0f;
        // fast tile (uniform over the CTA): all 16 frames live and stored, none inside the time mask, and either all of
        // them kept or none (frames beyond the partial-segment cut still count for the max; the fix-up pads them later)
        const bool all_kept = t0 + kTileFrames <= keep;
        const bool fast = out_vec_ok && (all_kept || keep <= t0) && t0 + kTileFrames <= p.n_frames &&
                          t0 + kTileFrames <= p.n_frames_out && (mk.y <= t0 || mk.x >= t0 + kTileFrames || mk.y <= mk.x);
        float* dst = p.out + (static_cast<size_t>(clip) * NM + mel_row) * p.n_frames_out + t0 + 4 * mel_q0;

        float mx = -INFINITY, mn_kept = INFINITY, mn_live = INFINITY;
        MelEdge e;
        if (!fast) {
          const int sh = 4 * mel_q0;
          e.live = frame_window(0, p.n_frames, t0) >> sh;
          e.kept = frame_window(0, keep, t0) >> sh;
          e.store = frame_window(0, p.n_frames < p.n_frames_out ? p.n_frames : p.n_frames_out, t0) >> sh;
          e.tmask = frame_window(mk.x, mk.y, t0) >> sh;
        }
        const int n_halves = (mel_any_wide<NM>() && mel_wide) ? 2 : 1;   // warp-uniform
#pragma unroll 1
        for (int half = 0; half < n_halves; ++half) {
          cpx acc[4];
          mel_dispatch<NM, 0>(mel_cls, mel_p, mel_w, acc);
          if (fast) {
            mel_post8<true>(acc, sc, of, p.mask_value, dst, e, mx, mn_kept, mn_live);
          } else {
            mel_post8<false>(acc, sc, of, p.mask_value, dst, e, mx, mn_kept, mn_live);
            e.live >>= 8; e.kept >>= 8; e.store >>= 8; e.tmask >>= 8;
          }
          mel_p += 8;
          dst += 8;
        }
        if (fast && all_kept) mn_kept = mn_live;
        mx = warp_max(mx);
        mn_kept = warp_min(mn_kept);
        mn_live = warp_min(mn_live);
        if (lane == 0) {
          float* red = reinterpret_cast<float*>(sm_ctl + kCtlRed) + 3 * warp;
          red[0] = mx;
          red[1] = mn_kept;
          red[2] = mn_live;
        }
      }
    } else {
      // nothing to compute: a pad-only tile (n_frames_out > n_frames) or a silent tile (all-zero PCM: every mel value is
      // the 1e-10 clamp, so only its statistics are recorded here and the fix-up later writes the constant rows)
      __syncthreads();  // the previous tile's last readers of the other descriptor slot are done
      if (tid == 0) describe_tile<PcmT>(tile_geom(p), nxt_claim, NDESC, sm_ctl + kCtlMemo);
      __syncthreads();
      if (NDESC[3] == kTileInterior && tid == kTmaThread) {
        const long long off = (static_cast<long long>(NDESC[5]) << 32) | static_cast<unsigned int>(NDESC[4]);
        prefetch_audio<PcmT>(sm_audio, reinterpret_cast<const PcmT*>(p.pcm) + off, audio_bar);
      }
      if (warp == 0 && lane < sm_ctl[kCtlNRing]) {
        const uint4 st = ld_stat(p.stats + sm_ctl[kCtlRingClip + lane]);
        seen_max = st.x;
        seen_done = st.z;
      }
      if (lane == 0) {
        const int clip = d_top.y, t0 = d_top.z, kind = d_top.w;   // (short path: the loop-top read is still in registers)
        const bool silent = (kind == kTileSilent || kind == kTileSilentRest) && t0 < p.n_frames;
        const bool silent_head = kind == kTileSilent;   // only the first silent tile of a clip touches the clip statistics
        float* red = reinterpret_cast<float*>(sm_ctl + kCtlRed) + 3 * warp;
        const float lc = silent_log_mel();
        red[0] = (silent && silent_head) ? lc : -INFINITY;                                                 // live frames
        red[1] = (silent && silent_head && t0 < kept_frames(p.n_valid, clip, p.n_frames)) ? lc : INFINITY;  // kept frames
        red[2] = silent ? lc : INFINITY;
      }
    }

    // warp 0 looks at the pending ring: a tile whose clip is complete either needs the fix-up (floor binds or it
    // carries pad frames) or is already final and simply leaves the ring; everything else keeps waiting
    if (warp == 0) {
      const bool pending = lane < sm_ctl[kCtlNRing];
      const int mine = pending ? sm_ctl[kCtlRing + lane] : -1;
      const int mine_clip = pending ? sm_ctl[kCtlRingClip + lane] : 0;
      const float mine_min = pending ? __int_as_float(sm_ctl[kCtlRingMin + lane]) : 0.0f;
      const bool complete = pending && seen_done >= tiles_per_clip_u;
      bool ready = false;
      if (complete) {
        const int mt0 = ((mine & kTileIdMask) - mine_clip * p.tiles_per_clip) * kTileFrames;
        ready = (mine & kSilentBit) != 0 ||
                tile_needs_fixup(mine_min, seen_max, mt0, kept_frames(p.n_valid, mine_clip, p.n_frames), p.n_frames_out);
      }
      const uint32_t ready_mask = __ballot_sync(0xffffffffu, ready);
      const uint32_t wait_mask = __ballot_sync(0xffffffffu, pending && !complete);
      const uint32_t below = (1u << lane) - 1u;
      __syncwarp();
      if (ready) {
        sm_ctl[kCtlList + __popc(ready_mask & below)] = mine;
        sm_ctl[kCtlListClip + __popc(ready_mask & below)] = mine_clip;
      } else if (pending && !complete) {
        sm_ctl[kCtlRing + __popc(wait_mask & below)] = mine;
        sm_ctl[kCtlRingClip + __popc(wait_mask & below)] = mine_clip;
        sm_ctl[kCtlRingMin + __popc(wait_mask & below)] = __float_as_int(mine_min);
      }
      if (lane == 0) {
        sm_ctl[kCtlNRing] = __popc(wait_mask);
        sm_ctl[kCtlReady] = __popc(ready_mask);
      }
    }
    __syncthreads();  // tile finished: power tile free, ready list and per-warp max/min visible

    // publish the tile's statistics: two returning atomics, then the completion count with a true data dependency on
    // their results (through p.zero) -- no fence, so nobody waits for the tile's stores to drain.  The tile itself
    // joins the pending ring (or is parked: a CTA never waits while tiles are unclaimed).
    if (warp == 0) {
      const float* red = reinterpret_cast<const float*>(sm_ctl + kCtlRed) + 3 * (lane < kWarps ? lane : 0);
      float mx = lane < kWarps ? red[0] : -INFINITY;
      float mn_kept = lane < kWarps ? red[1] : INFINITY;
      float mn_live = lane < kWarps ? red[2] : INFINITY;
      mx = warp_max(mx);
      mn_kept = warp_min(mn_kept);
      mn_live = warp_min(mn_live);
      if (lane == 0) {
        const int4 d_pub = DESC4;
        const int cur = d_pub.x, clip = d_pub.y;
        const bool silent = (d_pub.w == kTileSilent || d_pub.w == kTileSilentRest) && d_pub.z < p.n_frames;
        ClipStat* cs = p.stats + clip;
        uint32_t dep = 0;
        if (mx > -INFINITY) dep |= atomicMax(&cs->max_enc, enc_ordered(mx));
        if (mn_kept < INFINITY) dep |= atomicMax(&cs->min_inv, ~enc_ordered(mn_kept));
        const int tagged = cur | (silent ? kSilentBit : 0);
        const int n_ring = sm_ctl[kCtlNRing];
        if (n_ring < kMaxPending) {
          sm_ctl[kCtlRing + n_ring] = tagged;
          sm_ctl[kCtlRingClip + n_ring] = clip;
          sm_ctl[kCtlRingMin + n_ring] = __float_as_int(mn_live);
          sm_ctl[kCtlNRing] = n_ring + 1;
        } else {
          p.next[cur] = sm_ctl[kCtlChain];           // parked tiles are re-examined (conservatively) in the drain
          sm_ctl[kCtlChain] = tagged;
        }
        // (deferring this count -- to the next tile's gather, or with the two results kept apart -- costs registers the
        // gather does not have: measured slower every time)
        atomicAdd(&cs->done, 1u + (dep & p.zero));
      }
      __syncwarp();
    }
    const int n_ready = sm_ctl[kCtlReady];
#pragma unroll 1
    for (int k = 0; k < n_ready; ++k) fixup_tile<NM>(make_fixup_args(p), sm_ctl[kCtlList + k], sm_ctl[kCtlListClip + k], tid);
    lstate ^= kCtlSlot;   // the next tile becomes current; its slot is rewritten only behind the tile-after-next's first barrier
  }
#undef DESC
#undef NDESC
#undef DESC4

  // drain: every tile is claimed by a running CTA now, so waiting on a clip's counter is safe
  for (;;) {
    __syncthreads();  // previous readers of sm_ctl are done
    if (tid == 0) {
      int t = -1, c = 0;
      int n_ring = sm_ctl[kCtlNRing], chain = sm_ctl[kCtlChain];
      for (;;) {
        float tmin = -INFINITY;  // parked tiles lost their minimum: treat them as needing the fix-up
        if (n_ring > 0) {
          --n_ring;
          t = sm_ctl[kCtlRing + n_ring];
          c = sm_ctl[kCtlRingClip + n_ring];
          tmin = __int_as_float(sm_ctl[kCtlRingMin + n_ring]);
        } else if (chain >= 0) {
          t = chain;
          c = (t & kTileIdMask) / p.tiles_per_clip;
          chain = p.next[t & kTileIdMask];
        } else {
          t = -1;
          break;
        }
        uint4 st = ld_stat(p.stats + c);
        while (st.z < tiles_per_clip_u) {
          __nanosleep(100);
          st = ld_stat(p.stats + c);
        }
        const int tt0 = ((t & kTileIdMask) - c * p.tiles_per_clip) * kTileFrames;
        if ((t & kSilentBit) != 0 ||
            tile_needs_fixup(tmin, st.x, tt0, kept_frames(p.n_valid, c, p.n_frames), p.n_frames_out)) break;
      }
      sm_ctl[kCtlNRing] = n_ring;
      sm_ctl[kCtlChain] = chain;
      sm_ctl[kCtlDrain] = t;
      sm_ctl[kCtlDrainClip] = c;
    }
    __syncthreads();
    const int t = sm_ctl[kCtlDrain];
    if (t < 0) break;
    fixup_tile<NM>(make_fixup_args(p), t, sm_ctl[kCtlDrainClip], tid);
  }
}

}  // namespace wft
