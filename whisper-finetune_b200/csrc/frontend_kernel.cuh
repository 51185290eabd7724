// Fused Whisper audio front end for sm_100a: PCM -> log-mel (+ cut / min-pad / SpecAugment masks), one launch.
//
// Replaces, for a whole batch, the per-clip CPU path of the reference
//   np.pad -> whisper.audio.log_mel_spectrogram -> mel[:, :T'] -> pad_or_trim -> time/freq masks -> collate
//   (src/whisper_finetune/data/data_loader.py:346, :278, :279-282, :286-287, :362-367; data/utils.py:380-404).
//
// Work decomposition
//   tile      = 32 consecutive frames of one clip = 16 frame PAIRS; 160 threads = 16 pairs x 10 items.
//   pair      = frames (2q, 2q+1) packed as re/im of ONE 400-point complex FFT (two real frames per transform).
//   400-point = 20 x 20 Cooley-Tukey; every thread runs two register-resident 20-point DFTs per stage
//               (dft20.cuh), stage A over n1 for n2 in {i, i+10}, stage B over n2 for k1 in {i', 20-i'}, so
//               that Z[k] and its mirror Z[400-k] meet in the same thread and the two real spectra are
//               separated without another exchange:  4|Xa|^2 = |Z[k]+conj Z[400-k]|^2, 4|Xb|^2 = |Z[k]-conj Z[400-k]|^2.
//   mel phase = warp g owns a row group, lane <-> frame; sparse triangular filters unrolled with immediate
//               weights (wft_tables.inc); log10 via MUFU.LG2; un-floored log-mel written once to `out`.
//   per-clip max / min = ordered-int atomicMax into the workspace; the max-8 floor, (x+4)/4, the min-value
//               pad and the SpecAugment masks are applied by a deferred in-place "fix-up" of the CTA's OWN
//               tiles once the clip's tile counter is complete -- while those lines are still L2-resident,
//               so HBM sees each output byte once.
//   scheduling = persistent CTAs pulling tiles from an atomic counter in clip-major order; a CTA never waits
//               while tiles are still unclaimed (pending fix-ups are parked), so the kernel is deadlock-free
//               for any grid size.
//
// Shared memory (per CTA, 3 CTAs/SM): one 56.6 KB region time-multiplexed as
//   audio tile (skewed so that stride-10 gathers are conflict-free) -> stage A->B exchange -> power spectrum,
//   plus 4.8 KB of window / twiddle tables.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dft20.cuh"
#include "wft_tables.inc"

namespace wft {

constexpr int kHop = 160;
constexpr int kNfft = 400;
constexpr int kTileFrames = 32;
constexpr int kPairs = kTileFrames / 2;
constexpr int kItems = 10;
constexpr int kThreads = kPairs * kItems;                        // 160
constexpr int kWarps = kThreads / 32;                            // 5 == mel row groups
constexpr int kTileSamples = kTileFrames * kHop + (kNfft - kHop);  // 5360
constexpr int kSkewBlock = 320;                                  // samples per frame pair
constexpr int kSkew = 10;                                        // extra floats per block: bank(tid) = tid + const
constexpr int kAudioFloats = kTileSamples + kSkew * ((kTileSamples - 1) / kSkewBlock);      // 5520
constexpr int kRowStride = 44;                                   // floats per k1 row (20 complex + pad)
constexpr int kPairStride = 20 * kRowStride + 4;                 // 884: pair stride == 20 (mod 32) banks
constexpr int kRegionFloats = kPairs * kPairStride;              // 14144 floats = 56576 B
constexpr int kPStride = 202;                                    // power rows: stride == 10 (mod 32)
constexpr int kPOddBase = 16 * kPStride + 1;                     // odd frames start one bank over
constexpr int kWinFloats = WFT_WINDOW_TABLE_LEN;                 // 400
constexpr int kTwFloats = WFT_TWIDDLE_TABLE_LEN;                 // 880
constexpr int kTwRow = WFT_TWIDDLE_ROW;                          // 44
constexpr int kAudioBase = kRegionFloats - kAudioFloats;         // audio tile lives at the TOP of the region so the
                                                                 // next tile can be prefetched under the power tile
constexpr int kMaxPending = 6;
constexpr int kSmemFloats = kRegionFloats + kWinFloats + kTwFloats;
constexpr int kSmemBytes = kSmemFloats * 4 + 64;

static_assert(kAudioFloats <= kRegionFloats, "audio tile must fit in the shared region");
static_assert(32 * kPStride + 1 <= kAudioBase, "power tile and prefetched audio tile must not overlap");
static_assert((kAudioBase & 1) == 0, "audio base must stay 8-byte aligned");

__device__ const float g_window_table[kWinFloats] = WFT_WINDOW_TABLE_INIT;
__device__ const float g_twiddle_table[kTwFloats] = WFT_TWIDDLE_TABLE_INIT;

struct ClipStat {
  uint32_t max_enc;   // ordered-int encoding of max log10(mel) over ALL frames of the clip
  uint32_t min_inv;   // ~encoding of min log10(mel) over the KEPT frames (pad value of pad_or_trim)
  uint32_t done;      // tiles of this clip whose un-floored values and stats are published
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
  int32_t* next;        // [total_tiles] per-CTA overflow chains of parked fix-ups
  int32_t n_samples;
  int32_t n_total;      // n_samples + padding
  int32_t batch;
  int32_t n_frames;     // n_total / 160
  int32_t n_frames_out;
  int32_t tiles_per_clip;
  int32_t total_tiles;
  float mask_value;
};

__device__ __forceinline__ uint32_t enc_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t e) {
  const uint32_t b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  return __uint_as_float(b);
}

__device__ __forceinline__ float pcm_to_float(float v) { return v; }
__device__ __forceinline__ float pcm_to_float(int16_t v) { return static_cast<float>(v) * (1.0f / 32768.0f); }

// p[j] of the reflect-padded, zero-extended clip; j is relative to sample 0 of the un-padded clip.
template <typename PcmT>
__device__ __forceinline__ float load_sample(const PcmT* __restrict__ x, int j, int len, int n_total) {
  int r = j < 0 ? -j : (j >= n_total ? 2 * (n_total - 1) - j : j);
  return (r >= 0 && r < len) ? pcm_to_float(x[r]) : 0.0f;
}

__device__ __forceinline__ int skewed(int m) { return m + kSkew * (m / kSkewBlock); }

// ---- stage 0: 5360 samples of a tile -> shared memory (float32, skewed) ---------------------------------------
// A tile is "interior" when every sample it touches is a plain in-range sample of the clip (no reflection, no
// zero extension) and the source is 8-byte aligned: then float32 PCM is copied global->shared asynchronously
// (LDGSTS.64) one tile AHEAD, under the mel phase of the previous tile.
template <typename PcmT>
__device__ __forceinline__ bool tile_is_interior(const PcmT* x, int g0, int len) {
  return g0 >= 0 && g0 + kTileSamples <= len && (reinterpret_cast<uintptr_t>(x + g0) & 7) == 0;
}

__device__ __forceinline__ void cp_async_8(float* smem_dst, const float* gsrc) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// float32 interior tile: 2680 8-byte async copies, 16.75 per thread; block b of 320 samples lands at 330*b.
__device__ __forceinline__ void prefetch_audio_f32(float* __restrict__ sm_audio, const float* __restrict__ src,
                                                   int tid) {
#pragma unroll
  for (int it = 0; it < 17; ++it) {
    const int m = 2 * tid + kSkewBlock * it;
    if (it < 16 || m < kTileSamples) cp_async_8(sm_audio + 2 * tid + (kSkewBlock + kSkew) * it, src + m);
  }
}

// generic synchronous staging (edges, int16, unaligned): loads are batched ahead of the stores
template <typename PcmT>
__device__ __forceinline__ void stage_audio_sync(float* __restrict__ sm_audio, const PcmT* __restrict__ x, int g0,
                                                 int len, int n_total, int tid, bool interior) {
  constexpr int kGroups = kTileSamples / 4;                        // 1340 groups of 4 samples
  constexpr int kIters = (kGroups + kThreads - 1) / kThreads;      // 9
  if (interior) {
    if constexpr (sizeof(PcmT) == 2) {
      short4 v[kIters];
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int gi = tid + kThreads * it;
        if (gi < kGroups) v[it] = __ldg(reinterpret_cast<const short4*>(x + g0 + 4 * gi));
      }
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int gi = tid + kThreads * it;
        if (gi < kGroups) {
          float2* dst = reinterpret_cast<float2*>(sm_audio + skewed(4 * gi));
          dst[0] = make_float2(pcm_to_float(v[it].x), pcm_to_float(v[it].y));
          dst[1] = make_float2(pcm_to_float(v[it].z), pcm_to_float(v[it].w));
        }
      }
    } else {
      float2 v[2 * kIters];
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int gi = tid + kThreads * it;
        if (gi < kGroups) {
          v[2 * it] = __ldg(reinterpret_cast<const float2*>(x + g0 + 4 * gi));
          v[2 * it + 1] = __ldg(reinterpret_cast<const float2*>(x + g0 + 4 * gi + 2));
        }
      }
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int gi = tid + kThreads * it;
        if (gi < kGroups) {
          float2* dst = reinterpret_cast<float2*>(sm_audio + skewed(4 * gi));
          dst[0] = v[2 * it];
          dst[1] = v[2 * it + 1];
        }
      }
    }
    return;
  }
  for (int gi = tid; gi < kGroups; gi += kThreads) {
    const int g = g0 + 4 * gi;
    float2* dst = reinterpret_cast<float2*>(sm_audio + skewed(4 * gi));
    dst[0] = make_float2(load_sample(x, g, len, n_total), load_sample(x, g + 1, len, n_total));
    dst[1] = make_float2(load_sample(x, g + 2, len, n_total), load_sample(x, g + 3, len, n_total));
  }
}

// ---- stage A: window, 2 x DFT20 over n1, twiddle, scatter to the exchange ---------------------------------
template <int D>
__device__ __forceinline__ void stage_a_one(const float (&u)[56], const float* __restrict__ sm_win,
                                            const float* __restrict__ sm_tw, float* __restrict__ sm_pair, int i) {
  float xr[20], xi[20];
  const float4* w4 = reinterpret_cast<const float4*>(sm_win + (D * kItems + i) * 20);
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    const float4 w = w4[a];
    const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n1 = 4 * a + e;
      xr[n1] = ww[e] * u[2 * n1 + D];
      xi[n1] = ww[e] * u[2 * n1 + D + 16];
    }
  }
  dft20(xr, xi);
  const float4* t4 = reinterpret_cast<const float4*>(sm_tw + (D * kItems + i) * kTwRow);
  float2* e2 = reinterpret_cast<float2*>(sm_pair) + (i + 10 * D);
#pragma unroll
  for (int h = 0; h < 10; ++h) {
    const float4 t = t4[h];
    const int k0 = 2 * h, k1 = 2 * h + 1;
    e2[k0 * (kRowStride / 2)] = make_float2(xr[k0] * t.x - xi[k0] * t.y, fmaf(xr[k0], t.y, xi[k0] * t.x));
    e2[k1 * (kRowStride / 2)] = make_float2(xr[k1] * t.z - xi[k1] * t.w, fmaf(xr[k1], t.w, xi[k1] * t.z));
  }
}

// ---- stage B helpers ----------------------------------------------------------------------------------------
__device__ __forceinline__ void load_row(const float* __restrict__ row, float (&yr)[20], float (&yi)[20]) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int a = 0; a < 10; ++a) {
    const float4 v = r4[a];
    yr[2 * a] = v.x; yi[2 * a] = v.y;
    yr[2 * a + 1] = v.z; yi[2 * a + 1] = v.w;
  }
}

// power of the two real frames hidden in (Z[k], Z[400-k]) = (z, m):  |z + conj m|^2 and |z - conj m|^2
__device__ __forceinline__ void pair_power(float zr, float zi, float mr, float mi, float& pa, float& pb) {
  const float sr = zr + mr, si = zi - mi;
  const float dr = zr - mr, di = zi + mi;
  pa = fmaf(sr, sr, si * si);
  pb = fmaf(dr, dr, di * di);
}

template <int NM, int G, class EmitT>
__device__ __forceinline__ void mel_group(const float* __restrict__ P, EmitT&& emit) {
  if constexpr (NM == 128) {
    if constexpr (G == 0) wft_mel128_g0(P, emit);
    if constexpr (G == 1) wft_mel128_g1(P, emit);
    if constexpr (G == 2) wft_mel128_g2(P, emit);
    if constexpr (G == 3) wft_mel128_g3(P, emit);
    if constexpr (G == 4) wft_mel128_g4(P, emit);
  } else {
    if constexpr (G == 0) wft_mel80_g0(P, emit);
    if constexpr (G == 1) wft_mel80_g1(P, emit);
    if constexpr (G == 2) wft_mel80_g2(P, emit);
    if constexpr (G == 3) wft_mel80_g3(P, emit);
    if constexpr (G == 4) wft_mel80_g4(P, emit);
  }
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

// ---- deferred fix-up of one tile: floor at max-8, (x+4)/4, min-value pad, SpecAugment masks -------------------
template <int NM>
__device__ __noinline__ void fixup_tile(const FrontendParams& p, int tile, int tid) {
  const TileCoord tc = tile_coord(p, tile);
  const int clip = tc.clip, t0 = tc.t0;
  const ClipStat* st = p.stats + clip;
  const float lmax = dec_ordered(__ldcg(&st->max_enc));
  const float lmin = dec_ordered(~__ldcg(&st->min_inv));
  const float floorv = lmax - 8.0f;
  const float padv = (fmaxf(lmin, floorv) + 4.0f) * 0.25f;
  int keep = p.n_frames;
  if (p.n_valid != nullptr) {
    const int nv = __ldg(p.n_valid + clip);
    if (nv >= 0 && nv < keep) keep = nv;
  }
  int mt0 = 0, mt1 = 0, mf0 = 0, mf1 = 0;
  if (p.masks != nullptr) {
    const int4 mk = __ldg(reinterpret_cast<const int4*>(p.masks) + clip);
    mt0 = mk.x; mt1 = mk.y; mf0 = mk.z; mf1 = mk.w;
  }
  const float mv = p.mask_value;
  const int pitch = p.n_frames_out;
  float* base = p.out + static_cast<size_t>(clip) * NM * pitch;
  if ((pitch & 3) == 0) {
    constexpr int kVec = NM * 8;                                  // float4 groups per tile
    constexpr int kIters = (kVec + kThreads - 1) / kThreads;      // 7 (128 mel) / 4 (80 mel)
    const int f = t0 + ((tid & 7) << 2);                          // kThreads % 8 == 0: same column group every iter
    if (f >= pitch) return;
    float4 v[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int row = (tid + kThreads * it) >> 3;
      v[it] = make_float4(padv, padv, padv, padv);
      if (row < NM && f < keep) v[it] = __ldcg(reinterpret_cast<const float4*>(base + static_cast<size_t>(row) * pitch + f));
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int row = (tid + kThreads * it) >> 3;
      if (row < NM) {
        const bool rowmask = row >= mf0 && row < mf1;
        float e[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int fc = f + c;
          float r = (fc < keep) ? (fmaxf(e[c], floorv) + 4.0f) * 0.25f : padv;
          if (rowmask || (fc >= mt0 && fc < mt1)) r = mv;
          e[c] = r;
        }
        *reinterpret_cast<float4*>(base + static_cast<size_t>(row) * pitch + f) = make_float4(e[0], e[1], e[2], e[3]);
      }
    }
  } else {
    for (int idx = tid; idx < NM * kTileFrames; idx += kThreads) {
      const int row = idx >> 5;
      const int f = t0 + (idx & 31);
      if (f >= pitch) continue;
      float* ptr = base + static_cast<size_t>(row) * pitch + f;
      float r = padv;
      if (f < keep) r = (fmaxf(__ldcg(ptr), floorv) + 4.0f) * 0.25f;
      if ((row >= mf0 && row < mf1) || (f >= mt0 && f < mt1)) r = mv;
      *ptr = r;
    }
  }
}

// sm_ctl slots
enum { kCtlNext = 0, kCtlReady = 1, kCtlDrain = 2, kCtlList = 4 };

template <int NM, typename PcmT>
__global__ void __launch_bounds__(kThreads, 3) frontend_kernel(const FrontendParams p) {
  extern __shared__ __align__(16) float smem[];
  float* sm_region = smem;
  float* sm_audio = smem + kAudioBase;
  float* sm_win = smem + kRegionFloats;
  float* sm_tw = sm_win + kWinFloats;
  int* sm_ctl = reinterpret_cast<int*>(sm_tw + kTwFloats);
  const int tid = threadIdx.x;
  const int q = tid / kItems;
  const int i = tid - q * kItems;
  const int warp = tid >> 5, lane = tid & 31;

  for (int k = tid; k < kWinFloats; k += kThreads) sm_win[k] = g_window_table[k];
  for (int k = tid; k < kTwFloats; k += kThreads) sm_tw[k] = g_twiddle_table[k];
  if (tid == 0) sm_ctl[kCtlNext] = static_cast<int>(atomicAdd(p.tile_counter, 1u));
  __syncthreads();
  int cur = sm_ctl[kCtlNext];
  bool prefetched = false;

  // thread-0 private scheduler state
  int ring[kMaxPending];
  int n_ring = 0;
  int chain = -1;  // head of this CTA's parked-fix-up chain in p.next

  while (cur < p.total_tiles) {
    const TileCoord tc = tile_coord(p, cur);
    const int clip = tc.clip, t0 = tc.t0;
    // claim the tile AFTER this one now; the answer is consumed two barriers later (latency hidden by stage A)
    int nxt_claim = 0;
    if (tid == 0) nxt_claim = static_cast<int>(atomicAdd(p.tile_counter, 1u));
    int nxt = p.total_tiles;

    if (t0 < p.n_frames) {
      // stage 0 ---------------------------------------------------------------------------------------------
      if (!prefetched) {
        const PcmT* x = reinterpret_cast<const PcmT*>(p.pcm) + static_cast<size_t>(clip) * p.clip_stride;
        int len = p.n_samples;
        if (p.lengths != nullptr) {
          const int l = __ldg(p.lengths + clip);
          len = l < 0 ? 0 : (l < len ? l : len);
        }
        const int g0 = t0 * kHop - kNfft / 2;
        stage_audio_sync<PcmT>(sm_audio, x, g0, len, p.n_total, tid, tile_is_interior(x, g0, len));
      } else {
        cp_async_commit_wait_all();
      }
      __syncthreads();

      // stage A ---------------------------------------------------------------------------------------------
      {
        float u[56];
        const float* a = sm_audio + (kSkewBlock + kSkew) * q + i;
#pragma unroll
        for (int j = 0; j < 56; ++j) u[j] = a[10 * j + (j >= 32 ? kSkew : 0)];
        __syncthreads();  // audio is dead from here on: the region becomes the exchange buffer
        float* sm_pair = sm_region + q * kPairStride;
        stage_a_one<0>(u, sm_win, sm_tw, sm_pair, i);
        stage_a_one<1>(u, sm_win, sm_tw, sm_pair, i);
      }
      if (tid == 0) sm_ctl[kCtlNext] = nxt_claim;
      __syncthreads();
      nxt = sm_ctl[kCtlNext];

      // stage B ---------------------------------------------------------------------------------------------
      {
        float ar[20], ai[20], br[20], bi[20];
        const int ra = i;                         // residue class k1 = i  (item 0: classes 0 and 10)
        const int rb = (i == 0) ? 10 : 20 - i;    // and its mirror class 20 - i
        const float* sm_pair = sm_region + q * kPairStride;
        load_row(sm_pair + ra * kRowStride, ar, ai);
        load_row(sm_pair + rb * kRowStride, br, bi);
        __syncthreads();  // exchange is dead: the region becomes power tile (bottom) + next audio tile (top)
        dft20(ar, ai);    // ZA[k2] = Z[ra + 20 k2]
        dft20(br, bi);    // ZB[k2] = Z[rb + 20 k2]
        float* pe = sm_region + q * kPStride;              // even frame 2q   -> power row q
        float* po = sm_region + kPOddBase + q * kPStride;  // odd frame 2q+1  -> power row 16+q
        if (i != 0) {
#pragma unroll
          for (int k2 = 0; k2 < 20; ++k2) {
            // k = i + 20 k2 ; mirror 400 - k = (20 - i) + 20 (19 - k2)
            float pa, pb;
            pair_power(ar[k2], ai[k2], br[19 - k2], bi[19 - k2], pa, pb);
            const int bin = (k2 < 10) ? (i + 20 * k2) : (400 - i - 20 * k2);
            pe[bin] = pa;
            po[bin] = pb;
          }
        } else {
#pragma unroll
          for (int k2 = 1; k2 < 10; ++k2) {  // class 0: k = 20 k2, mirror = 20 (20 - k2)
            float pa, pb;
            pair_power(ar[k2], ai[k2], ar[20 - k2], ai[20 - k2], pa, pb);
            pe[20 * k2] = pa;
            po[20 * k2] = pb;
          }
#pragma unroll
          for (int k2 = 0; k2 < 10; ++k2) {  // class 10: k = 10 + 20 k2, mirror = 10 + 20 (19 - k2)
            float pa, pb;
            pair_power(br[k2], bi[k2], br[19 - k2], bi[19 - k2], pa, pb);
            pe[10 + 20 * k2] = pa;
            po[10 + 20 * k2] = pb;
          }
        }
      }

      // prefetch the NEXT tile's audio into the top of the region (free since the barrier above) ------------------
      prefetched = false;
      if constexpr (sizeof(PcmT) == 4) {
        if (nxt < p.total_tiles) {
          const TileCoord nc = tile_coord(p, nxt);
          if (nc.t0 < p.n_frames) {
            const float* x = reinterpret_cast<const float*>(p.pcm) + static_cast<size_t>(nc.clip) * p.clip_stride;
            int len = p.n_samples;
            if (p.lengths != nullptr) {
              const int l = __ldg(p.lengths + nc.clip);
              len = l < 0 ? 0 : (l < len ? l : len);
            }
            const int g0 = nc.t0 * kHop - kNfft / 2;
            if (tile_is_interior(x, g0, len)) {
              prefetch_audio_f32(sm_audio, x + g0, tid);
              prefetched = true;
            }
          }
        }
      }
      __syncthreads();

      // mel phase -------------------------------------------------------------------------------------------
      {
        const int frame = t0 + (lane < 16 ? 2 * lane : 2 * (lane - 16) + 1);
        const float* P = sm_region + (lane < 16 ? lane * kPStride : kPOddBase + (lane - 16) * kPStride);
        int keep = p.n_frames;
        if (p.n_valid != nullptr) {
          const int nv = __ldg(p.n_valid + clip);
          if (nv >= 0 && nv < keep) keep = nv;
        }
        const bool live = frame < p.n_frames;           // real frame of the clip: counts for the max
        const bool kept = frame < keep;                 // survives the partial-segment cut: counts for the min
        const bool store = live && frame < p.n_frames_out;
        float* orow = p.out + static_cast<size_t>(clip) * NM * p.n_frames_out + frame;
        const size_t pitch = p.n_frames_out;
        float mx = -INFINITY, mn = INFINITY;
        auto emit = [&](int m, float acc) {
          const float L = __log2f(fmaxf(acc, 1e-10f)) * 0.301029995663981195f;
          mx = fmaxf(mx, L);
          mn = kept ? fminf(mn, L) : mn;
          if (store) orow[m * pitch] = L;
        };
        switch (warp) {
          case 0: mel_group<NM, 0>(P, emit); break;
          case 1: mel_group<NM, 1>(P, emit); break;
          case 2: mel_group<NM, 2>(P, emit); break;
          case 3: mel_group<NM, 3>(P, emit); break;
          default: mel_group<NM, 4>(P, emit); break;
        }
        if (!live) mx = -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (lane == 0) {
          ClipStat* st = p.stats + clip;
          if (mx > -INFINITY) atomicMax(&st->max_enc, enc_ordered(mx));
          if (mn < INFINITY) atomicMax(&st->min_inv, ~enc_ordered(mn));
        }
      }
    } else {
      // pad-only tile (n_frames_out > n_frames): nothing to compute
      if (tid == 0) sm_ctl[kCtlNext] = nxt_claim;
      __syncthreads();
      nxt = sm_ctl[kCtlNext];
      prefetched = false;
    }
    __syncthreads();  // all stores / atomics of the tile issued; power tile free

    if (tid == 0) {
      __threadfence();
      atomicAdd(&p.stats[clip].done, 1u);
      if (n_ring == kMaxPending) {  // park the oldest: never wait while tiles are unclaimed
        p.next[ring[0]] = chain;
        chain = ring[0];
#pragma unroll
        for (int k = 1; k < kMaxPending; ++k) ring[k - 1] = ring[k];
        --n_ring;
      }
      ring[n_ring++] = cur;
      int n_ready = 0, w = 0;
#pragma unroll
      for (int k = 0; k < kMaxPending; ++k) {
        if (k < n_ring) {
          const int t = ring[k];
          const uint32_t d = *reinterpret_cast<volatile uint32_t*>(&p.stats[t / p.tiles_per_clip].done);
          if (d >= static_cast<uint32_t>(p.tiles_per_clip)) sm_ctl[kCtlList + n_ready++] = t;
          else ring[w++] = t;
        }
      }
      n_ring = w;
      sm_ctl[kCtlReady] = n_ready;
      if (n_ready > 0) __threadfence();
    }
    __syncthreads();
    const int n_ready = sm_ctl[kCtlReady];
    for (int r = 0; r < n_ready; ++r) fixup_tile<NM>(p, sm_ctl[kCtlList + r], tid);
    cur = nxt;
  }

  // drain: every tile is claimed by a running CTA now, so waiting on a clip's counter is safe
  for (;;) {
    __syncthreads();  // previous readers of sm_ctl[kCtlDrain] / sm_ctl[kCtlList] are done
    if (tid == 0) {
      int t = -1;
      if (n_ring > 0) t = ring[--n_ring];
      else if (chain >= 0) { t = chain; chain = p.next[chain]; }
      if (t >= 0) {
        volatile uint32_t* d = reinterpret_cast<volatile uint32_t*>(&p.stats[t / p.tiles_per_clip].done);
        while (*d < static_cast<uint32_t>(p.tiles_per_clip)) __nanosleep(100);
        __threadfence();
      }
      sm_ctl[kCtlDrain] = t;
    }
    __syncthreads();
    const int t = sm_ctl[kCtlDrain];
    if (t < 0) break;
    fixup_tile<NM>(p, t, tid);
  }
}

}  // namespace wft
