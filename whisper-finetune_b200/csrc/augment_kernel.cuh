// Augmentation epilogue of the front end for sm_100a: time-warp -> time mask -> frequency mask -> extremes mask in ONE read and ONE
// write of the features (src/whisper_finetune/data/data_loader.py:284-290; TimeWarpAugmenter / ExtremesFrequencyMasking,
// data/utils.py:41-190), optionally finishing the cells the front-end kernel could not (wft_frontend_augment_forward).
//
// Two kernels compute the same function, bit for bit:
//   augment_staged_kernel  the production instance (n_frames % 4 == 0, 16-byte aligned tensors).  A CTA owns 1024 output frames
//                          x 8 rows; the window of source columns its taps fall into travels global -> shared memory as ONE
//                          bulk copy (TMA, mbarrier completion) per source row, all rows of the CTA in flight at once; the taps
//                          are then shared-memory loads at immediate offsets.
//   augment_kernel         any shape / alignment: taps straight from global memory (L1).
// Both keep ONE source column per frame: the bilinear taps are the in-row pair e, e + 1 with the weights re-assigned at the row
// ends, so that no load is conditional and the second one sits at an immediate offset from the first.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "frontend_kernel.cuh"

namespace wft {

// counter-based draw of (warp_p, warp_d): warp_p uniform in [W, T-W), warp_d uniform in [-W, W) (the reference's randint
// ranges, data/utils.py:107-111), Philox block 2 of the clip's counter; (-1, 0) == "no warp" when the p gate rejects
__device__ __noinline__ int2 draw_warp_point(uint64_t seed, uint64_t idx, int32_t n_frames, int32_t W, float p) {
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const uint32_t lo = static_cast<uint32_t>(idx), hi = static_cast<uint32_t>(idx >> 32);
  bool apply = p >= 1.0f;
  if (!apply && p > 0.0f) {
    uint32_t g[4];
    philox4x32_10(lo, hi, 1u, 0u, k0, k1, g);
    apply = u01(g[0]) < p;
  }
  int2 w = make_int2(-1, 0);   // "no warp": the warp kernels copy such a clip
  if (apply && W > 0 && n_frames > 2 * W) {
    uint32_t r[4];
    philox4x32_10(lo, hi, 2u, 0u, k0, k1, r);
    w.x = W + static_cast<int>(__fmul_rn(u01(r[0]), static_cast<float>(n_frames - 2 * W)));
    w.y = -W + static_cast<int>(__fmul_rn(u01(r[1]), static_cast<float>(2 * W)));
  }
  return w;
}

__global__ void time_warp_draw_kernel(uint64_t seed, uint64_t clip_offset, int32_t batch, int32_t n_frames, int32_t W,
                                      float p, int32_t* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  reinterpret_cast<int2*>(out)[b] = draw_warp_point(seed, clip_offset + static_cast<uint64_t>(b), n_frames, W, p);
}

// the augmentation epilogue may draw its clip's parameters itself (wft_augment_drawn_f32): same draws as wft_specaug_draw and
// wft_time_warp_draw for (seed, clip_offset + b)
struct AugDraw {
  int32_t enabled, tparam, fparam, W;
  float p;
  uint64_t seed, clip_offset;
};

// ---- fused augmentation epilogue: time-warp -> time mask -> frequency mask -> extremes mask in ONE read + write of the
// features (data_loader.py:284-290: time_warping, time_masking, freq_masking, extreme_freq_masking).  Every step after the
// warp only overwrites cells with the mask value, so out[b, r, t] = masked(b, r, t) ? mask_value : warp(in[b])[r, t].
//
// Source coordinate of output frame t (normalised, align_corners): the reference's 3-knot cubic Hermite spline
// (data/utils.py:65-93).  kF32 = false evaluates it in float64 and rounds once; kF32 = true restates the reference's own
// float32 evaluation order (knot slopes, (xs - x0) / dx, powers of t, the 4x4 basis product as a k-ascending FMA chain, the
// four products summed left to right) so that the coordinate lands on the reference's float32 value wherever torch's pow
// returns the correctly rounded power.
template <bool kF32>
__device__ __forceinline__ float warp_source_coord(int t, int T, int warp_p, int warp_d) {
  if (kF32) {
    const float y0 = -1.0f, y2 = 1.0f;
    const float y1 = __fsub_rn(__fdiv_rn(static_cast<float>((warp_p - warp_d) * 2), static_cast<float>(T - 1)), 1.0f);
    const float dxa = static_cast<float>(warp_p), dxb = static_cast<float>(T - 1 - warp_p);
    const float s0 = __fdiv_rn(__fsub_rn(y1, y0), dxa), s1 = __fdiv_rn(__fsub_rn(y2, y1), dxb);
    const float mm = __fdiv_rn(__fadd_rn(s1, s0), 2.0f);
    const bool second = t > warp_p;
    const float xa = second ? static_cast<float>(warp_p) : 0.0f, dx = second ? dxb : dxa;
    const float ya = second ? y1 : y0, yb = second ? y2 : y1, ma = second ? mm : s0, mb = second ? s1 : mm;
    const float u = __fdiv_rn(__fsub_rn(static_cast<float>(t), xa), dx);
    const float u2 = __fmul_rn(u, u);
    const float u3 = static_cast<float>(static_cast<double>(u) * static_cast<double>(u) * static_cast<double>(u));
    // A @ [1, u, u2, u3]^T, rows of A = (1,0,-3,2), (0,1,-2,1), (0,0,3,-2), (0,0,-1,1)
    const float h0 = __fmaf_rn(2.0f, u3, __fmaf_rn(-3.0f, u2, 1.0f));
    const float h1 = __fmaf_rn(1.0f, u3, __fmaf_rn(-2.0f, u2, u));
    const float h2 = __fmaf_rn(-2.0f, u3, __fmul_rn(3.0f, u2));
    const float h3 = __fmaf_rn(1.0f, u3, __fmul_rn(-1.0f, u2));
    float g = __fmul_rn(h0, ya);
    g = __fadd_rn(g, __fmul_rn(__fmul_rn(h1, ma), dx));
    g = __fadd_rn(g, __fmul_rn(h2, yb));
    g = __fadd_rn(g, __fmul_rn(__fmul_rn(h3, mb), dx));
    return g;
  } else {
    const double x1 = static_cast<double>(warp_p), x2 = static_cast<double>(T - 1);
    const double y0 = -1.0, y1 = static_cast<double>(warp_p - warp_d) * 2.0 / (T - 1.0) - 1.0, y2 = 1.0;
    const double s0 = (y1 - y0) / x1, s1 = (y2 - y1) / (x2 - x1);
    const double m0 = s0, m1 = 0.5 * (s0 + s1), m2 = s1;
    const bool second = static_cast<double>(t) > x1;
    const double xa = second ? x1 : 0.0, dx = second ? (x2 - x1) : x1;
    const double ya = second ? y1 : y0, yb = second ? y2 : y1, ma = second ? m1 : m0, mb = second ? m2 : m1;
    const double u = (static_cast<double>(t) - xa) / dx, u2 = u * u, u3 = u2 * u;
    return static_cast<float>((1.0 - 3.0 * u2 + 2.0 * u3) * ya + (u - 2.0 * u2 + u3) * ma * dx + (3.0 * u2 - 2.0 * u3) * yb +
                              (-u2 + u3) * mb * dx);
  }
}

// The float64 spline once per CTA instead of once per frame: the source map is a cubic in u = (t - xa) / dx on each of its two
// segments, g(u) = A h00 + B h10 + C h01 + D h11 with A = ya, B = ma dx, C = yb, D = mb dx, i.e.
//   g(u) = A + B u + (-3A - 2B + 3C - D) u^2 + (2A + B - 2C + D) u^3.
// One thread derives {xa, 1 / dx, c0 .. c3} for both segments (the three float64 divisions of the knot slopes live here: with
// every thread evaluating its own frames they were half of the kernel's instructions), every frame is then 1 multiply + 3 FMAs.
__device__ __forceinline__ void spline_segments(int T, int warp_p, int warp_d, double* __restrict__ seg /* [2][6] */) {
#if defined(WFT_AUG_KO) && (WFT_AUG_KO & 1)   // knock-out build (timing only): no float64 divisions, the identity map
  for (int k = 0; k < 2; ++k) { seg[6 * k] = 0.0; seg[6 * k + 1] = 1.0 / 2999.0; seg[6 * k + 2] = -1.0; seg[6 * k + 3] = 2.0; seg[6 * k + 4] = 0.0; seg[6 * k + 5] = 0.0; }
  return;
#endif
  const double x1 = static_cast<double>(warp_p), x2 = static_cast<double>(T - 1);
  const double y0 = -1.0, y1 = static_cast<double>(warp_p - warp_d) * 2.0 / (T - 1.0) - 1.0, y2 = 1.0;
  const double s0 = (y1 - y0) / x1, s1 = (y2 - y1) / (x2 - x1);
  const double m0 = s0, m1 = 0.5 * (s0 + s1), m2 = s1;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double xa = k ? x1 : 0.0, dx = k ? (x2 - x1) : x1;
    const double A = k ? y1 : y0, C = k ? y2 : y1, B = (k ? m1 : m0) * dx, D = (k ? m2 : m1) * dx;
    seg[6 * k + 0] = xa;
    seg[6 * k + 1] = 1.0 / dx;
    seg[6 * k + 2] = A;
    seg[6 * k + 3] = B;
    seg[6 * k + 4] = -3.0 * A - 2.0 * B + 3.0 * C - D;
    seg[6 * k + 5] = 2.0 * A + B - 2.0 * C + D;
  }
}
__device__ __forceinline__ float spline_eval(int t, int warp_p, const double* __restrict__ seg) {
  const double* c = seg + (t > warp_p ? 6 : 0);
  const double u = (static_cast<double>(t) - c[0]) * c[1];
  return static_cast<float>(fma(fma(fma(c[5], u, c[4]), u, c[3]), u, c[2]));
}

constexpr int kAugThreads = 256;
constexpr int kAugFramesPerThread = 4;
constexpr int kAugRowsPerCta = 16;

// kFix instances of the epilogue run directly behind a front-end grid that was launched WITHOUT its fix-up grid
// (wft_frontend_augment_forward): `in` then holds what the front-end kernel wrote -- final features except for what can only
// be finished once the whole clip is known -- and every tap is finished on load exactly like fixup_tile would have
// rewritten it: max(v, floor) for the kept frames, the clamp value for tiles that were never computed (silent / pad-only),
// the min-value pad beyond the kept frames (data/utils.py:380-404).
struct AugFix {
  const ClipStat* stats;    // this call's clip statistics (complete once the front-end grid is)
  const int32_t* lengths;
  const int32_t* n_valid;
  int32_t n_samples, n_total, n_frames;   // of the front-end call (frames the clip really has; T is n_frames_out)
};

// base + 4 * idx (the compiler, left to itself, widens, adds and scales in four instructions per load)
__device__ __forceinline__ const float* elem_ptr(const float* base, uint32_t idx) {
  uint64_t a;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(a) : "r"(idx), "l"(reinterpret_cast<uint64_t>(base)));
  return reinterpret_cast<const float*>(a);
}

// the value fixup_tile would have left in a cell the front-end kernel wrote as v (kind: 0 = computed, 1 = never computed, 2 = pad)
// (s_fix: floor feature, pad value, kept frames, clip length, clamp feature; the last two values a ragged clip needs are read
// from shared memory where they are used -- a register each would spill the float64-spline instance)
__device__ __forceinline__ float aug_finish(float v, uint32_t kind, bool ragged, float floorn, const int* __restrict__ s_fix) {
  if (!ragged) return fmaxf(v, floorn);
  return kind == 2u ? __int_as_float(s_fix[1]) : fmaxf(kind == 1u ? __int_as_float(s_fix[4]) : v, floorn);
}

// What both kernels do before their first __syncthreads: the CTA's row table, the clip's parameters (drawn before the wait
// for the producing grid, loaded after it) and, for the kFix instances, what finishing a cell needs.
// kFix instances: what finishing a cell needs.  The front-end grid is complete once griddepcontrol.wait has returned, so its
// statistics are final.  Load and publication are separate so that a kernel can keep the loads in flight across a barrier.
__device__ __forceinline__ int4 aug_fix_load(int b, const AugFix& fix) {
  int4 raw;
  raw.x = static_cast<int>(__ldcg(&fix.stats[b].max_enc));
  raw.y = static_cast<int>(__ldcg(&fix.stats[b].min_inv));
  raw.z = fix.lengths != nullptr ? __ldg(fix.lengths + b) : fix.n_samples;
  raw.w = fix.n_valid != nullptr ? __ldg(fix.n_valid + b) : -1;
  return raw;
}
__device__ __forceinline__ void aug_fix_publish(int4 raw, const AugFix& fix, int* __restrict__ s_fix) {
  const float floorn = floor_feature(dec_ordered(static_cast<uint32_t>(raw.x)));
  const float padv = fmaxf(feature_of_l2(dec_ordered(~static_cast<uint32_t>(raw.y))), floorn);
  const int len = raw.z < 0 ? 0 : (raw.z < fix.n_samples ? raw.z : fix.n_samples);
  s_fix[0] = __float_as_int(floorn);
  s_fix[1] = __float_as_int(padv);
  s_fix[2] = (raw.w >= 0 && raw.w < fix.n_frames) ? raw.w : fix.n_frames;   // kept_frames()
  s_fix[3] = len;
  s_fix[4] = __float_as_int(feature_of_l2(silent_l2()));
}

template <bool kF32, int kFix, bool kPublishFix, int kRows>
__device__ __forceinline__ void aug_prologue(int b, int32_t R, int32_t T, const int32_t* __restrict__ warp_params,
                                             const int32_t* __restrict__ mask_params, const AugDraw& draw, const AugFix& fix,
                                             int* __restrict__ s_draw, int* __restrict__ s_fix, double* __restrict__ s_seg,
                                             int4* __restrict__ s_row) {
  // source row(s) of every output row of the group: grid_sample's y coordinate of row r (the identity up to float32 rounding,
  // which can put a sliver of weight on the next row -- restated, not assumed), once per CTA instead of once per thread and row
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kRows) {
    const int r = blockIdx.y * kRows + (threadIdx.x - 64);
    const float step = 2.0f / static_cast<float>(R - 1);  // torch.linspace(-1, 1, R)
    const float gy = (r < R / 2) ? (-1.0f + step * static_cast<float>(r)) : (1.0f - step * static_cast<float>(R - 1 - r));
    const float iy = ((gy + 1.0f) / 2.0f) * static_cast<float>(R - 1);
    const float iy0f = floorf(iy);
    const int iy0 = static_cast<int>(iy0f), iy1 = iy0 + 1;
    const float wy1r = iy - iy0f, wy0r = (iy0f + 1.0f) - iy;
    const bool use0 = iy0 >= 0 && iy0 < R, use1 = iy1 >= 0 && iy1 < R && wy1r != 0.0f;
    s_row[threadIdx.x - 64] = make_int4(use0 ? iy0 : 0, __float_as_int(use0 ? wy0r : 0.0f), __float_as_int(use1 ? wy1r : 0.0f),
                                        use1 ? iy1 : -1);
  }
  // clip parameters -> shared memory: thread 0 the mask intervals, thread 32 the warp point and the spline's segment
  // coefficients.  Drawn parameters depend on nothing a grid in front produced: they are ready before this grid's wait.
  if (draw.enabled) {
    if (threadIdx.x == 0) {
      const int4 m = draw_mask_intervals(draw.seed, draw.clip_offset + static_cast<uint64_t>(b), R, T, draw.tparam, draw.fparam, draw.p);
      s_draw[0] = m.x; s_draw[1] = m.y; s_draw[2] = m.z; s_draw[3] = m.w;
    } else if (threadIdx.x == 32) {
      const int2 w = draw_warp_point(draw.seed, draw.clip_offset + static_cast<uint64_t>(b), T, draw.W, draw.p);
      s_draw[4] = w.x; s_draw[5] = w.y;
      if (!kF32 && w.x > 0 && w.x < T - 1) spline_segments(T, w.x, w.y, s_seg);
    }
  }
  // a programmatic dependent of whatever produced `in`: the grid behind this one may be scheduled, this one waits
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!draw.enabled) {
    if (threadIdx.x == 0) {
      int4 m = make_int4(0, 0, 0, 0);
      if (mask_params != nullptr) m = __ldg(reinterpret_cast<const int4*>(mask_params) + b);
      s_draw[0] = m.x; s_draw[1] = m.y; s_draw[2] = m.z; s_draw[3] = m.w;
    } else if (threadIdx.x == 32) {
      int2 w = make_int2(-1, 0);
      if (warp_params != nullptr) w = __ldg(reinterpret_cast<const int2*>(warp_params) + b);
      s_draw[4] = w.x; s_draw[5] = w.y;
      if (!kF32 && w.x > 0 && w.x < T - 1) spline_segments(T, w.x, w.y, s_seg);
    }
  }
  if (kFix != 0 && kPublishFix && threadIdx.x == 96) aug_fix_publish(aug_fix_load(b, fix), fix, s_fix);
}

// ---- the generic instance -------------------------------------------------------------------------------------------------
// grid = (frame blocks of 1024, row groups of 16, clips); a thread owns 4 output frames, 256 apart (lane <-> consecutive frames:
// a warp's store is one 128-byte line and the two source taps of a smooth map fall into one or two lines -- with 4 ADJACENT
// frames per thread every scalar load of a warp was spread over 4-8 lines); it evaluates the 4 source coordinates once and
// walks the 16 rows of its group with 8 independent loads in flight per row (bilinear taps mirror grid_sample's float32
// arithmetic, zeros outside).
// kFix: 0 = `in` holds finished features; 1 = finish on load, full-length clips without a cut (the floor is all there is);
// 2 = finish on load, ragged batch (lengths / cuts / output longer than the clip: per-tap kinds; 64 registers, 4 CTAs per SM)
template <bool kF32, int kFix>
__global__ void __launch_bounds__(kAugThreads, kFix == 2 ? 4 : 5) augment_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             int32_t R, int32_t T, const int32_t* __restrict__ warp_params,
                                                             const int32_t* __restrict__ mask_params,
                                                             const int32_t* __restrict__ extremes, float mask_value,
                                                             const AugDraw draw, const AugFix fix) {
  const int b = blockIdx.z;
  __shared__ int s_draw[8];
  __shared__ int s_fix[5];    // kFix: floor feature, pad value (float bits), kept frames, clip length in samples, clamp feature
  __shared__ double s_seg[12];
  __shared__ int4 s_row[kAugRowsPerCta];   // per output row of this CTA: source row, its weight, the next row's weight (bits), -
  aug_prologue<kF32, kFix, true, kAugRowsPerCta>(b, R, T, warp_params, mask_params, draw, fix, s_draw, s_fix, s_seg, s_row);
  __syncthreads();
  const int tbase = blockIdx.x * kAugThreads * kAugFramesPerThread + threadIdx.x;
  constexpr int kStep = kAugThreads;   // frame k of this thread = tbase + k * kStep
  if (tbase >= T) return;
  const int wp = s_draw[4], wd = s_draw[5];
  const int4 mk = make_int4(s_draw[0], s_draw[1], s_draw[2], s_draw[3]);
  // the CTA's 16 rows against the frequency mask and the extremes mask, once: bit i = row blockIdx.y * 16 + i is masked
  uint32_t rowbits = 0;
  {
    int lo_rows = 0, hi_rows = 0;
    if (extremes != nullptr) {
      const int2 e = __ldg(reinterpret_cast<const int2*>(extremes) + b);
      lo_rows = e.x; hi_rows = e.y;
    }
    const int r0 = blockIdx.y * kAugRowsPerCta;   // (block scope; the same value again below)
    auto rows = [r0](int a, int e) {   // rows [a, e) as bits of this CTA's group
      a = min(max(a - r0, 0), kAugRowsPerCta);
      e = min(max(e - r0, 0), kAugRowsPerCta);
      return e > a ? (1u << e) - (1u << a) : 0u;
    };
    rowbits = rows(mk.z, mk.w) | rows(0, lo_rows) | rows(R - hi_rows, R);
  }
  const bool warp = wp > 0 && wp < T - 1;          // anything else (the draw's "gate rejected" marker is -1) = no warp
  const int r0 = blockIdx.y * kAugRowsPerCta;
  const size_t origin = (static_cast<size_t>(b) * R + r0) * T;   // cell (b, r0, 0)
  // Per frame, once: the left one of two neighbouring source columns e, e + 1, their weights, and whether the cell is live.
  // The taps of grid_sample are columns a = floor(ix) and a + 1 with "zeros" padding; here they are always the in-row pair
  // e = clamp(a, 0, T - 2), e + 1 -- both loads unconditional, the second one at an immediate offset -- and a tap that falls
  // outside gets weight 0 on a finite cell (0 * finite adds exactly nothing; at a = -1 / a = T - 1 the one live tap keeps its
  // weight and the two products swap places in the sum, which changes no bit: x + 0 = 0 + x).  A row then costs one uniform
  // row pointer and one address per frame for its eight loads (ncu on the version with two clamped columns per frame: 28 M
  // warp-instructions per B = 64 launch, 36 per output cell, three quarters of them 64-bit address arithmetic).
  uint32_t e[kAugFramesPerThread];
  float wa[kAugFramesPerThread], wc[kAugFramesPerThread];
  bool on[kAugFramesPerThread];
#pragma unroll
  for (int k = 0; k < kAugFramesPerThread; ++k) {
    const int t = tbase + k * kStep;
    on[k] = t < T && !(t >= mk.x && t < mk.y);
    int a = t < T ? t : T - 1;
    float wx0 = 1.0f, wx1 = 0.0f;
    if (warp && t < T) {
      const float gx = kF32 ? warp_source_coord<true>(t, T, wp, wd) : spline_eval(t, wp, s_seg);
      const float ix = ((gx + 1.0f) / 2.0f) * static_cast<float>(T - 1);
      const float f = floorf(ix);
      a = static_cast<int>(f);
      wx1 = ix - f;
      wx0 = (f + 1.0f) - ix;
    }
    const int ek = min(max(a, 0), T - 2);
    // weight of column e: tap a if a == e, tap a + 1 if a + 1 == e (a == -1); of column e + 1: tap a + 1 if a == e, tap a if a == e + 1
    wa[k] = a == ek ? wx0 : (a + 1 == ek ? wx1 : 0.0f);
    wc[k] = a == ek ? wx1 : (a == ek + 1 ? wx0 : 0.0f);
    e[k] = static_cast<uint32_t>(ek);
  }
  // kFix: what a tap at source column e / e + 1 still needs (per frame, once): 0 = floor only, 1 = never computed (the clamp
  // value, then the floor), 2 = beyond the kept frames (the pad value).  Full-length clips without a cut need no table.
  // (Without a warp the one tap is column t itself.)
  float floorn = 0.0f;
  bool ragged = false;
  uint32_t kinds = 0;   // 2 bits per tap: left tap of frame k at bit 4k, right tap at bit 4k + 2
  if constexpr (kFix != 0) floorn = __int_as_float(s_fix[0]);
  if constexpr (kFix == 2) {
    const int keep = s_fix[2], len = s_fix[3];
    ragged = keep < T || len < fix.n_samples || fix.n_frames < T;
    if (ragged) {
#pragma unroll
      for (int k = 0; k < kAugFramesPerThread; ++k) {
        const int a = warp ? static_cast<int>(e[k]) : min(tbase + k * kStep, T - 1), c = min(a + 1, T - 1);
        const uint32_t ka = a >= keep ? 2u : (tile_is_silent(a & ~(kTileFrames - 1), len, fix.n_total) ? 1u : 0u);
        const uint32_t kc = c >= keep ? 2u : (tile_is_silent(c & ~(kTileFrames - 1), len, fix.n_total) ? 1u : 0u);
        kinds |= (ka | (kc << 2)) << (4 * k);
      }
    }
  }
#define AUG_FINISH(v, k, tap) aug_finish(v, (kinds >> (4 * (k) + 2 * (tap))) & 3u, ragged, floorn, s_fix)
  const int n_rows = min(R - r0, kAugRowsPerCta);
  const float* src = in + origin + tbase;    // (no warp) this thread's cells of row r0; one pointer, frames at immediate offsets
  float* dst = out + origin + tbase;
  for (int i = 0; i < n_rows; ++i, src += T, dst += T) {
    asm volatile("" : "+l"(src), "+l"(dst));   // carried in registers: re-deriving them from %tid costs ten instructions a row
    float v[kAugFramesPerThread];
    const bool rowmask = ((rowbits >> i) & 1u) != 0u;
    if (rowmask) {
#pragma unroll
      for (int k = 0; k < kAugFramesPerThread; ++k) v[k] = mask_value;
    } else if (!warp) {
#pragma unroll
      for (int k = 0; k < kAugFramesPerThread; ++k) {
        if constexpr (kFix != 0) v[k] = on[k] ? AUG_FINISH(__ldg(src + k * kStep), k, 0) : mask_value;
        else v[k] = on[k] ? __ldg(src + k * kStep) : mask_value;
      }
    } else {
      const int4 rr = s_row[i];
      const float wy0 = __int_as_float(rr.y), wy1 = __int_as_float(rr.z);
      const bool use1 = rr.w >= 0;
      const float* row0 = in + origin + (rr.x - r0) * T;    // source row (warp-uniform)
      // every tap of the row is requested before the first one is used (8 independent loads in flight per thread)
      float t0a[kAugFramesPerThread], t0c[kAugFramesPerThread];
#pragma unroll
      for (int k = 0; k < kAugFramesPerThread; ++k) {
        const float* q = elem_ptr(row0, e[k]);
        t0a[k] = __ldg(q);
        t0c[k] = __ldg(q + 1);
      }
      if constexpr (kFix != 0) {
#pragma unroll
        for (int k = 0; k < kAugFramesPerThread; ++k) {
          t0a[k] = AUG_FINISH(t0a[k], k, 0);
          t0c[k] = AUG_FINISH(t0c[k], k, 1);
        }
      }
      if (!use1) {      // warp-uniform (depends on the row alone); taps accumulate in grid_sample's order
#pragma unroll
        for (int k = 0; k < kAugFramesPerThread; ++k) {
          float acc = 0.0f;
          acc += t0a[k] * (wa[k] * wy0);
          acc += t0c[k] * (wc[k] * wy0);
          v[k] = on[k] ? acc : mask_value;
        }
      } else {
        const float* row1 = in + origin + (rr.w - r0) * T;
        float t1a[kAugFramesPerThread], t1c[kAugFramesPerThread];
#pragma unroll
        for (int k = 0; k < kAugFramesPerThread; ++k) {
          const float* q = elem_ptr(row1, e[k]);
          t1a[k] = __ldg(q);
          t1c[k] = __ldg(q + 1);
        }
        if constexpr (kFix != 0) {
#pragma unroll
          for (int k = 0; k < kAugFramesPerThread; ++k) {
            t1a[k] = AUG_FINISH(t1a[k], k, 0);
            t1c[k] = AUG_FINISH(t1c[k], k, 1);
          }
        }
#pragma unroll
        for (int k = 0; k < kAugFramesPerThread; ++k) {
          float acc = 0.0f;
          acc += t0a[k] * (wa[k] * wy0);
          acc += t0c[k] * (wc[k] * wy0);
          acc += t1a[k] * (wa[k] * wy1);
          acc += t1c[k] * (wc[k] * wy1);
          v[k] = on[k] ? acc : mask_value;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kAugFramesPerThread; ++k)
      if (tbase + k * kStep < T) dst[k * kStep] = v[k];
  }
}

#undef AUG_FINISH

// ---- the staged instance ------------------------------------------------------------------------------------------------
// ncu on augment_kernel (B = 64): 28 M warp-instructions for 24.6 M cells, three quarters of them 64-bit address arithmetic for
// 4-byte loads, long-scoreboard stalls on top: 47 us for 196.6 MB, and the same 47 us with the stores removed.  Here the loads
// are not instructions at all.  A CTA owns 1024 output frames x 8 rows.  Once its threads know their source columns (block
// min / max: the window is exact, whatever the spline does), warp 0 sends the window of every source row the CTA needs global
// -> shared memory as bulk copies (TMA; one mbarrier counts the bytes of all of them, ~40 KB in flight per CTA, 4 CTAs per SM);
// a thread's taps are then two 4-byte shared-memory loads per cell at an immediate offset from ONE per-frame register, and the
// row loop is 2 LDS + 2 FMUL + 2 FFMA + a select + a store per cell.  A window wider than the buffer (output frames on a steep
// piece of the map, ~2 % of all CTAs) and a clip without a warp take the global path inside the same kernel.
constexpr int kStgThreads = 256;
constexpr int kStgFrames = 4;                          // frames per thread, 256 apart
constexpr int kStgBlock = kStgThreads * kStgFrames;    // 1024 output frames per CTA
constexpr int kStgRows = 8;                            // output rows per CTA
constexpr int kStgSrcRows = kStgRows + 2;              // source rows a row group can touch: r0 - 1 .. r0 + 8
constexpr int kStgCols = 1216;                         // staged source columns per row (multiple of 4)
constexpr int kStgSmemBytes = kStgSrcRows * kStgCols * 4;   // 48 640 B dynamic

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32_4(uint32_t addr) {   // the neighbouring column
  float v;
  asm volatile("ld.shared.f32 %0, [%1 + 4];" : "=f"(v) : "r"(addr));
  return v;
}

template <bool kF32, int kFix>
__global__ void __launch_bounds__(kStgThreads, 4) augment_staged_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                    int32_t R, int32_t T, const int32_t* __restrict__ warp_params,
                                                                    const int32_t* __restrict__ mask_params,
                                                                    const int32_t* __restrict__ extremes, float mask_value,
                                                                    const AugDraw draw, const AugFix fix) {
  extern __shared__ __align__(16) float s_buf[];   // [kStgSrcRows][kStgCols]
  const int b = blockIdx.z;
  __shared__ int s_draw[8];
  __shared__ int s_fix[5];
  __shared__ double s_seg[12];
  __shared__ int4 s_row[kStgRows];
  __shared__ int s_win[2];                       // min / max left source column over the CTA's live frames
  __shared__ __align__(8) uint64_t s_bar;        // completion of all the CTA's bulk copies
  const int tid = threadIdx.x;
  if (tid == 128) {
    s_win[0] = 0x7fffffff;
    s_win[1] = -1;
    mbar_init(&s_bar, 1);
  }
  aug_prologue<kF32, kFix, false, kStgRows>(b, R, T, warp_params, mask_params, draw, fix, s_draw, s_fix, s_seg, s_row);
  // (thread 96: the loads for s_fix are requested here and used behind the second barrier -- nothing in front of the row loop
  // needs them, and a round trip to L2 would otherwise sit on every CTA's way to its bulk copies)
  int4 fix_raw = make_int4(0, 0, 0, 0);
  if (kFix != 0 && tid == 96) fix_raw = aug_fix_load(b, fix);
  __syncthreads();
  const int tbase = blockIdx.x * kStgBlock + tid;
  constexpr int kStep = kStgThreads;
  const int wp = s_draw[4], wd = s_draw[5];
  const int4 mk = make_int4(s_draw[0], s_draw[1], s_draw[2], s_draw[3]);
  const int r0 = blockIdx.y * kStgRows;
  uint32_t rowbits = 0;
  {
    int lo_rows = 0, hi_rows = 0;
    if (extremes != nullptr) {
      const int2 ex = __ldg(reinterpret_cast<const int2*>(extremes) + b);
      lo_rows = ex.x; hi_rows = ex.y;
    }
    auto rows = [r0](int a, int e) {
      a = min(max(a - r0, 0), kStgRows);
      e = min(max(e - r0, 0), kStgRows);
      return e > a ? (1u << e) - (1u << a) : 0u;
    };
    rowbits = rows(mk.z, mk.w) | rows(0, lo_rows) | rows(R - hi_rows, R);
  }
  const bool warp = wp > 0 && wp < T - 1;
  const size_t origin = (static_cast<size_t>(b) * R + r0) * T;
  // per frame, once (see augment_kernel): left source column e, the weights of columns e and e + 1, whether the cell is live
  uint32_t e[kStgFrames];
  float wa[kStgFrames], wc[kStgFrames];
  bool on[kStgFrames];
  int emin = 0x7fffffff, emax = -1;
#pragma unroll
  for (int k = 0; k < kStgFrames; ++k) {
    const int t = tbase + k * kStep;
    on[k] = t < T && !(t >= mk.x && t < mk.y);
    int a = t < T ? t : T - 1;
    float wx0 = 1.0f, wx1 = 0.0f;
    if (warp && t < T) {
      const float gx = kF32 ? warp_source_coord<true>(t, T, wp, wd) : spline_eval(t, wp, s_seg);
      const float ix = ((gx + 1.0f) / 2.0f) * static_cast<float>(T - 1);
      const float f = floorf(ix);
      a = static_cast<int>(f);
      wx1 = ix - f;
      wx0 = (f + 1.0f) - ix;
    }
    const int ek = min(max(a, 0), T - 2);
    wa[k] = a == ek ? wx0 : (a + 1 == ek ? wx1 : 0.0f);
    wc[k] = a == ek ? wx1 : (a == ek + 1 ? wx0 : 0.0f);
    e[k] = static_cast<uint32_t>(ek);
    if (t < T) {
      emin = min(emin, ek);
      emax = max(emax, ek);
    }
  }
  // the CTA's window of source columns: [c0, c0 + n), 16-byte aligned at both ends (T % 4 == 0 keeps it inside the row)
  if (warp) {
    emin = __reduce_min_sync(0xffffffffu, emin);
    emax = __reduce_max_sync(0xffffffffu, emax);
    if ((tid & 31) == 0) {
      atomicMin(&s_win[0], emin);
      atomicMax(&s_win[1], emax);
    }
  }
  __syncthreads();
  const int c0 = s_win[0] & ~3;
  const int n = (s_win[1] + 2 - c0 + 3) & ~3;
  const bool staged = warp && s_win[1] >= 0 && n <= kStgCols;
  const int smin = max(r0 - 1, 0);                       // first source row the group can touch
  const int n_rows = min(R - r0, kStgRows);
  if (staged && tid < 32) {
    // every source row a live output row of the group reads, as one bulk copy each; warp 0 finds the rows (lane <-> output row),
    // lane 0 announces the bytes of all of them and lane s sends source row s (one thread issuing them all took ~2 us of every
    // CTA's life)
    uint32_t bits = 0;
    if (tid < n_rows && !((rowbits >> tid) & 1u)) {
      const int4 rr = s_row[tid];
      bits = 1u << (rr.x - smin);
      if (rr.w >= 0) bits |= 1u << (rr.w - smin);
    }
    const uint32_t need = __reduce_or_sync(0xffffffffu, bits);
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&s_bar, static_cast<uint32_t>(__popc(need)) * static_cast<uint32_t>(n) * 4u);
    }
    __syncwarp();
    if (tid < kStgSrcRows && ((need >> tid) & 1u))
      tma_bulk_g2s(s_buf + tid * kStgCols, in + (static_cast<size_t>(b) * R + smin + tid) * T + c0, static_cast<uint32_t>(n) * 4u,
                   &s_bar);
  }
  if constexpr (kFix != 0) {
    if (tid == 96) aug_fix_publish(fix_raw, fix, s_fix);
    __syncthreads();
  }
  if (tbase >= T) return;
  float floorn = 0.0f;
  bool ragged = false;
  uint32_t kinds = 0;
  if constexpr (kFix != 0) floorn = __int_as_float(s_fix[0]);
  if constexpr (kFix == 2) {
    const int keep = s_fix[2], len = s_fix[3];
    ragged = keep < T || len < fix.n_samples || fix.n_frames < T;
    if (ragged) {
#pragma unroll
      for (int k = 0; k < kStgFrames; ++k) {
        const int a = warp ? static_cast<int>(e[k]) : min(tbase + k * kStep, T - 1), c = min(a + 1, T - 1);
        const uint32_t ka = a >= keep ? 2u : (tile_is_silent(a & ~(kTileFrames - 1), len, fix.n_total) ? 1u : 0u);
        const uint32_t kc = c >= keep ? 2u : (tile_is_silent(c & ~(kTileFrames - 1), len, fix.n_total) ? 1u : 0u);
        kinds |= (ka | (kc << 2)) << (4 * k);
      }
    }
  }
#define AUG_FINISH(v, k, tap) aug_finish(v, (kinds >> (4 * (k) + 2 * (tap))) & 3u, ragged, floorn, s_fix)
  // shared-memory address of this thread's left tap in staged row 0 (a frame beyond T is not part of the window: any valid
  // address will do)
  uint32_t lo[kStgFrames];
  const uint32_t buf0 = smem_u32(s_buf);
#pragma unroll
  for (int k = 0; k < kStgFrames; ++k)
    lo[k] = buf0 + ((staged && tbase + k * kStep < T) ? (e[k] - static_cast<uint32_t>(c0)) * 4u : 0u);
  if (staged) mbar_wait(&s_bar, 0u);            // all source rows of the CTA have landed
  float* dst = out + origin + tbase;
  for (int i = 0; i < n_rows; ++i, dst += T) {
    float v[kStgFrames];
    if ((rowbits >> i) & 1u) {
#pragma unroll
      for (int k = 0; k < kStgFrames; ++k) v[k] = mask_value;
    } else if (!warp) {
      const float* src = in + (dst - out);
#pragma unroll
      for (int k = 0; k < kStgFrames; ++k) {
        if constexpr (kFix != 0) v[k] = on[k] ? AUG_FINISH(__ldg(src + k * kStep), k, 0) : mask_value;
        else v[k] = on[k] ? __ldg(src + k * kStep) : mask_value;
      }
    } else {
      const int4 rr = s_row[i];
      const float wy0 = __int_as_float(rr.y), wy1 = __int_as_float(rr.z);
      const bool use1 = rr.w >= 0;
      // the taps of source row `which` (0: rr.x, 1: rr.w) for this thread's frames, finished if the instance finishes cells
      float ta[kStgFrames], tc[kStgFrames], acc[kStgFrames];
      auto taps = [&](int srow) {
        if (staged) {
          const uint32_t o = static_cast<uint32_t>(srow - smin) * (kStgCols * 4u);
#pragma unroll
          for (int k = 0; k < kStgFrames; ++k) {
            ta[k] = lds_f32(lo[k] + o);
            tc[k] = lds_f32_4(lo[k] + o);
          }
        } else {
          const float* row = in + origin + (srow - r0) * T;
#pragma unroll
          for (int k = 0; k < kStgFrames; ++k) {
            const float* q = elem_ptr(row, e[k]);
            ta[k] = __ldg(q);
            tc[k] = __ldg(q + 1);
          }
        }
        if constexpr (kFix != 0) {
#pragma unroll
          for (int k = 0; k < kStgFrames; ++k) {
            ta[k] = AUG_FINISH(ta[k], k, 0);
            tc[k] = AUG_FINISH(tc[k], k, 1);
          }
        }
      };
      taps(rr.x);
#pragma unroll
      for (int k = 0; k < kStgFrames; ++k) {   // taps accumulate in grid_sample's order
        acc[k] = 0.0f;
        acc[k] += ta[k] * (wa[k] * wy0);
        acc[k] += tc[k] * (wc[k] * wy0);
      }
      if (use1) {      // warp-uniform: a sliver of weight on the next row (float32 rounding of the row coordinate)
        taps(rr.w);
#pragma unroll
        for (int k = 0; k < kStgFrames; ++k) {
          acc[k] += ta[k] * (wa[k] * wy1);
          acc[k] += tc[k] * (wc[k] * wy1);
        }
      }
#pragma unroll
      for (int k = 0; k < kStgFrames; ++k) v[k] = on[k] ? acc[k] : mask_value;
    }
#pragma unroll
    for (int k = 0; k < kStgFrames; ++k) {
#if defined(WFT_AUG_KO) && (WFT_AUG_KO & 2)   // knock-out build (timing only): no stores
      if (tbase + k * kStep < T && v[k] != v[k]) dst[k * kStep] = v[k];
#else
      if (tbase + k * kStep < T) dst[k * kStep] = v[k];
#endif
    }
  }
#undef AUG_FINISH
}

}  // namespace wft
