// C-ABI shim of the B200-native Whisper front end (see include/wft.h for the contract and the reference
// file:line each entry point replaces).  Host code only validates, sizes the persistent grid and launches;
// all arithmetic lives in frontend_kernel.cuh and the three small kernels below.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "../../include/wft.h"
#include "frontend_kernel.cuh"

namespace {

thread_local std::string g_last_error;
thread_local int64_t g_launches = 0;
int g_debug_chunk = 0;      // WFT_DEBUG_CHUNK (development): force the tiles-per-claim of the fused kernel
int g_debug_max_ctas = 0;   // wft_debug_set_max_ctas: caps the persistent grids (results must not depend on the grid)
int g_debug_extra_smem = 0; // wft_debug_set_extra_smem: pads the front-end CTA's shared memory, i.e. lowers its CTAs per SM

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(WFT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define WFT_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
  } while (0)

struct GridInfo {
  int ctas = 0;
  int smem = 0;
  bool ready = false;
};

template <typename KernelT>
int grid_for(KernelT kernel, GridInfo* cache, int* ctas) {
  int dev = 0;
  WFT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(WFT_ERR_CUDA, "device ordinal out of range");
  GridInfo& gi = cache[dev];
  if (!gi.ready) {
    WFT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wft::kSmemBytes));
    int per_sm = 0, sms = 0;
    WFT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, wft::kThreads, wft::kSmemBytes));
    WFT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm < 1) return fail(WFT_ERR_CUDA, "front-end kernel does not fit on this device");
    gi.ctas = per_sm * sms;
    gi.smem = wft::kSmemBytes;
    gi.ready = true;
  }
  *ctas = gi.ctas;
  return WFT_OK;
}

template <int NM, typename PcmT>
int launch_frontend(const wft::FrontendParams& p, cudaStream_t stream, int launch_flags, bool with_fixup) {
  static GridInfo cache[64];
  static const bool env_read = [] {
    if (const char* e = getenv("WFT_DEBUG_CHUNK")) g_debug_chunk = atoi(e);
    return true;
  }();
  (void)env_read;
  int ctas = 0;
  int rc = grid_for(wft::frontend_kernel<NM, PcmT>, cache, &ctas);
  if (rc != WFT_OK) return rc;
  if (g_debug_extra_smem > 0) {   // fewer front-end CTAs per SM: the rest of the SM is left to whatever runs next to this grid
    int per_sm = 0, sms = 0, dev = 0;
    WFT_CUDA(cudaGetDevice(&dev));
    WFT_CUDA(cudaFuncSetAttribute(wft::frontend_kernel<NM, PcmT>, cudaFuncAttributeMaxDynamicSharedMemorySize, wft::kSmemBytes + g_debug_extra_smem));
    WFT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wft::frontend_kernel<NM, PcmT>, wft::kThreads, wft::kSmemBytes + g_debug_extra_smem));
    WFT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm >= 1) ctas = per_sm * sms;
  }
  const int ctas_full = ctas;
  if (ctas > p.total_tiles) ctas = p.total_tiles;
  if (g_debug_max_ctas > 0 && ctas > g_debug_max_ctas) ctas = g_debug_max_ctas;
  // tiles a CTA takes per claim: at most 1/24 of its share (measured: chunks of 8 at 54 tiles per CTA cost 5 % in the tail)
  wft::FrontendParams q = p;
  q.overlap = (launch_flags & WFT_LAUNCH_OVERLAP) ? 1 : 0;
  const int per_cta = p.total_tiles / ctas;
  q.chunk = g_debug_chunk > 0 ? g_debug_chunk : (per_cta / 24 < 1 ? 1 : (per_cta / 24 > 4 ? 4 : per_cta / 24));
  // programmatic stream serialization: this grid may be scheduled while the previous kernel on the stream drains; the
  // kernel itself waits (griddepcontrol.wait) before it touches anything the previous one produced
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(wft::kThreads);
  cfg.dynamicSmemBytes = wft::kSmemBytes + (g_debug_extra_smem > 0 ? g_debug_extra_smem : 0);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (launch_flags & (WFT_LAUNCH_PDL | WFT_LAUNCH_OVERLAP)) ? 1 : 0;
  WFT_CUDA(cudaLaunchKernelEx(&cfg, wft::frontend_kernel<NM, PcmT>, q));
  ++g_launches;
  if (!with_fixup) return WFT_OK;   // wft_frontend_augment_forward: the epilogue grid behind this one finishes the cells on load

  // the fix-up grid right behind it: always a programmatic dependent (it waits for the front-end grid on the device), one
  // lean CTA per SM -- 128 threads x 32 registers and no shared memory fit NEXT TO six front-end CTAs, so the blocks that
  // sit waiting for the front-end grid (or for the tail of it, while an overlapping launch already runs) cost no SM slot
  wft::FixupParams f{};
  f.out = p.out; f.stats = p.stats; f.tile_min = p.tile_min; f.lengths = p.lengths; f.n_valid = p.n_valid; f.masks = p.masks;
  f.draw = p.draw; f.draw_seed = p.draw_seed; f.draw_clip_offset = p.draw_clip_offset; f.draw_tparam = p.draw_tparam;
  f.draw_fparam = p.draw_fparam; f.draw_p = p.draw_p;
  f.n_samples = p.n_samples; f.n_total = p.n_total; f.n_frames = p.n_frames; f.n_frames_out = p.n_frames_out;
  f.tiles_per_clip = p.tiles_per_clip; f.total_tiles = p.total_tiles; f.mask_value = p.mask_value;
  const int groups = (p.total_tiles + wft::kFixTiles * wft::kFixWarps - 1) / (wft::kFixTiles * wft::kFixWarps);
  // ragged inputs (lengths / cuts / output longer than the clip) mean many constant-fill tiles: more CTAs per SM
  const bool heavy = p.lengths != nullptr || p.n_valid != nullptr || p.n_frames_out > p.n_frames;
  int fix_ctas = (ctas_full / 6) * (heavy ? 8 : 1);
  if (fix_ctas > groups) fix_ctas = groups;
  if (g_debug_max_ctas > 0 && fix_ctas > g_debug_max_ctas) fix_ctas = g_debug_max_ctas;
  cudaLaunchConfig_t fcfg{};
  fcfg.gridDim = dim3(fix_ctas < 1 ? 1 : fix_ctas);
  fcfg.blockDim = dim3(wft::kFixThreads);
  fcfg.dynamicSmemBytes = 0;
  fcfg.stream = stream;
  fcfg.attrs = attr;
  fcfg.numAttrs = 1;
  if (heavy) WFT_CUDA(cudaLaunchKernelEx(&fcfg, wft::fixup_kernel<NM, false>, f));
  else WFT_CUDA(cudaLaunchKernelEx(&fcfg, wft::fixup_kernel<NM, true>, f));
  ++g_launches;
  return WFT_OK;
}

template <int NM, typename PcmT>
int query_grid(int32_t* ctas) {
  static GridInfo cache[64];
  int c = 0;
  int rc = grid_for(wft::frontend_kernel<NM, PcmT>, cache, &c);
  *ctas = c;
  return rc;
}

// workspace = WFT_WS_PHASES headers {16 bytes (tile counter), ClipStat[batch]} (the part that has to be zero before a launch),
// then WFT_WS_PHASES bodies {tile_min[tiles]}
size_t ws_header_bytes(int32_t batch) { return (16 + sizeof(wft::ClipStat) * static_cast<size_t>(batch) + 15) & ~static_cast<size_t>(15); }
size_t ws_body_bytes(int32_t /*batch*/, int64_t tiles) {
  return (static_cast<size_t>(tiles) * sizeof(float) + 15) & ~static_cast<size_t>(15);
}

// ---- small stand-alone kernels ---------------------------------------------------------------------------

// scratch[0] = ~enc(min), scratch[1] = 1 if any element is NaN (torch.min propagates NaN; fminf would drop it)
__global__ void min_reduce_kernel(const float* __restrict__ in, int64_t n, uint32_t* __restrict__ min_inv) {
  float mn = INFINITY;
  bool nan = false;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = __ldg(in + k);
    nan |= v != v;
    mn = fminf(mn, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  nan = __any_sync(0xffffffffu, nan);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(min_inv, ~wft::enc_ordered(mn));   // always published: an all-(+inf) input pads with +inf
    if (nan) min_inv[1] = 1u;
  }
}

// out[o, l, i] = l < len_in ? in[o, l, i] : min(in)      (data/utils.py:380-404)
__global__ void pad_or_trim_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t outer,
                                   int64_t len_in, int64_t inner, int64_t len_out,
                                   const uint32_t* __restrict__ min_inv) {
  const int64_t n = outer * len_out * inner;
  const float fill = (len_out > len_in) ? (min_inv[1] != 0u ? __int_as_float(0x7fc00000) : wft::dec_ordered(~min_inv[0])) : 0.0f;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = k % inner;
    const int64_t l = (k / inner) % len_out;
    const int64_t o = k / (inner * len_out);
    out[k] = (l < len_in) ? __ldg(in + (o * len_in + l) * inner + i) : fill;
  }
}

// torchaudio mask_along_axis for time then frequency (data_loader.py:286-287), explicit intervals
__global__ void specaug_apply_kernel(const float* __restrict__ in, float* __restrict__ out, int32_t n_rows,
                                     int32_t n_frames, const int32_t* __restrict__ mask_params, float mask_value) {
  const int b = blockIdx.y;
  const int4 mk = __ldg(reinterpret_cast<const int4*>(mask_params) + b);
  const int64_t per_clip = static_cast<int64_t>(n_rows) * n_frames;
  const float* src = in + b * per_clip;
  float* dst = out + b * per_clip;
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < per_clip;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(k / n_frames);
    const int t = static_cast<int>(k - static_cast<int64_t>(r) * n_frames);
    const bool masked = (t >= mk.x && t < mk.y) || (r >= mk.z && r < mk.w);
    if (masked) dst[k] = mask_value;
    else if (src != dst) dst[k] = src[k];
  }
}

// Deep SpecAugment on encoder activations (model/model_utils.py:382-437): x is [batch, seq, dim], one (time, feature) mask
// for the whole batch.  Works on 16-byte vectors of 16- or 32-bit elements; masked rows are written without being read,
// a vector that straddles the feature mask's edge is patched element by element.  blockDim = (vectors, rows).
template <typename ElemT>
__device__ __forceinline__ uint4 patch_vector(uint4 v, int c0, int f0, int f1, ElemT fill) {
  constexpr int kPer = 16 / sizeof(ElemT);
  ElemT e[kPer];
  memcpy(e, &v, 16);
#pragma unroll
  for (int i = 0; i < kPer; ++i)
    if (c0 + i >= f0 && c0 + i < f1) e[i] = fill;
  memcpy(&v, e, 16);
  return v;
}

template <typename ElemT>
__global__ void mask_bsd_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n_rows, int32_t seq,
                                int32_t vec_per_row, int32_t t0, int32_t t1, int32_t f0, int32_t f1, uint32_t fill_bits) {
  constexpr int kPer = 16 / sizeof(ElemT);
  ElemT fill;
  memcpy(&fill, &fill_bits, sizeof(ElemT));
  const uint32_t f32 = sizeof(ElemT) == 2 ? (fill_bits & 0xffffu) * 0x10001u : fill_bits;
  const uint4 fill4 = make_uint4(f32, f32, f32, f32);
  const bool in_place = in == out;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.y + threadIdx.y; row < n_rows;
       row += static_cast<int64_t>(gridDim.x) * blockDim.y) {
    const int s = static_cast<int>(row % seq);
    const bool row_masked = s >= t0 && s < t1;
    const uint4* src = in + row * vec_per_row;
    uint4* dst = out + row * vec_per_row;
    for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) {
      const int c0 = v * kPer;
      const bool all_in = row_masked || (c0 >= f0 && c0 + kPer <= f1);
      const bool none_in = !row_masked && (c0 + kPer <= f0 || c0 >= f1);
      if (all_in) {
        dst[v] = fill4;
      } else if (none_in) {
        if (!in_place) dst[v] = __ldcs(src + v);
      } else {
        dst[v] = patch_vector<ElemT>(__ldcs(src + v), c0, f0, f1, fill);
      }
    }
  }
}

// any shape / alignment: one element per thread
template <typename ElemT>
__global__ void mask_bsd_scalar_kernel(const ElemT* __restrict__ in, ElemT* __restrict__ out, int64_t n_elems, int32_t seq,
                                       int32_t dim, int32_t t0, int32_t t1, int32_t f0, int32_t f1, uint32_t fill_bits) {
  ElemT fill;
  memcpy(&fill, &fill_bits, sizeof(ElemT));
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n_elems;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = k / dim;
    const int d = static_cast<int>(k - row * dim);
    const int s = static_cast<int>(row % seq);
    if ((s >= t0 && s < t1) || (d >= f0 && d < f1)) out[k] = fill;
    else if (in != out) out[k] = in[k];
  }
}

__global__ void specaug_draw_kernel(uint64_t seed, uint64_t clip_offset, int32_t batch, int32_t n_mels,
                                    int32_t n_frames, int32_t tparam, int32_t fparam, float p,
                                    int32_t* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  reinterpret_cast<int4*>(out)[b] = wft::draw_mask_intervals(seed, clip_offset + static_cast<uint64_t>(b), n_mels, n_frames, tparam, fparam, p);
}


// counter-based draw of (warp_p, warp_d): warp_p uniform in [W, T-W), warp_d uniform in [-W, W) (the reference's randint
// ranges, data/utils.py:107-111), Philox block 2 of the clip's counter; (-1, 0) == "no warp" when the p gate rejects
__device__ __noinline__ int2 draw_warp_point(uint64_t seed, uint64_t idx, int32_t n_frames, int32_t W, float p) {
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const uint32_t lo = static_cast<uint32_t>(idx), hi = static_cast<uint32_t>(idx >> 32);
  bool apply = p >= 1.0f;
  if (!apply && p > 0.0f) {
    uint32_t g[4];
    wft::philox4x32_10(lo, hi, 1u, 0u, k0, k1, g);
    apply = wft::u01(g[0]) < p;
  }
  int2 w = make_int2(-1, 0);   // "no warp": the warp kernels copy such a clip
  if (apply && W > 0 && n_frames > 2 * W) {
    uint32_t r[4];
    wft::philox4x32_10(lo, hi, 2u, 0u, k0, k1, r);
    w.x = W + static_cast<int>(__fmul_rn(wft::u01(r[0]), static_cast<float>(n_frames - 2 * W)));
    w.y = -W + static_cast<int>(__fmul_rn(wft::u01(r[1]), static_cast<float>(2 * W)));
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
// be finished once the whole clip is known -- and every tap is finished on load exactly like wft::fixup_tile would have
// rewritten it: max(v, floor) for the kept frames, the clamp value for tiles that were never computed (silent / pad-only),
// the min-value pad beyond the kept frames (data/utils.py:380-404).
struct AugFix {
  const wft::ClipStat* stats;    // this call's clip statistics (complete once the front-end grid is)
  const int32_t* lengths;
  const int32_t* n_valid;
  int32_t n_samples, n_total, n_frames;   // of the front-end call (frames the clip really has; T is n_frames_out)
};

// grid = (frame blocks of 1024, row groups of 16, clips); a thread owns 4 output frames, 256 apart (lane <-> consecutive frames:
// a warp's store is one 128-byte line and the two source taps of a smooth, monotone map fall into one or two lines -- with 4
// ADJACENT frames per thread every scalar load of a warp was spread over 4-8 lines and the kernel sat at 0.40 of the HBM peak
// on the L1 data pipe); it evaluates the 4 source coordinates once and walks the 16 rows of its group with 8 independent
// loads in flight per row (bilinear taps mirror grid_sample's float32 arithmetic, zeros outside).
// the value fixup_tile would have left in a cell the front-end kernel wrote as v (kind: 0 = computed, 1 = never computed, 2 = pad)
// (s_fix: floor feature, pad value, kept frames, clip length, clamp feature; the last two values a ragged clip needs are read
// from shared memory where they are used -- a register each would spill the float64-spline instance)
__device__ __forceinline__ float aug_finish(float v, uint32_t kind, bool ragged, float floorn, const int* __restrict__ s_fix) {
  if (!ragged) return fmaxf(v, floorn);
  return kind == 2u ? __int_as_float(s_fix[1]) : fmaxf(kind == 1u ? __int_as_float(s_fix[4]) : v, floorn);
}

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
  // source row(s) of every output row of the group: grid_sample's y coordinate of row r (the identity up to float32 rounding,
  // which can put a sliver of weight on the next row -- restated, not assumed), once per CTA instead of once per thread and row
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kAugRowsPerCta) {
    const int r = blockIdx.y * kAugRowsPerCta + (threadIdx.x - 64);
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
      const int4 m = wft::draw_mask_intervals(draw.seed, draw.clip_offset + static_cast<uint64_t>(b), R, T, draw.tparam, draw.fparam, draw.p);
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
  if (kFix != 0 && threadIdx.x == 96) {   // the front-end grid is complete (griddepcontrol.wait above): its statistics are final
    const float floorn = wft::floor_feature(wft::dec_ordered(__ldcg(&fix.stats[b].max_enc)));
    const float padv = fmaxf(wft::feature_of_l2(wft::dec_ordered(~__ldcg(&fix.stats[b].min_inv))), floorn);
    int len = fix.n_samples;
    if (fix.lengths != nullptr) {
      const int l = __ldg(fix.lengths + b);
      len = l < 0 ? 0 : (l < len ? l : len);
    }
    s_fix[0] = __float_as_int(floorn);
    s_fix[1] = __float_as_int(padv);
    s_fix[2] = wft::kept_frames(fix.n_valid, b, fix.n_frames);
    s_fix[3] = len;
    s_fix[4] = __float_as_int(wft::feature_of_l2(wft::silent_l2()));
  }
  __syncthreads();
  const int tbase = blockIdx.x * kAugThreads * kAugFramesPerThread + threadIdx.x;
  constexpr int kStep = kAugThreads;   // frame k of this thread = tbase + k * kStep
  if (tbase >= T) return;
  const int wp = s_draw[4], wd = s_draw[5];
  const int4 mk = make_int4(s_draw[0], s_draw[1], s_draw[2], s_draw[3]);
  int lo_rows = 0, hi_rows = 0;
  if (extremes != nullptr) {
    const int2 e = __ldg(reinterpret_cast<const int2*>(extremes) + b);
    lo_rows = e.x; hi_rows = e.y;
  }
  const bool warp = wp > 0 && wp < T - 1;          // anything else (the draw's "gate rejected" marker is -1) = no warp
  // per frame, once: the two source columns (clamped into the row so that every load is unconditional) and their weights
  // (0 for a tap that falls outside: "zeros" padding -- 0 * finite contributes exactly nothing), and whether the cell is live
  int oa[kAugFramesPerThread];
  uint32_t cstep = 0;   // bit k: the second tap of frame k sits one column right of the first (0 where both clamp to one column)
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
    const int c = a + 1;
    wa[k] = (a >= 0 && a < T) ? wx0 : 0.0f;
    wc[k] = (c >= 0 && c < T) ? wx1 : 0.0f;
    oa[k] = min(max(a, 0), T - 1);
    cstep |= static_cast<uint32_t>(min(max(c, 0), T - 1) - oa[k]) << k;
  }
  // kFix: what a tap at source column oa[k] / oc[k] still needs (per frame, once): 0 = floor only, 1 = never computed (the
  // clamp value, then the floor), 2 = beyond the kept frames (the pad value).  Full-length clips without a cut need no table.
  float floorn = 0.0f;
  bool ragged = false;
  uint32_t kinds = 0;   // 2 bits per tap: tap a of frame k at bit 4k, tap c at bit 4k + 2
  if constexpr (kFix != 0) floorn = __int_as_float(s_fix[0]);
  if constexpr (kFix == 2) {
    const int keep = s_fix[2], len = s_fix[3];
    ragged = keep < T || len < fix.n_samples || fix.n_frames < T;
    if (ragged) {
#pragma unroll
      for (int k = 0; k < kAugFramesPerThread; ++k) {
        const int a = oa[k], c = oa[k] + static_cast<int>((cstep >> k) & 1u);
        const uint32_t ka = a >= keep ? 2u : (wft::tile_is_silent(a & ~(wft::kTileFrames - 1), len, fix.n_total) ? 1u : 0u);
        const uint32_t kc = c >= keep ? 2u : (wft::tile_is_silent(c & ~(wft::kTileFrames - 1), len, fix.n_total) ? 1u : 0u);
        kinds |= (ka | (kc << 2)) << (4 * k);
      }
    }
  }
#define AUG_FINISH(v, k, tap) aug_finish(v, (kinds >> (4 * (k) + 2 * (tap))) & 3u, ragged, floorn, s_fix)
  const size_t clip = static_cast<size_t>(b) * R * T;
  const int r_end = min(R, static_cast<int>(blockIdx.y + 1) * kAugRowsPerCta);
  for (int r = blockIdx.y * kAugRowsPerCta; r < r_end; ++r) {
    float v[kAugFramesPerThread];
    const bool rowmask = (r >= mk.z && r < mk.w) || r < lo_rows || r >= R - hi_rows;
    if (rowmask) {
#pragma unroll
      for (int k = 0; k < kAugFramesPerThread; ++k) v[k] = mask_value;
    } else if (!warp) {
      const float* row = in + clip + static_cast<size_t>(r) * T;
#pragma unroll
      for (int k = 0; k < kAugFramesPerThread; ++k) {
        if constexpr (kFix != 0) v[k] = on[k] ? AUG_FINISH(__ldg(row + oa[k]), k, 0) : mask_value;
        else v[k] = on[k] ? __ldg(row + oa[k]) : mask_value;
      }
    } else {
      const int4 rr = s_row[r - blockIdx.y * kAugRowsPerCta];
      const float wy0 = __int_as_float(rr.y), wy1 = __int_as_float(rr.z);
      const bool use1 = rr.w >= 0;
      const int iy1 = rr.w;
      const float* row0 = in + clip + static_cast<size_t>(rr.x) * T;
      // every tap of the row is requested before the first one is used (8 independent loads in flight per thread)
      float t0a[kAugFramesPerThread], t0c[kAugFramesPerThread];
#pragma unroll
      for (int k = 0; k < kAugFramesPerThread; ++k) {
        t0a[k] = __ldg(row0 + oa[k]);
        t0c[k] = __ldg(row0 + oa[k] + ((cstep >> k) & 1u));
      }
      if constexpr (kFix != 0) {
#pragma unroll
        for (int k = 0; k < kAugFramesPerThread; ++k) {
          t0a[k] = AUG_FINISH(t0a[k], k, 0);
          t0c[k] = AUG_FINISH(t0c[k], k, 1);
        }
      }
      if (!use1) {      // warp-uniform (depends on r alone); taps accumulate in grid_sample's order
#pragma unroll
        for (int k = 0; k < kAugFramesPerThread; ++k) {
          float acc = 0.0f;
          acc += t0a[k] * (wa[k] * wy0);
          acc += t0c[k] * (wc[k] * wy0);
          v[k] = on[k] ? acc : mask_value;
        }
      } else {
        const float* row1 = in + clip + static_cast<size_t>(iy1) * T;
        float t1a[kAugFramesPerThread], t1c[kAugFramesPerThread];
#pragma unroll
        for (int k = 0; k < kAugFramesPerThread; ++k) {
          t1a[k] = __ldg(row1 + oa[k]);
          t1c[k] = __ldg(row1 + oa[k] + ((cstep >> k) & 1u));
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
    float* dst = out + clip + static_cast<size_t>(r) * T + tbase;
#pragma unroll
    for (int k = 0; k < kAugFramesPerThread; ++k)
      if (tbase + k * kStep < T) dst[k * kStep] = v[k];
  }
}

#undef AUG_FINISH

int grid_1d(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int wft_abi_version(void) { return WFT_ABI_VERSION; }

const char* wft_last_error(void) { return g_last_error.c_str(); }

#ifdef WFT_TIMELINE
// development build only: copy the kernel's clock stamps to the host (kTlMaxCtas x kTlIters x kWarps x kTlPoints words)
int wft_debug_timeline(uint32_t* host_out, int32_t* dims) {
  dims[0] = wft::kTlMaxCtas; dims[1] = wft::kTlIters; dims[2] = wft::kWarps; dims[3] = wft::kTlPoints;
  if (host_out == nullptr) return WFT_OK;
  WFT_CUDA(cudaMemcpyFromSymbol(host_out, wft::g_timeline, sizeof(wft::g_timeline)));
  return WFT_OK;
}
#endif

int wft_debug_set_max_ctas(int32_t max_ctas) {
  g_debug_max_ctas = max_ctas > 0 ? max_ctas : 0;
  return WFT_OK;
}

int wft_debug_set_extra_smem(int32_t bytes) {
  g_debug_extra_smem = bytes > 0 ? bytes : 0;
  return WFT_OK;
}

int64_t wft_launch_count(int reset) {
  const int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int wft_frontend_workspace_bytes(int32_t batch, int32_t n_samples_total, int32_t n_frames_out, size_t* bytes) {
  if (bytes == nullptr) return fail(WFT_ERR_INVALID, "bytes is NULL");
  if (batch < 1) return fail(WFT_ERR_INVALID, "batch must be >= 1");
  if (n_samples_total <= WFT_N_FFT / 2) return fail(WFT_ERR_INVALID, "clips must be longer than 200 samples (reflect pad)");
  if (n_frames_out < 0) return fail(WFT_ERR_INVALID, "n_frames_out must be >= 0");
  const int64_t n_frames = n_samples_total / WFT_HOP_LENGTH;
  const int64_t span = n_frames_out > n_frames ? n_frames_out : n_frames;
  const int64_t tiles = (span + wft::kTileFrames - 1) / wft::kTileFrames * batch;
  if (tiles < 1 || tiles >= (int64_t(1) << 30)) return fail(WFT_ERR_INVALID, "batch x frames out of range (tile ids must stay below 2^30)");
  size_t b = WFT_WS_PHASES * (ws_header_bytes(batch) + ws_body_bytes(batch, tiles));
  *bytes = (b + 255) & ~static_cast<size_t>(255);
  return WFT_OK;
}

// validate, lay out the workspace, launch the front-end grid (and, with_fixup, its fix-up grid); *used = what was launched
static int frontend_forward_impl(const wft_frontend_args* a, cudaStream_t stream, bool with_fixup, wft::FrontendParams* used) {
  if (a == nullptr) return fail(WFT_ERR_INVALID, "args is NULL");
  if (a->pcm == nullptr || a->out == nullptr || a->workspace == nullptr)
    return fail(WFT_ERR_INVALID, "pcm, out and workspace must be non-NULL device pointers");
  if (a->n_mels != 80 && a->n_mels != 128) return fail(WFT_ERR_INVALID, "Unsupported n_mels: " + std::to_string(a->n_mels));
  if (a->pcm_dtype != WFT_PCM_F32 && a->pcm_dtype != WFT_PCM_I16) return fail(WFT_ERR_INVALID, "pcm_dtype must be WFT_PCM_F32 or WFT_PCM_I16");
  if (a->batch < 1) return fail(WFT_ERR_INVALID, "batch must be >= 1");
  if (a->n_samples < 1 || a->padding < 0) return fail(WFT_ERR_INVALID, "n_samples must be >= 1 and padding >= 0");
  if (a->clip_stride < a->n_samples) return fail(WFT_ERR_INVALID, "clip_stride must be >= n_samples");
  const int64_t n_total64 = static_cast<int64_t>(a->n_samples) + a->padding;
  if (n_total64 > INT32_MAX / 4) return fail(WFT_ERR_INVALID, "clip too long");
  const int32_t n_total = static_cast<int32_t>(n_total64);
  if (n_total <= WFT_N_FFT / 2) return fail(WFT_ERR_INVALID, "clips must be longer than 200 samples (reflect pad)");
  const int32_t n_frames = n_total / WFT_HOP_LENGTH;
  if (n_frames < 1) return fail(WFT_ERR_INVALID, "clip shorter than one hop");
  const int32_t n_frames_out = a->n_frames_out > 0 ? a->n_frames_out : n_frames;
  size_t need = 0;
  int rc = wft_frontend_workspace_bytes(a->batch, n_total, n_frames_out, &need);
  if (rc != WFT_OK) return rc;
  if (static_cast<int64_t>(a->n_mels) * n_frames_out * 4 >= (int64_t(1) << 32))
    return fail(WFT_ERR_INVALID, "n_frames_out too large");
  if (a->workspace_bytes < need) return fail(WFT_ERR_INVALID, "workspace too small: need " + std::to_string(need) + " bytes");
  if ((reinterpret_cast<uintptr_t>(a->workspace) & 15) != 0) return fail(WFT_ERR_INVALID, "workspace must be 16-byte aligned");
  if (a->mask_params != nullptr && (reinterpret_cast<uintptr_t>(a->mask_params) & 15) != 0)
    return fail(WFT_ERR_INVALID, "mask_params must be 16-byte aligned");
  if (a->draw_masks != 0) {
    if (a->mask_params != nullptr) return fail(WFT_ERR_INVALID, "draw_masks and mask_params are mutually exclusive");
    if (!(a->draw_p >= 0.0f && a->draw_p <= 1.0f)) return fail(WFT_ERR_INVALID, "spec_augment p must be between 0 and 1");
  }
  if (((n_frames_out & 3) == 0) && (reinterpret_cast<uintptr_t>(a->out) & 15) != 0)
    return fail(WFT_ERR_INVALID, "out must be 16-byte aligned");

  wft::FrontendParams p{};
  p.pcm = a->pcm;
  p.clip_stride = a->clip_stride;
  p.lengths = a->lengths;
  p.n_valid = a->n_valid_frames;
  p.masks = a->mask_params;
  p.out = a->out;
  const int mode = a->workspace_mode;
  const bool ring = mode >= WFT_WS_RING && mode < WFT_WS_RING + WFT_WS_PHASES;
  if (!ring && (mode < WFT_WS_MEMSET || mode > WFT_WS_PHASE_B))
    return fail(WFT_ERR_INVALID, "workspace_mode must be WFT_WS_MEMSET, WFT_WS_PHASE_A, WFT_WS_PHASE_B or WFT_WS_RING + k");
  if ((a->launch_flags & WFT_LAUNCH_OVERLAP) != 0 && !ring)
    return fail(WFT_ERR_INVALID, "WFT_LAUNCH_OVERLAP needs workspace_mode WFT_WS_RING + k (launches in flight must not share counters)");
  const size_t hdr = ws_header_bytes(a->batch);
  const int phase = ring ? mode - WFT_WS_RING : (mode == WFT_WS_PHASE_B ? 1 : 0);
  uint8_t* wsbase = static_cast<uint8_t*>(a->workspace);
  uint8_t* ws = wsbase + phase * hdr;
  p.tile_counter = reinterpret_cast<uint32_t*>(ws);
  p.stats = reinterpret_cast<wft::ClipStat*>(ws + 16);
  const int32_t span0 = n_frames_out > n_frames ? n_frames_out : n_frames;
  const int64_t tiles0 = static_cast<int64_t>((span0 + wft::kTileFrames - 1) / wft::kTileFrames) * a->batch;
  uint8_t* body = wsbase + WFT_WS_PHASES * hdr + phase * ws_body_bytes(a->batch, tiles0);
  p.tile_min = reinterpret_cast<float*>(body);
  p.clean = nullptr;
  p.clean_vec = 0;
  if (mode == WFT_WS_PHASE_A || mode == WFT_WS_PHASE_B) {   // self-cleaning: this launch zeroes the other phase for the next one
    p.clean = reinterpret_cast<uint4*>(wsbase + (1 - phase) * hdr);
    p.clean_vec = static_cast<int32_t>(hdr / 16);
  }
  p.n_samples = a->n_samples;
  p.n_total = n_total;
  p.batch = a->batch;
  p.n_frames = n_frames;
  p.n_frames_out = n_frames_out;
  const int32_t span = n_frames_out > n_frames ? n_frames_out : n_frames;
  p.tiles_per_clip = (span + wft::kTileFrames - 1) / wft::kTileFrames;
  p.total_tiles = p.tiles_per_clip * a->batch;
  p.mask_value = a->mask_value;
  const uint64_t magic = (uint64_t(1) << 32) / static_cast<uint64_t>(p.tiles_per_clip);
  p.tpc_magic = magic > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(magic);
  p.vec_ok = ((reinterpret_cast<uintptr_t>(a->out) & 31) == 0 && (n_frames_out & 7) == 0) ? 1 : 0;

  // MEMSET: this call's counters; RING: all phases at once when the ring wraps (a plain stream operation, i.e. the one
  // point per WFT_WS_PHASES calls where everything in front has to be complete)
  if (mode == WFT_WS_MEMSET) WFT_CUDA(cudaMemsetAsync(ws, 0, hdr, stream));
  if (ring && phase == 0) WFT_CUDA(cudaMemsetAsync(wsbase, 0, WFT_WS_PHASES * hdr, stream));
  int launch_flags = a->launch_flags;
  if (a->draw_masks != 0) {   // the intervals are a pure function of (seed, clip index): every CTA draws what it needs itself
    p.draw = 1;
    p.draw_seed = a->draw_seed;
    p.draw_clip_offset = a->draw_clip_offset;
    p.draw_tparam = a->draw_time_mask_param;
    p.draw_fparam = a->draw_freq_mask_param;
    p.draw_p = a->draw_p;
  }
  if (used != nullptr) *used = p;
  if (a->n_mels == 128) {
    return a->pcm_dtype == WFT_PCM_F32 ? launch_frontend<128, float>(p, stream, launch_flags, with_fixup)
                                       : launch_frontend<128, int16_t>(p, stream, launch_flags, with_fixup);
  }
  return a->pcm_dtype == WFT_PCM_F32 ? launch_frontend<80, float>(p, stream, launch_flags, with_fixup)
                                     : launch_frontend<80, int16_t>(p, stream, launch_flags, with_fixup);
}

int wft_frontend_forward(const wft_frontend_args* a, void* stream_) {
  return frontend_forward_impl(a, static_cast<cudaStream_t>(stream_), true, nullptr);
}

int wft_frontend_grid(int32_t n_mels, int32_t pcm_dtype, int32_t* ctas, int32_t* threads, int32_t* smem_bytes) {
  if (ctas == nullptr || threads == nullptr || smem_bytes == nullptr) return fail(WFT_ERR_INVALID, "NULL output pointer");
  if (n_mels != 80 && n_mels != 128) return fail(WFT_ERR_INVALID, "Unsupported n_mels: " + std::to_string(n_mels));
  *threads = wft::kThreads;
  *smem_bytes = wft::kSmemBytes;
  if (n_mels == 128) return pcm_dtype == WFT_PCM_I16 ? query_grid<128, int16_t>(ctas) : query_grid<128, float>(ctas);
  return pcm_dtype == WFT_PCM_I16 ? query_grid<80, int16_t>(ctas) : query_grid<80, float>(ctas);
}

int wft_pad_or_trim_f32(const float* in, int64_t outer, int64_t len_in, int64_t inner, int64_t len_out, float* out,
                        void* scratch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (outer < 0 || len_in < 0 || inner < 0 || len_out < 0) return fail(WFT_ERR_INVALID, "negative extent");
  const int64_t n_out = outer * len_out * inner;
  if (n_out == 0) return WFT_OK;
  if (out == nullptr) return fail(WFT_ERR_INVALID, "out is NULL");
  const int64_t n_in = outer * len_in * inner;
  if (len_out > len_in) {
    if (n_in == 0) return fail(WFT_ERR_INVALID, "pad_or_trim: cannot take the minimum of an empty array");
    if (scratch == nullptr) return fail(WFT_ERR_INVALID, "scratch is NULL");
    if (in == nullptr) return fail(WFT_ERR_INVALID, "in is NULL");
    WFT_CUDA(cudaMemsetAsync(scratch, 0, 16, stream));
    min_reduce_kernel<<<grid_1d(n_in, 256), 256, 0, stream>>>(in, n_in, static_cast<uint32_t*>(scratch));
    ++g_launches;
    WFT_CUDA(cudaGetLastError());
  } else if (in == nullptr) {
    return fail(WFT_ERR_INVALID, "in is NULL");
  }
  pad_or_trim_kernel<<<grid_1d(n_out, 256), 256, 0, stream>>>(in, out, outer, len_in, inner, len_out,
                                                               static_cast<const uint32_t*>(scratch));
  ++g_launches;
  WFT_CUDA(cudaGetLastError());
  return WFT_OK;
}

int wft_specaug_apply_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames,
                          const int32_t* mask_params, float mask_value, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (batch < 0 || n_rows < 0 || n_frames < 0) return fail(WFT_ERR_INVALID, "negative extent");
  if (static_cast<int64_t>(batch) * n_rows * n_frames == 0) return WFT_OK;
  if (in == nullptr || out == nullptr || mask_params == nullptr) return fail(WFT_ERR_INVALID, "NULL pointer");
  if ((reinterpret_cast<uintptr_t>(mask_params) & 15) != 0) return fail(WFT_ERR_INVALID, "mask_params must be 16-byte aligned");
  if (batch > 65535) return fail(WFT_ERR_INVALID, "batch too large for one launch (max 65535)");
  dim3 grid(grid_1d(static_cast<int64_t>(n_rows) * n_frames, 256), batch);
  specaug_apply_kernel<<<grid, 256, 0, stream>>>(in, out, n_rows, n_frames, mask_params, mask_value);
  ++g_launches;
  WFT_CUDA(cudaGetLastError());
  return WFT_OK;
}

int wft_mask_bsd(const void* in, void* out, int32_t elem_bytes, int64_t batch, int32_t seq, int32_t dim, int32_t t0,
                 int32_t t1, int32_t f0, int32_t f1, uint32_t fill_bits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (elem_bytes != 2 && elem_bytes != 4) return fail(WFT_ERR_INVALID, "elem_bytes must be 2 (fp16 / bf16) or 4 (fp32)");
  if (batch < 0 || seq < 0 || dim < 0) return fail(WFT_ERR_INVALID, "negative extent");
  const int64_t n_rows = batch * seq;
  if (n_rows == 0 || dim == 0) return WFT_OK;
  if (in == nullptr || out == nullptr) return fail(WFT_ERR_INVALID, "NULL pointer");
  const int64_t row_bytes = static_cast<int64_t>(dim) * elem_bytes;
  const bool vec_ok = row_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  int dev = 0, sms = 148;
  WFT_CUDA(cudaGetDevice(&dev));
  WFT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (vec_ok) {
    const int vec_per_row = static_cast<int>(row_bytes / 16);
    int bx = (vec_per_row + 31) / 32 * 32;
    if (bx > 256) bx = 256;
    const int by = 256 / bx > 0 ? 256 / bx : 1;
    int64_t ctas = (n_rows + by - 1) / by;
    if (ctas > static_cast<int64_t>(sms) * 8) ctas = static_cast<int64_t>(sms) * 8;
    const dim3 block(bx, by), grid(static_cast<unsigned>(ctas));
    if (elem_bytes == 2)
      mask_bsd_kernel<uint16_t><<<grid, block, 0, stream>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), n_rows, seq,
                                                           vec_per_row, t0, t1, f0, f1, fill_bits);
    else
      mask_bsd_kernel<uint32_t><<<grid, block, 0, stream>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), n_rows, seq,
                                                           vec_per_row, t0, t1, f0, f1, fill_bits);
  } else {
    const int64_t n = n_rows * dim;
    const int grid = grid_1d(n, 256);
    if (elem_bytes == 2)
      mask_bsd_scalar_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), n,
                                                                seq, dim, t0, t1, f0, f1, fill_bits);
    else
      mask_bsd_scalar_kernel<uint32_t><<<grid, 256, 0, stream>>>(static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), n,
                                                                seq, dim, t0, t1, f0, f1, fill_bits);
  }
  ++g_launches;
  WFT_CUDA(cudaGetLastError());
  return WFT_OK;
}

int wft_specaug_draw(uint64_t seed, uint64_t clip_offset, int32_t batch, int32_t n_mels, int32_t n_frames,
                     int32_t time_mask_param, int32_t freq_mask_param, float p, int32_t* mask_params_out,
                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (batch < 1) return fail(WFT_ERR_INVALID, "batch must be >= 1");
  if (n_mels < 1 || n_frames < 1) return fail(WFT_ERR_INVALID, "n_mels and n_frames must be >= 1");
  if (!(p >= 0.0f && p <= 1.0f)) return fail(WFT_ERR_INVALID, "spec_augment p must be between 0 and 1");
  if (mask_params_out == nullptr) return fail(WFT_ERR_INVALID, "mask_params_out is NULL");
  if ((reinterpret_cast<uintptr_t>(mask_params_out) & 15) != 0) return fail(WFT_ERR_INVALID, "mask_params_out must be 16-byte aligned");
  specaug_draw_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(seed, clip_offset, batch, n_mels, n_frames,
                                                               time_mask_param, freq_mask_param, p, mask_params_out);
  ++g_launches;
  WFT_CUDA(cudaGetLastError());
  return WFT_OK;
}

static int launch_augment(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames, const int32_t* warp_params,
                          const int32_t* mask_params, const int32_t* extremes, float mask_value, int32_t spline_f32,
                          const AugDraw& draw, cudaStream_t stream, const AugFix* fix = nullptr) {
  if (batch < 0 || n_rows < 0 || n_frames < 0) return fail(WFT_ERR_INVALID, "negative extent");
  if (static_cast<int64_t>(batch) * n_rows * n_frames == 0) return WFT_OK;
  if (in == nullptr || out == nullptr) return fail(WFT_ERR_INVALID, "NULL pointer");
  if (warp_params != nullptr || (draw.enabled && draw.W > 0)) {
    if (in == out) return fail(WFT_ERR_INVALID, "time warp cannot run in place");
    if (n_rows < 2 || n_frames < 3) return fail(WFT_ERR_INVALID, "time warp needs at least 2 rows and 3 frames");
    if ((reinterpret_cast<uintptr_t>(warp_params) & 7) != 0) return fail(WFT_ERR_INVALID, "warp_params must be 8-byte aligned");
  }
  if (mask_params != nullptr && (reinterpret_cast<uintptr_t>(mask_params) & 15) != 0)
    return fail(WFT_ERR_INVALID, "mask_params must be 16-byte aligned");
  if (extremes != nullptr && (reinterpret_cast<uintptr_t>(extremes) & 7) != 0)
    return fail(WFT_ERR_INVALID, "extremes must be 8-byte aligned");
  if ((n_frames & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) != 0)
    return fail(WFT_ERR_INVALID, "in and out must be 16-byte aligned");
  if (batch > 65535) return fail(WFT_ERR_INVALID, "batch too large for one launch (max 65535)");
  const int per_cta = kAugThreads * kAugFramesPerThread;
  // always a programmatic dependent: the kernel waits for its predecessor on the device (griddepcontrol.wait), so its
  // launch latency and -- for the drawn variant -- its draws hide under the tail of the kernel that produces `in`
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((n_frames + per_cta - 1) / per_cta, (n_rows + kAugRowsPerCta - 1) / kAugRowsPerCta, batch);
  cfg.blockDim = dim3(kAugThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const AugFix fx = fix != nullptr ? *fix : AugFix{};
  // (same rule as the fix-up grid's heavy instance: lengths / cuts / an output longer than the clip)
  const bool ragged = fix != nullptr && (fix->lengths != nullptr || fix->n_valid != nullptr || n_frames > fix->n_frames);
#define WFT_AUG_LAUNCH(F32, FIX) \
  WFT_CUDA(cudaLaunchKernelEx(&cfg, augment_kernel<F32, FIX>, in, out, n_rows, n_frames, warp_params, mask_params, extremes, mask_value, draw, fx))
  if (fix == nullptr) {
    if (spline_f32) WFT_AUG_LAUNCH(true, 0); else WFT_AUG_LAUNCH(false, 0);
  } else if (!ragged) {
    if (spline_f32) WFT_AUG_LAUNCH(true, 1); else WFT_AUG_LAUNCH(false, 1);
  } else {
    if (spline_f32) WFT_AUG_LAUNCH(true, 2); else WFT_AUG_LAUNCH(false, 2);
  }
#undef WFT_AUG_LAUNCH
  ++g_launches;
  return WFT_OK;
}

int wft_augment_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames, const int32_t* warp_params,
                    const int32_t* mask_params, const int32_t* extremes, float mask_value, int32_t spline_f32, void* stream_) {
  return launch_augment(in, out, batch, n_rows, n_frames, warp_params, mask_params, extremes, mask_value, spline_f32, AugDraw{},
                        static_cast<cudaStream_t>(stream_));
}

int wft_augment_drawn_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames, uint64_t seed,
                          uint64_t clip_offset, int32_t time_mask_param, int32_t freq_mask_param, int32_t time_warp_w, float p,
                          const int32_t* extremes, float mask_value, int32_t spline_f32, void* stream_) {
  if (!(p >= 0.0f && p <= 1.0f)) return fail(WFT_ERR_INVALID, "spec_augment p must be between 0 and 1");
  if (time_warp_w < 0) return fail(WFT_ERR_INVALID, "time_warp_w must be >= 0");
  AugDraw d{};
  d.enabled = 1; d.tparam = time_mask_param; d.fparam = freq_mask_param; d.W = time_warp_w; d.p = p;
  d.seed = seed; d.clip_offset = clip_offset;
  return launch_augment(in, out, batch, n_rows, n_frames, nullptr, nullptr, extremes, mask_value, spline_f32, d,
                        static_cast<cudaStream_t>(stream_));
}

int wft_frontend_augment_forward(const wft_frontend_args* a, const wft_augment_args* g, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr || g == nullptr) return fail(WFT_ERR_INVALID, "args is NULL");
  if (g->out == nullptr) return fail(WFT_ERR_INVALID, "augment out is NULL");
  if (a->mask_params != nullptr || a->draw_masks != 0)
    return fail(WFT_ERR_INVALID, "the masks of a front-end + augmentation call belong to the augmentation arguments");
  if (g->out == a->out) return fail(WFT_ERR_INVALID, "the un-augmented features (scratch) and the output must be different buffers");
  AugDraw d{};
  if (g->draw != 0) {
    if (g->warp_params != nullptr || g->mask_params != nullptr) return fail(WFT_ERR_INVALID, "draw and explicit parameters are mutually exclusive");
    if (!(g->draw_p >= 0.0f && g->draw_p <= 1.0f)) return fail(WFT_ERR_INVALID, "spec_augment p must be between 0 and 1");
    if (g->draw_time_warp_w < 0) return fail(WFT_ERR_INVALID, "time_warp_w must be >= 0");
    d.enabled = 1; d.tparam = g->draw_time_mask_param; d.fparam = g->draw_freq_mask_param; d.W = g->draw_time_warp_w; d.p = g->draw_p;
    d.seed = g->draw_seed; d.clip_offset = g->draw_clip_offset;
  }
  wft::FrontendParams used{};
  int rc = frontend_forward_impl(a, stream, false, &used);
  if (rc != WFT_OK) return rc;
  AugFix fix{};
  fix.stats = used.stats; fix.lengths = used.lengths; fix.n_valid = used.n_valid;
  fix.n_samples = used.n_samples; fix.n_total = used.n_total; fix.n_frames = used.n_frames;
  return launch_augment(a->out, g->out, a->batch, a->n_mels, used.n_frames_out, g->warp_params, g->mask_params, g->extremes,
                        g->mask_value, g->spline_f32, d, stream, &fix);
}

int wft_time_warp_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames,
                      const int32_t* warp_params, void* stream_) {
  if (static_cast<int64_t>(batch) * n_rows * n_frames != 0 && warp_params == nullptr) return fail(WFT_ERR_INVALID, "NULL pointer");
  return wft_augment_f32(in, out, batch, n_rows, n_frames, warp_params, nullptr, nullptr, 0.0f, 0, stream_);
}

int wft_time_warp_draw(uint64_t seed, uint64_t clip_offset, int32_t batch, int32_t n_frames, int32_t time_warp_w, float p,
                       int32_t* warp_params_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (batch < 1) return fail(WFT_ERR_INVALID, "batch must be >= 1");
  if (n_frames < 3 || time_warp_w < 0) return fail(WFT_ERR_INVALID, "n_frames must be >= 3 and time_warp_w >= 0");
  if (!(p >= 0.0f && p <= 1.0f)) return fail(WFT_ERR_INVALID, "spec_augment p must be between 0 and 1");
  if (warp_params_out == nullptr) return fail(WFT_ERR_INVALID, "warp_params_out is NULL");
  if ((reinterpret_cast<uintptr_t>(warp_params_out) & 7) != 0) return fail(WFT_ERR_INVALID, "warp_params_out must be 8-byte aligned");
  time_warp_draw_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(seed, clip_offset, batch, n_frames, time_warp_w, p,
                                                                 warp_params_out);
  ++g_launches;
  WFT_CUDA(cudaGetLastError());
  return WFT_OK;
}

}  // extern "C"
