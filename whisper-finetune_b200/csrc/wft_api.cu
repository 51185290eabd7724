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
#include "augment_kernel.cuh"

namespace {

thread_local std::string g_last_error;
thread_local int64_t g_launches = 0;
int g_debug_chunk = 0;      // WFT_DEBUG_CHUNK (development): force the tiles-per-claim of the fused kernel
int g_debug_max_ctas = 0;   // wft_debug_set_max_ctas: caps the persistent grids (results must not depend on the grid)
int g_debug_aug_generic = 0;   // wft_debug_set_augment_generic: the epilogue's generic instance even where the staged one applies
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
  // ragged inputs (lengths / cuts / output longer than the clip) mean many constant-fill tiles: more CTAs per SM, smaller groups
  const bool heavy = p.lengths != nullptr || p.n_valid != nullptr || p.n_frames_out > p.n_frames;
  const int fix_per_cta = (heavy ? wft::kFixTilesHeavy : wft::kFixTiles) * wft::kFixWarps;
  const int groups = (p.total_tiles + fix_per_cta - 1) / fix_per_cta;
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
// grid.y walks `outer`, grid.x x threads walk the len_out * inner cells of one outer slice: the cell index splits into (l, i)
// with 32-bit arithmetic (or none at all when inner == 1, the axis = -1 case of data_loader.py:282) instead of two 64-bit
// divisions per element.
__global__ void pad_or_trim_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t outer,
                                   int64_t len_in, int64_t inner, int64_t len_out,
                                   const uint32_t* __restrict__ min_inv) {
  const float fill = (len_out > len_in) ? (min_inv[1] != 0u ? __int_as_float(0x7fc00000) : wft::dec_ordered(~min_inv[0])) : 0.0f;
  const int64_t per_out = len_out * inner, per_in = len_in * inner;
  const int64_t live = (len_in < len_out ? len_in : len_out) * inner;   // cells [0, live) of a slice are copies, the rest the fill
  for (int64_t o = blockIdx.y; o < outer; o += gridDim.y) {
    const float* src = in + o * per_in;
    float* dst = out + o * per_out;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < per_out;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x)
      dst[j] = j < live ? __ldg(src + j) : fill;
  }
}

// torchaudio mask_along_axis for time then frequency (data_loader.py:286-287), explicit intervals
// grid = (frame chunks, row chunks, clips): rows and frames are walked by their own indices (no division); a masked row is
// written without being read, an in-place call only touches the masked cells
__global__ void specaug_apply_kernel(const float* __restrict__ in, float* __restrict__ out, int32_t n_rows,
                                     int32_t n_frames, const int32_t* __restrict__ mask_params, float mask_value) {
  const int b = blockIdx.z;
  const int4 mk = __ldg(reinterpret_cast<const int4*>(mask_params) + b);
  const int64_t per_clip = static_cast<int64_t>(n_rows) * n_frames;
  const float* src = in + b * per_clip;
  float* dst = out + b * per_clip;
  const bool in_place = src == dst;
  for (int r = blockIdx.y; r < n_rows; r += gridDim.y) {
    const bool rowmask = r >= mk.z && r < mk.w;
    const float* s = src + static_cast<int64_t>(r) * n_frames;
    float* d = dst + static_cast<int64_t>(r) * n_frames;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_frames; t += gridDim.x * blockDim.x) {
      const bool masked = rowmask || (t >= mk.x && t < mk.y);
      if (masked) d[t] = mask_value;
      else if (!in_place) d[t] = __ldg(s + t);
    }
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
  // kRows rows per pass, every load of the pass requested before the first store.  The position of a row inside its sequence
  // is carried along (one 64-bit modulo per thread, then add-and-wrap): row % seq per 16-byte vector was ~100 instructions of
  // 64-bit division and held the kernel at 0.71 of the HBM peak on instruction issue.
  constexpr int kRows = 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.y;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.y + threadIdx.y;
  int sq_next = static_cast<int>(first % seq);
  const int step = static_cast<int>(stride % seq);
  for (int64_t row0 = first; row0 < n_rows; row0 += kRows * stride) {
    int sq_of[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      sq_of[k] = sq_next;
      sq_next += step;
      if (sq_next >= seq) sq_next -= seq;
    }
    for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) {
      const int c0 = v * kPer;
      const bool col_all = c0 >= f0 && c0 + kPer <= f1;
      const bool col_none = c0 + kPer <= f0 || c0 >= f1;
      uint4 x[kRows];
      bool masked[kRows], live[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const int64_t row = row0 + k * stride;
        live[k] = row < n_rows;
        masked[k] = col_all || (sq_of[k] >= t0 && sq_of[k] < t1);
        // an untouched vector of an in-place call is neither read nor written
        if (live[k] && !masked[k] && !(in_place && col_none)) x[k] = __ldcs(in + row * vec_per_row + v);
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        if (!live[k]) continue;
        uint4* dst = out + (row0 + k * stride) * vec_per_row + v;
        if (masked[k]) *dst = fill4;
        else if (col_none) { if (!in_place) *dst = x[k]; }
        else *dst = patch_vector<ElemT>(x[k], c0, f0, f1, fill);
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

int wft_debug_set_augment_generic(int32_t on) {
  g_debug_aug_generic = on != 0;
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
  {
    const int gx = grid_1d(len_out * inner, 256);
    const int64_t gy_want = outer < 1 ? 1 : outer, gy_cap = (148 * 16 + gx - 1) / gx;
    const dim3 grid(gx, static_cast<unsigned>(gy_want < gy_cap ? gy_want : (gy_cap < 1 ? 1 : gy_cap)));
    pad_or_trim_kernel<<<grid, 256, 0, stream>>>(in, out, outer, len_in, inner, len_out, static_cast<const uint32_t*>(scratch));
  }
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
  const int gx = grid_1d(n_frames, 256);
  int gy = (148 * 16 + gx * batch - 1) / (gx * batch);
  gy = gy < 1 ? 1 : (gy > n_rows ? n_rows : gy);
  dim3 grid(gx, gy, batch);
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
                          const wft::AugDraw& draw, cudaStream_t stream, const wft::AugFix* fix = nullptr) {
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
  const int per_cta = wft::kAugThreads * wft::kAugFramesPerThread;
  // always a programmatic dependent: the kernel waits for its predecessor on the device (griddepcontrol.wait), so its
  // launch latency and -- for the drawn variant -- its draws hide under the tail of the kernel that produces `in`
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((n_frames + per_cta - 1) / per_cta, (n_rows + wft::kAugRowsPerCta - 1) / wft::kAugRowsPerCta, batch);
  cfg.blockDim = dim3(wft::kAugThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const wft::AugFix fx = fix != nullptr ? *fix : wft::AugFix{};
  // (same rule as the fix-up grid's heavy instance: lengths / cuts / an output longer than the clip)
  const bool ragged = fix != nullptr && (fix->lengths != nullptr || fix->n_valid != nullptr || n_frames > fix->n_frames);
  const int fixmode = fix == nullptr ? 0 : (ragged ? 2 : 1);
  // the staged instance (source windows by bulk copy through shared memory) needs 16-byte rows; g_debug_aug_generic forces the
  // generic instance (tests compare the two bit for bit)
  const bool stage = (n_frames & 3) == 0 && n_frames >= 8 && !g_debug_aug_generic;
  if (stage) {
    cfg.gridDim = dim3((n_frames + wft::kStgBlock - 1) / wft::kStgBlock, (n_rows + wft::kStgRows - 1) / wft::kStgRows, batch);
    cfg.blockDim = dim3(wft::kStgThreads);
    cfg.dynamicSmemBytes = wft::kStgSmemBytes;
  }
#define WFT_AUG_LAUNCH(F32, FIX)                                                                                              \
  do {                                                                                                                        \
    if (stage) {                                                                                                              \
      static bool attr_set[64] = {};                                                                                          \
      int dev_ = 0;                                                                                                           \
      WFT_CUDA(cudaGetDevice(&dev_));                                                                                         \
      if (dev_ >= 0 && dev_ < 64 && !attr_set[dev_]) {                                                                        \
        WFT_CUDA(cudaFuncSetAttribute(wft::augment_staged_kernel<F32, FIX>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                      wft::kStgSmemBytes));                                                                   \
        attr_set[dev_] = true;                                                                                                \
      }                                                                                                                       \
      WFT_CUDA(cudaLaunchKernelEx(&cfg, wft::augment_staged_kernel<F32, FIX>, in, out, n_rows, n_frames, warp_params,         \
                                  mask_params, extremes, mask_value, draw, fx));                                             \
    } else {                                                                                                                  \
      WFT_CUDA(cudaLaunchKernelEx(&cfg, wft::augment_kernel<F32, FIX>, in, out, n_rows, n_frames, warp_params, mask_params,   \
                                  extremes, mask_value, draw, fx));                                                           \
    }                                                                                                                         \
  } while (0)
  if (fixmode == 0) {
    if (spline_f32) WFT_AUG_LAUNCH(true, 0); else WFT_AUG_LAUNCH(false, 0);
  } else if (fixmode == 1) {
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
  return launch_augment(in, out, batch, n_rows, n_frames, warp_params, mask_params, extremes, mask_value, spline_f32, wft::AugDraw{},
                        static_cast<cudaStream_t>(stream_));
}

int wft_augment_drawn_f32(const float* in, float* out, int32_t batch, int32_t n_rows, int32_t n_frames, uint64_t seed,
                          uint64_t clip_offset, int32_t time_mask_param, int32_t freq_mask_param, int32_t time_warp_w, float p,
                          const int32_t* extremes, float mask_value, int32_t spline_f32, void* stream_) {
  if (!(p >= 0.0f && p <= 1.0f)) return fail(WFT_ERR_INVALID, "spec_augment p must be between 0 and 1");
  if (time_warp_w < 0) return fail(WFT_ERR_INVALID, "time_warp_w must be >= 0");
  wft::AugDraw d{};
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
  wft::AugDraw d{};
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
  wft::AugFix fix{};
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
  wft::time_warp_draw_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(seed, clip_offset, batch, n_frames, time_warp_w, p,
                                                                 warp_params_out);
  ++g_launches;
  WFT_CUDA(cudaGetLastError());
  return WFT_OK;
}

}  // extern "C"
