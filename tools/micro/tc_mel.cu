// Micro-benchmark: the mel projection of the fused front end on the 5th-generation tensor cores (tcgen05), to put NUMBERS
// next to the "FMA pipes or tensor cores?" decision of DESIGN.md section 3 (BASELINE.json north_star: "decided by ncu").
//
//   mel[128 x 16] = W[128 x 208] . P[208 x 16]        per 16-frame tile (W = 0.25 * Slaney bank, P = power spectrum)
//
// fp32-class accuracy out of tf32 tensor cores needs the split  x = hi + lo  (hi = x with 13 mantissa bits cleared):
//   W.P  ~=  Whi.Phi + Whi.Plo + Wlo.Phi        (3 MMAs per K step, error ~2^-21)
// This is the BEST case for the tensor formulation: one CTA per SM, W resident in TENSOR MEMORY as the A operand
// (2 x 208 columns of the 512; it does not fit shared memory twice over, and not at all at the 6 CTAs / SM the fused kernel
// runs), P staged to shared memory in the canonical K-major no-swizzle layout as the B operand, accumulator in TMEM, read
// back with tcgen05.ld.  26 K-steps x 3 = 78 tcgen05.mma (M=128, N=16, K=8, kind::tf32) per tile, issued by one thread.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/micro/tc_mel tools/micro/tc_mel.cu
//   tools/micro/tc_mel [tiles per CTA]          prints cycles per tile (MMA only, and with P staging + read-back), the
//                                               max relative error against an fp32 FMA evaluation, and the FMA-path time
//   cuobjdump -sass tools/micro/tc_mel | grep -E "UTC|LDTM|STTM"      the SASS evidence (profiles/r02_tc_mel_micro.md)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int kM = 128, kN = 16, kK = 208, kKStep = 8;
constexpr int kThreads = 128;
constexpr int kColsWhi = 0, kColsWlo = kK, kColsD = 2 * kK;   // TMEM columns: W hi [0,208), W lo [208,416), D [416,432)
constexpr int kTmemCols = 512;
// B operand (P^T as [N=16][K=208], K-major): core matrix = 8 rows x 16 bytes, K-adjacent cores 128 B apart (LBO), the two
// 8-row groups kSbo bytes apart
constexpr int kLbo = 128, kSbo = (kK / 4) * 128;
constexpr int kPBytes = 2 * kSbo;   // one (hi or lo) copy of the P tile: 13312 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);            // start address
  d |= static_cast<uint64_t>(kLbo >> 4) << 16;                    // leading (K) byte offset
  d |= static_cast<uint64_t>(kSbo >> 4) << 32;                    // stride (N) byte offset
  d |= 1ull << 46;                                                // descriptor version (sm_100)
  return d;                                                       // layout type 0 = no swizzle
}
// kind::tf32, D = f32, A / B = tf32, both K-major, N = 16, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kN >> 3) << 17) | ((kM >> 4) << 24);

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
tc_mel_kernel(const float* __restrict__ W, const float* __restrict__ P, float* __restrict__ out, int tiles_per_cta,
              unsigned long long* __restrict__ cycles, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sm_phi = smem;
  uint8_t* sm_plo = smem + kPBytes;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(8) uint64_t mma_bar;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_slot;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(32 * warp) << 16);   // this warp's 32 TMEM lanes

  // W row `tid` -> TMEM lane `tid`: hi in columns [0, 208), lo in [208, 416)
  for (int c = 0; c < kK; c += 8) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float w = W[tid * kK + c + i];
      const float h = tf32_hi(w);
      hi[i] = __float_as_uint(h);
      lo[i] = __float_as_uint(w - h);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + kColsWhi + c),
                 "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + kColsWlo + c),
                 "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  unsigned long long mma_cycles = 0, all_cycles = 0;
  uint32_t parity = 0;
  const long long t_all0 = clock64();
  for (int it = 0; it < tiles_per_cta; ++it) {
    const int tile = blockIdx.x * tiles_per_cta + it;
    // stage P (this tile: [208 bins][16 frames] in global) as hi / lo into the canonical B layout
    const float* pt = P + static_cast<size_t>(tile) * kK * kN;
    for (int e = tid; e < kK * kN; e += kThreads) {
      const int k = e / kN, n = e % kN;
      const float v = pt[e];
      const float h = tf32_hi(v);
      const int off = ((n >> 3) * (kK / 4) + (k >> 2)) * 128 + (n & 7) * 16 + (k & 3) * 4;
      *reinterpret_cast<float*>(sm_phi + off) = h;
      *reinterpret_cast<float*>(sm_plo + off) = v - h;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = 0;
    if (tid == 0) {
      t0 = clock64();
      const uint32_t d = tmem + kColsD;
      for (int ks = 0; ks < kK / kKStep; ++ks) {
        const uint64_t bhi = make_b_desc(smem_u32(sm_phi) + ks * 2 * kLbo);
        const uint64_t blo = make_b_desc(smem_u32(sm_plo) + ks * 2 * kLbo);
        mma_ts(d, tmem + kColsWhi + ks * kKStep, bhi, ks > 0 ? 1u : 0u);
        mma_ts(d, tmem + kColsWhi + ks * kKStep, blo, 1u);
        mma_ts(d, tmem + kColsWlo + ks * kKStep, bhi, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar)) : "memory");
    }
    // everybody waits for the accumulator (bounded spin: a wrong descriptor must not hang the box)
    {
      uint32_t done = 0;
      for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&mma_bar)), "r"(parity) : "memory");
      if (!done) {
        if (tid == 0) *err = 1;
        break;
      }
      parity ^= 1;
    }
    if (tid == 0) mma_cycles += static_cast<unsigned long long>(clock64() - t0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(lane_base + kColsD) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* o = out + (static_cast<size_t>(tile) * kM + tid) * kN;
#pragma unroll
    for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4*>(o + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // D and the P tile are free again
  }
  all_cycles = static_cast<unsigned long long>(clock64() - t_all0);
  if (tid == 0) {
    cycles[2 * blockIdx.x] = mma_cycles;
    cycles[2 * blockIdx.x + 1] = all_cycles;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

// the FMA-pipe evaluation of the same dense product (reference for the error, and a time to compare: the fused kernel's
// mel phase uses the SPARSE form of this, 394 of the 26 624 weights)
__global__ void ref_mel_kernel(const float* __restrict__ W, const float* __restrict__ P, float* __restrict__ out, int tiles) {
  const int tile = blockIdx.x, m = threadIdx.x;
  if (tile >= tiles) return;
  const float* pt = P + static_cast<size_t>(tile) * kK * kN;
  float acc[kN];
  for (int n = 0; n < kN; ++n) acc[n] = 0.0f;
  for (int k = 0; k < kK; ++k) {
    const float w = W[m * kK + k];
    if (w != 0.0f)
      for (int n = 0; n < kN; ++n) acc[n] = fmaf(w, pt[k * kN + n], acc[n]);
  }
  for (int n = 0; n < kN; ++n) out[(static_cast<size_t>(tile) * kM + m) * kN + n] = acc[n];
}

int main(int argc, char** argv) {
  const int tiles_per_cta = argc > 1 ? atoi(argv[1]) : 64;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = sms * tiles_per_cta;
  // a banded triangular bank like the Slaney one (row m covers a few bins around 1 + 198 (m / 127)^1.6), times 0.25
  std::vector<float> hW(kM * kK, 0.0f), hP(static_cast<size_t>(tiles) * kK * kN);
  for (int m = 0; m < kM; ++m) {
    const double c = 1.0 + 197.0 * pow(m / 127.0, 1.6), half = 0.6 + 4.0 * m / 127.0;
    for (int k = 1; k < 200; ++k) {
      const double t = 1.0 - fabs(k - c) / half;
      if (t > 0) hW[m * kK + k] = static_cast<float>(0.25 * t / half);
    }
  }
  unsigned s = 12345;
  for (size_t i = 0; i < hP.size(); ++i) {   // power values over 12 decades, like a real spectrum with a loud tone
    s = s * 1664525u + 1013904223u;
    const double u = (s >> 8) * (1.0 / 16777216.0);
    s = s * 1664525u + 1013904223u;
    hP[i] = static_cast<float>(pow(10.0, -8.0 + 12.0 * ((s >> 8) * (1.0 / 16777216.0))) * (0.5 + u));
  }
  float *dW, *dP, *dOut, *dRef;
  unsigned long long* dCyc;
  int* dErr;
  cudaMalloc(&dW, hW.size() * 4);
  cudaMalloc(&dP, hP.size() * 4);
  cudaMalloc(&dOut, static_cast<size_t>(tiles) * kM * kN * 4);
  cudaMalloc(&dRef, static_cast<size_t>(tiles) * kM * kN * 4);
  cudaMalloc(&dCyc, sms * 16);
  cudaMalloc(&dErr, 4);
  cudaMemset(dErr, 0, 4);
  cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dP, hP.data(), hP.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 2 * kPBytes + 1024;
  cudaFuncSetAttribute(tc_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms_tc = 0, ms_ref = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    tc_mel_kernel<<<sms, kThreads, smem>>>(dW, dP, dOut, tiles_per_cta, dCyc, dErr);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tc_mel_kernel: %s\n", cudaGetErrorString(e)); return 1; }
    cudaEventElapsedTime(&ms_tc, e0, e1);
  }
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    ref_mel_kernel<<<tiles, kM>>>(dW, dP, dRef, tiles);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms_ref, e0, e1);
  }
  int herr = 0;
  cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
  std::vector<float> a(static_cast<size_t>(tiles) * kM * kN), b(a.size());
  std::vector<unsigned long long> cyc(2 * sms);
  cudaMemcpy(a.data(), dOut, a.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), dRef, b.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc.data(), dCyc, cyc.size() * 8, cudaMemcpyDeviceToHost);
  double max_rel = 0, max_log = 0;
  size_t bad = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    if (b[i] > 0) {
      const double rel = fabs(static_cast<double>(a[i]) - b[i]) / b[i];
      if (rel > max_rel) max_rel = rel;
      const double dl = fabs(log10(fmax(a[i], 1e-10)) - log10(fmax(b[i], 1e-10))) / 4.0;   // what a feature would move by
      if (dl > max_log) max_log = dl;
      if (rel > 1e-4) ++bad;
    }
  }
  double mma = 0, all = 0;
  for (int i = 0; i < sms; ++i) { mma += cyc[2 * i]; all += cyc[2 * i + 1]; }
  printf("tcgen05 mel micro: %d SMs x %d tiles, mbarrier timeout flag = %d\n", sms, tiles_per_cta, herr);
  printf("  tensor path : %.1f cycles/tile from first tcgen05.mma to the commit's arrival (78 MMAs: M=128 N=16 K=8 tf32)\n", mma / sms / tiles_per_cta);
  printf("                %.1f cycles/tile including P staging (hi/lo split to shared memory) and tcgen05.ld read-back\n", all / sms / tiles_per_cta);
  printf("                kernel %.3f ms = %.3f us per tile per SM\n", ms_tc, ms_tc * 1e3 / tiles_per_cta);
  printf("  accuracy    : max relative error vs fp32 FMA %.3e (%zu cells > 1e-4), max feature error %.3e\n", max_rel, bad, max_log);
  printf("  FMA path    : dense-with-zero-skip reference kernel %.3f ms (%.3f us per tile per SM at 1 tile per CTA)\n", ms_ref, ms_ref * 1e3 * sms / tiles);
  printf("  budget      : the fused kernel spends ~13 400 cycles per tile per CTA (6 CTAs / SM -> ~2 200 cycles per tile per SM) in total\n");
  return herr;
}
