// standalone timing of the fused augmentation epilogue (wft_augment_f32) without torch:
//   tools/aug_harness [B] [spline_f32] [iters]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/wft.h"
int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 64, f32 = argc > 2 ? atoi(argv[2]) : 0, iters = argc > 3 ? atoi(argv[3]) : 20;
  const int R = 128, T = 3000;
  size_t n = (size_t)B * R * T;
  const int NB = 4;   // rotating buffers: 4 x 98 MB in + out > L2
  float *in[NB], *out[NB];
  std::vector<float> h(n);
  unsigned s = 1;
  for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) * (1.0f / 16777216.0f); }
  for (int i = 0; i < NB; ++i) { cudaMalloc(&in[i], n * 4); cudaMalloc(&out[i], n * 4); cudaMemcpy(in[i], h.data(), n * 4, cudaMemcpyHostToDevice); }
  std::vector<int32_t> hw(2 * B), hm(4 * B);
  for (int b = 0; b < B; ++b) { hw[2 * b] = 80 + (b * 37) % (T - 160); hw[2 * b + 1] = -80 + (b * 13) % 160; hm[4 * b] = 100 + b; hm[4 * b + 1] = 150 + b; hm[4 * b + 2] = 10; hm[4 * b + 3] = 30; }
  int32_t *dw, *dm; cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&dm, hm.size() * 4);
  cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dm, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 4; ++it) wft_augment_f32(in[it % NB], out[it % NB], B, R, T, dw, dm, nullptr, 0.0f, f32, 0);
  cudaEventRecord(e0);
  for (int it = 0; it < iters; ++it) { int rc = wft_augment_f32(in[it % NB], out[it % NB], B, R, T, dw, dm, nullptr, 0.0f, f32, 0); if (rc) { printf("rc=%d %s\n", rc, wft_last_error()); return 1; } }
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  double bytes = 2.0 * n * 4;
  printf("augment B=%d spline_f32=%d: %.1f us per launch, %.0f GB/s (%s)\n", B, f32, ms * 1e3 / iters, bytes / (ms * 1e-3 / iters) / 1e9, cudaGetErrorString(e));
  // reference: a device-to-device copy of the same buffers (what one read + one write of this size can reach at all)
  for (int it = 0; it < 4; ++it) cudaMemcpyAsync(out[it % NB], in[it % NB], n * 4, cudaMemcpyDeviceToDevice, 0);
  cudaEventRecord(e0);
  for (int it = 0; it < iters; ++it) cudaMemcpyAsync(out[it % NB], in[it % NB], n * 4, cudaMemcpyDeviceToDevice, 0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemcpyAsync D2D of the same %.1f MB: %.1f us per copy, %.0f GB/s\n", n * 4 / 1e6, ms * 1e3 / iters, bytes / (ms * 1e-3 / iters) / 1e9);
  return 0;
}
