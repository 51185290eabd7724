// Development aid: per-phase work and barrier-wait times of the fused kernel, from the %clock stamps of a
// -DWFT_TIMELINE build of libwft_b200.so (see WFT_TL in csrc/frontend_kernel.cuh).
//   nvcc -O2 -o tools/timeline tools/timeline.cu -L<dir of the timeline build> -lwft_b200
//   LD_LIBRARY_PATH=<dir> tools/timeline [batch] [n_mels]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../include/wft.h"
extern "C" int wft_debug_timeline(uint32_t*, int32_t*);
int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 256, nm = argc > 2 ? atoi(argv[2]) : 128;
  size_t n = (size_t)B * 480000;
  std::vector<float> h(n);
  unsigned s = 12345;
  for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 0.2f; }
  float *d_pcm, *d_out; void* ws; size_t wsb = 0;
  cudaMalloc(&d_pcm, n * 4); cudaMemcpy(d_pcm, h.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, (size_t)B * nm * 3000 * 4);
  wft_frontend_workspace_bytes(B, 480000, 3000, &wsb); cudaMalloc(&ws, wsb);
  wft_frontend_args a{}; a.pcm = d_pcm; a.pcm_dtype = WFT_PCM_F32; a.batch = B; a.clip_stride = 480000; a.n_samples = 480000;
  a.n_mels = nm; a.n_frames_out = 3000; a.out = d_out; a.workspace = ws; a.workspace_bytes = wsb;
  for (int it = 0; it < 3; ++it) { wft_frontend_forward(&a, 0); cudaDeviceSynchronize(); }
  int32_t dims[4]; wft_debug_timeline(nullptr, dims);
  const int C = dims[0], I = dims[1], W = dims[2], P = dims[3];
  std::vector<uint32_t> tl((size_t)C * I * W * P);
  wft_debug_timeline(tl.data(), dims);
  auto at = [&](int c, int i, int w, int p) { return tl[(((size_t)c * I + i) * W + w) * P + p]; };
  // absolute view: every stamp relative to the CTA's earliest warp at point 0 of the same iteration (one SM, one clock)
  const char* pname[] = {"0 loop top", "1 audio arrived (mbarrier passed)", "2 gather done -> BAR b", "3 past BAR b (issue)", "4 DFT A + exchange st (+describe) -> BAR c",
                         "5 past BAR c (issue)", "6 DFT B + shuffles -> BAR d", "7 past BAR d (issue)", "8 power stored (+TMA issue, stat loads) -> BAR e",
                         "9 past BAR e (issue)", "10 mel done (+ring check) -> BAR f", "11 past BAR f (issue)", "12 publish / fix-up done"};
  printf("B=%d n_mels=%d: mean cycles since the CTA's first warp entered the iteration; iterations 8..%d of every CTA\n", B, nm, I - 1);
  printf("%-52s %8s %8s %8s %8s %8s   %8s\n", "point", "warp0", "warp1", "warp2", "warp3", "warp4", "last");
  for (int k = 0; k < 13; ++k) {
    double sum[8] = {0}, summax = 0; long cnt = 0;
    for (int c = 0; c < C; ++c) for (int i = 8; i < I; ++i) {
      uint32_t base = at(c, i, 0, 0);
      for (int w = 1; w < W; ++w) if ((int32_t)(at(c, i, w, 0) - base) < 0) base = at(c, i, w, 0);
      double mx = 0;
      for (int w = 0; w < W; ++w) { double d = (double)(int32_t)(at(c, i, w, k) - base); sum[w] += d; mx = std::max(mx, d); }
      summax += mx; ++cnt;
    }
    printf("%-52s", pname[k]);
    for (int w = 0; w < W; ++w) printf(" %8.0f", sum[w] / cnt);
    printf("   %8.0f\n", summax / cnt);
  }
  {
    const int extra[][2] = {{7, 13}, {13, 14}, {14, 8}, {3, 15}, {15, 4}};
    const char* en[] = {"past BAR d -> before TMA issue", "TMA issue (fence + expect_tx + copies)", "after TMA issue -> power stored",
                        "past BAR b -> DFT A + exchange stores done", "publish-finish + describe (thread 0)"};
    printf("intervals of lane 0 (mean cycles)\n");
    for (int k = 0; k < 5; ++k) {
      double sum[8] = {0}; long cnt2 = 0;
      for (int c = 0; c < C; ++c) for (int i = 8; i < I; ++i) { for (int w = 0; w < W; ++w) sum[w] += (double)(int32_t)(at(c, i, w, extra[k][1]) - at(c, i, w, extra[k][0])); ++cnt2; }
      printf("%-52s", en[k]);
      for (int w = 0; w < W; ++w) printf(" %8.0f", sum[w] / cnt2);
      printf("\n");
    }
  }
  double iter = 0; long cnt = 0;
  for (int c = 0; c < C; ++c) for (int i = 8; i < I - 1; ++i) { iter += (double)(uint32_t)(at(c, i + 1, 1, 0) - at(c, i, 1, 0)); ++cnt; }
  printf("tile-to-tile period of one CTA: %.0f cycles\n", iter / cnt);
  return 0;
}
