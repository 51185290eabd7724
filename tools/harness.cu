// standalone driver: calls the C ABI without torch (fast start under ncu)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/wft.h"
extern "C" int wft_debug_read(int*);
int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 16, nm = argc > 2 ? atoi(argv[2]) : 128, iters = argc > 3 ? atoi(argv[3]) : 4; int ragged = argc > 4 ? atoi(argv[4]) : 0; int selfclean = argc > 5 ? atoi(argv[5]) : 1; int pdl = argc > 6 ? atoi(argv[6]) : 1; int overlap = argc > 7 ? atoi(argv[7]) : 0; int nbuf = argc > 8 ? atoi(argv[8]) : 1;
  size_t n = (size_t)B * 480000;
  std::vector<float> h(n);
  unsigned s = 12345;
  for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 0.2f; }
  float *d_pcm, *d_out; void* ws; size_t wsb = 0;
  cudaMalloc(&d_pcm, n * 4); cudaMemcpy(d_pcm, h.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, (size_t)B * nm * 3000 * 4);
  wft_frontend_workspace_bytes(B, 480000, 3000, &wsb); cudaMalloc(&ws, wsb); cudaMemset(ws, 0, wsb);
  int32_t* d_len = nullptr;
  if (ragged) { std::vector<int32_t> hl(B); unsigned r = 777; for (int i = 0; i < B; ++i) { r = r * 1664525u + 1013904223u; hl[i] = 16000 + (r >> 8) % 464001; }
    cudaMalloc(&d_len, B * 4); cudaMemcpy(d_len, hl.data(), B * 4, cudaMemcpyHostToDevice); }
  wft_frontend_args a{}; a.pcm = d_pcm; a.pcm_dtype = WFT_PCM_F32; a.batch = B; a.clip_stride = 480000; a.n_samples = 480000;
  a.lengths = d_len; a.n_mels = nm; a.n_frames_out = 3000; a.out = d_out; a.workspace = ws; a.workspace_bytes = wsb; a.launch_flags = pdl ? WFT_LAUNCH_PDL : 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < iters; ++it) {
    a.workspace_mode = selfclean == 2 ? WFT_WS_RING + it % WFT_WS_PHASES : selfclean ? (it & 1 ? WFT_WS_PHASE_B : WFT_WS_PHASE_A) : WFT_WS_MEMSET;
    cudaEventRecord(e0);
    int rc = wft_frontend_forward(&a, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("iter %d rc=%d cuda=%s %.1f us (%.3f us/clip)\n", it, rc, cudaGetErrorString(e), ms * 1e3, ms * 1e3 / B);
  }
  if (overlap) {   // independent batches: ring workspace, rotating output buffers, no wait for the launch in front
    const int K = 32;
    std::vector<float*> outs(nbuf, d_out);
    for (int i = 1; i < nbuf; ++i) cudaMalloc(&outs[i], (size_t)B * nm * 3000 * 4);
    a.launch_flags = WFT_LAUNCH_PDL | WFT_LAUNCH_OVERLAP;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      for (int it = 0; it < K; ++it) { a.workspace_mode = WFT_WS_RING + it % WFT_WS_PHASES; a.out = outs[it % nbuf]; int rc = wft_frontend_forward(&a, 0); if (rc) { printf("rc=%d %s\n", rc, wft_last_error()); return 1; } }
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      printf("overlapped x%d: %.1f us per launch (%s)\n", K, ms * 1e3 / K, cudaGetErrorString(e));
    }
    return 0;
  }
  {
    const int K = 20;
    cudaEventRecord(e0);
    for (int it = 0; it < K; ++it) { a.workspace_mode = selfclean == 2 ? WFT_WS_RING + (iters + it) % WFT_WS_PHASES : selfclean ? ((iters + it) & 1 ? WFT_WS_PHASE_B : WFT_WS_PHASE_A) : WFT_WS_MEMSET; wft_frontend_forward(&a, 0); }
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("back-to-back x%d: %.1f us per launch\n", K, ms * 1e3 / K);
  }
#ifdef WFT_DEBUG_SPIN
  int dbg[8]; wft_debug_read(dbg);
  printf("debug: hit=%d tile=%d done=%d counter=%d cta=%d n_ring=%d chain=%d total=%d\n", dbg[0], dbg[1], dbg[2], dbg[3], dbg[4], dbg[5], dbg[6], dbg[7]);
#endif
  return 0;
}
