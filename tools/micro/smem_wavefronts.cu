// Shared-memory wavefront cost of one warp-wide access pattern on sm_100a (development aid, see tools/README.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/micro/smem_wavefronts tools/micro/smem_wavefronts.cu
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum ./smem_wavefronts
// Every pattern is one launch of ONE warp doing kIters accesses; wavefronts / kIters = cost of the pattern.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <functional>
constexpr int kIters = 1000;
template <int OP>  // 0 LDS.32 1 LDS.64 2 LDS.128 3 STS.32 4 STS.64 5 STS.128
__global__ void probe(const int* __restrict__ lane_off, float* sink, int iters) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 8192; i += 32) sm[i] = i;
  __syncwarp();
  const int off = lane_off[threadIdx.x];  // in floats
  float acc = 0.f;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) { float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + off)) : "memory"); acc += v; }
    if (OP == 1) { float2 v; asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm + off)) : "memory"); acc += v.x + v.y; }
    if (OP == 2) { float4 v; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(sm + off)) : "memory"); acc += v.x + v.y + v.z + v.w; }
    if (OP == 3) asm volatile("st.shared.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(sm + off)), "f"(acc) : "memory");
    if (OP == 4) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"((unsigned)__cvta_generic_to_shared(sm + off)), "f"(acc), "f"(acc) : "memory");
    if (OP == 5) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"((unsigned)__cvta_generic_to_shared(sm + off)), "f"(acc), "f"(acc), "f"(acc), "f"(acc) : "memory");
  }
  if (acc == 123.456f) sink[0] = acc;
}
int main() {
  int* d; float* s; cudaMalloc(&d, 128); cudaMalloc(&s, 4);
  struct Pat { const char* name; int op; std::function<int(int)> f; };
  std::vector<Pat> pats = {
    {"LDS.128 uniform", 2, [](int l) { return 0; }},
    {"LDS.128 4 addrs, quarter-warp uniform (lane/8)*12 floats... distinct banks", 2, [](int l) { return (l / 8) * 44; }},
    {"LDS.128 4 addrs interleaved (lane%4)*44", 2, [](int l) { return (l % 4) * 44; }},
    {"LDS.128 2 addrs by half warp", 2, [](int l) { return (l / 16) * 44; }},
    {"LDS.128 8 addrs (lane/4)*44", 2, [](int l) { return (l / 4) * 44; }},
    {"LDS.128 8 addrs (lane%8)*44", 2, [](int l) { return (l % 8) * 44; }},
    {"LDS.128 16 addrs (lane/2)*4 contiguous", 2, [](int l) { return (l / 2) * 4; }},
    {"LDS.128 distinct contiguous", 2, [](int l) { return l * 4; }},
    {"LDS.128 rows r=lane%20 stride 44 (current twiddle)", 2, [](int l) { return (l % 20) * 44; }},
    {"LDS.64 uniform", 1, [](int l) { return 0; }},
    {"LDS.64 distinct contiguous", 1, [](int l) { return l * 2; }},
    {"LDS.64 4 addrs (lane/8)*22", 1, [](int l) { return (l / 8) * 22; }},
    {"LDS.32 uniform", 0, [](int l) { return 0; }},
    {"LDS.32 distinct", 0, [](int l) { return l; }},
    {"STS.64 contiguous", 4, [](int l) { return l * 2; }},
    {"STS.64 new power: q=lane%8 r=lane/8, r*20+2q", 4, [](int l) { return (l / 8) * 20 + 2 * (l % 8); }},
    {"STS.64 new power PS=16: r*16+2q", 4, [](int l) { return (l / 8) * 16 + 2 * (l % 8); }},
    {"STS.64 new power PS=24", 4, [](int l) { return (l / 8) * 24 + 2 * (l % 8); }},
    {"STS.64 new power PS=40", 4, [](int l) { return (l / 8) * 40 + 2 * (l % 8); }},
    {"STS.64 current power: r=lane%20,q=lane/20: r*20+2q", 4, [](int l) { return (l % 20) * 20 + 2 * (l / 20); }},
    {"STS.64 exchange new: q*900 + r  (float2 index) q=lane%8 r=lane/8", 4, [](int l) { return (l % 8) * 900 + 2 * (l / 8); }},
    {"STS.64 exchange new: q*904", 4, [](int l) { return (l % 8) * 904 + 2 * (l / 8); }},
    {"STS.64 exchange new: q*908", 4, [](int l) { return (l % 8) * 908 + 2 * (l / 8); }},
    {"STS.64 exchange new: q*888", 4, [](int l) { return (l % 8) * 888 + 2 * (l / 8); }},
    {"LDS.128 stageB new: q*900 + r*44, q=lane%8 r=lane/8", 2, [](int l) { return (l % 8) * 900 + (l / 8) * 44; }},
    {"LDS.128 stageB new: q*904 + r*44", 2, [](int l) { return (l % 8) * 904 + (l / 8) * 44; }},
    {"LDS.128 stageB new: q*908 + r*44", 2, [](int l) { return (l % 8) * 908 + (l / 8) * 44; }},
    {"LDS.128 stageB new: q*888 + r*44", 2, [](int l) { return (l % 8) * 888 + (l / 8) * 44; }},
    {"STS.128 mirror new: q*900 + r*44", 5, [](int l) { return (l % 8) * 900 + (l / 8) * 44; }},
    {"STS.128 contiguous", 5, [](int l) { return l * 4; }},
    {"STS.128 uniform", 5, [](int l) { return 0; }},
    {"LDS.32 audio new: q*340 + r, q=lane%8 r=lane/8", 0, [](int l) { return (l % 8) * 340 + (l / 8); }},
    {"LDS.128 window new: (lane/8)*20", 2, [](int l) { return (l / 8) * 20; }},
    {"LDS.128 window alt: (lane%4)*20", 2, [](int l) { return (l % 4) * 20; }},
    {"STS.64 power PS=12", 4, [](int l) { return (l / 8) * 12 + 2 * (l % 8); }},
    {"STS.64 power PS=28", 4, [](int l) { return (l / 8) * 28 + 2 * (l % 8); }},
    {"STS.64 power PS=32", 4, [](int l) { return (l / 8) * 32 + 2 * (l % 8); }},
    {"STS.64 power PS=36", 4, [](int l) { return (l / 8) * 36 + 2 * (l % 8); }},
    {"STS.64 power PS=48", 4, [](int l) { return (l / 8) * 48 + 2 * (l % 8); }},
    {"STS.64 power rfast: r=lane%4 q=lane/4, r*20+2q", 4, [](int l) { return (l % 4) * 20 + 2 * (l / 4); }},
    {"STS.64 exchange rfast: q*900 + 2r", 4, [](int l) { return (l / 4) * 900 + 2 * (l % 4); }},
    {"STS.64 exchange rfast: q*904 + 2r", 4, [](int l) { return (l / 4) * 904 + 2 * (l % 4); }},
    {"STS.64 exchange rfast: q*888 + 2r", 4, [](int l) { return (l / 4) * 888 + 2 * (l % 4); }},
    {"LDS.128 stageB rfast: q*900 + r*44", 2, [](int l) { return (l / 4) * 900 + (l % 4) * 44; }},
    {"LDS.128 stageB rfast: q*904 + r*44", 2, [](int l) { return (l / 4) * 904 + (l % 4) * 44; }},
    {"LDS.128 stageB rfast: q*888 + r*44", 2, [](int l) { return (l / 4) * 888 + (l % 4) * 44; }},
    {"LDS.32 audio rfast: q*340 + r", 0, [](int l) { return (l / 4) * 340 + (l % 4); }},
    {"LDS.128 mel-like: 32 distinct rows stride 20", 2, [](int l) { return l * 20; }},
    {"LDS.128 mel-like: pairs share (l/2)*20", 2, [](int l) { return (l / 2) * 20; }},
    {"LDS.64 4 distinct quarter-uniform stride 20", 1, [](int l) { return (l / 8) * 20; }},
    {"LDS.32 4 distinct quarter-uniform stride 20", 0, [](int l) { return (l / 8) * 20; }},
  };
  for (auto& p : pats) {
    int h[32]; for (int l = 0; l < 32; ++l) h[l] = p.f(l);
    cudaMemcpy(d, h, 128, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto launch = [&]() {
      switch (p.op) {
        case 0: probe<0><<<1, 32, 40000>>>(d, s, kIters); break; case 1: probe<1><<<1, 32, 40000>>>(d, s, kIters); break;
        case 2: probe<2><<<1, 32, 40000>>>(d, s, kIters); break; case 3: probe<3><<<1, 32, 40000>>>(d, s, kIters); break;
        case 4: probe<4><<<1, 32, 40000>>>(d, s, kIters); break; case 5: probe<5><<<1, 32, 40000>>>(d, s, kIters); break;
      }
    };
    launch(); cudaDeviceSynchronize();
    printf("PATTERN %s : %s\n", p.name, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
