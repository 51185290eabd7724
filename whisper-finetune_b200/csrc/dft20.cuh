// 20-point complex DFT held entirely in registers (prime-factor 4 x 5, no internal twiddles), written for
// Blackwell's packed fp32 pipe: every complex value is one float2 in an aligned register pair and the butterflies
// are FADD2 / FFMA2 / FMUL2 (add.rn.f32x2 / fma.rn.f32x2 / mul.rn.f32x2, sm_100+) -- 144 instructions instead of
// 224 scalar ones.
//
// Building block of the 400-point frame transform of the fused Whisper front end: 400 = 20 x 20 Cooley-Tukey, each
// of the two stages being a batch of these.  Forward sign convention exp(-i...) like torch.stft (reference path:
// whisper.audio.log_mel_spectrogram, data_loader.py:278).
//
// Good-Thomas index maps (gcd(4,5)=1):  n = (5a + 4b) mod 20,  k = (5ka + 16kb) mod 20
//   X[k] = sum_a sum_b x[n(a,b)] W4^(a ka) W5^(b kb)
// so 5 radix-4 butterflies (over a) feed 4 radix-5 butterflies (over b).
#pragma once

#include <cuda_runtime.h>

namespace wft {

typedef float2 cpx;  // (re, im)

__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ cpx cfma(cpx a, cpx b, cpx c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ cpx splat(float s) { return make_float2(s, s); }
// a - i b = (a.x + b.y, a.y - b.x)   and   a + i b = (a.x - b.y, a.y + b.x)
__device__ __forceinline__ cpx add_mi(cpx a, cpx b) { return cfma(make_float2(b.y, b.x), make_float2(1.0f, -1.0f), a); }
__device__ __forceinline__ cpx add_pi(cpx a, cpx b) { return cfma(make_float2(b.y, b.x), make_float2(-1.0f, 1.0f), a); }

// y[k] = sum_n x[n] (-i)^(nk)
__device__ __forceinline__ void radix4(cpx a, cpx b, cpx c, cpx d, cpx (&y)[4]) {
  const cpx t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
  y[0] = cadd(t0, t2);
  y[2] = csub(t0, t2);
  y[1] = add_mi(t1, t3);
  y[3] = add_pi(t1, t3);
}

__device__ __forceinline__ void radix5(const cpx (&x)[5], cpx (&y)[5]) {
  constexpr float C1 = 0.309016994374947424f;   // cos(2pi/5)
  constexpr float C2 = -0.809016994374947424f;  // cos(4pi/5)
  constexpr float S1 = 0.951056516295153572f;   // sin(2pi/5)
  constexpr float S2 = 0.587785252292473129f;   // sin(4pi/5)
  const cpx t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
  y[0] = cadd(cadd(x[0], t1), t2);
  const cpx a = cfma(splat(C2), t2, cfma(splat(C1), t1, x[0]));
  const cpx b = cfma(splat(C1), t2, cfma(splat(C2), t1, x[0]));
  const cpx c = cfma(splat(S2), t4, cmul(splat(S1), t3));
  const cpx d = cfma(splat(-S1), t4, cmul(splat(S2), t3));
  // y1 = a - i c, y4 = a + i c, y2 = b - i d, y3 = b + i d
  y[1] = add_mi(a, c);
  y[4] = add_pi(a, c);
  y[2] = add_mi(b, d);
  y[3] = add_pi(b, d);
}

// In-place: x[n] -> x[k].  Everything is compile-time indexed so the array lives in registers.
__device__ __forceinline__ void dft20(cpx (&x)[20]) {
  cpx u[4][5];
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    cpx y[4];
    radix4(x[(4 * b) % 20], x[(5 + 4 * b) % 20], x[(10 + 4 * b) % 20], x[(15 + 4 * b) % 20], y);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) u[ka][b] = y[ka];
  }
#pragma unroll
  for (int ka = 0; ka < 4; ++ka) {
    cpx y[5];
    radix5(u[ka], y);
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) x[(5 * ka + 16 * kb) % 20] = y[kb];
  }
}

}  // namespace wft
