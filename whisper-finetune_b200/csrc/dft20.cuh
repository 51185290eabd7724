// 20-point complex DFT held entirely in registers (prime-factor 4 x 5, no internal twiddles).
//
// Building block of the 400-point frame transform of the fused Whisper front end: 400 = 20 x 20
// Cooley-Tukey, each of the two stages being a batch of these.  Forward sign convention exp(-i...) like
// torch.stft (reference path: whisper.audio.log_mel_spectrogram, data_loader.py:278).
//
// Good-Thomas index maps (gcd(4,5)=1):  n = (5a + 4b) mod 20,  k = (5ka + 16kb) mod 20
//   X[k] = sum_a sum_b x[n(a,b)] W4^(a ka) W5^(b kb)
// so 5 radix-4 butterflies (over a) feed 4 radix-5 butterflies (over b); 224 flops-instructions total.
#pragma once

namespace wft {

struct cpx {
  float re, im;
};

__device__ __forceinline__ void radix4(const float ar, const float ai, const float br, const float bi,
                                       const float cr, const float ci, const float dr, const float di,
                                       float (&yr)[4], float (&yi)[4]) {
  // inputs x0..x3 = a,b,c,d ; y[k] = sum_n x[n] (-i)^(nk)
  const float t0r = ar + cr, t0i = ai + ci;
  const float t1r = ar - cr, t1i = ai - ci;
  const float t2r = br + dr, t2i = bi + di;
  const float t3r = br - dr, t3i = bi - di;
  yr[0] = t0r + t2r; yi[0] = t0i + t2i;
  yr[2] = t0r - t2r; yi[2] = t0i - t2i;
  yr[1] = t1r + t3i; yi[1] = t1i - t3r;   // t1 - i t3
  yr[3] = t1r - t3i; yi[3] = t1i + t3r;   // t1 + i t3
}

__device__ __forceinline__ void radix5(const float (&xr)[5], const float (&xi)[5], float (&yr)[5], float (&yi)[5]) {
  constexpr float C1 = 0.309016994374947424f;   // cos(2pi/5)
  constexpr float C2 = -0.809016994374947424f;  // cos(4pi/5)
  constexpr float S1 = 0.951056516295153572f;   // sin(2pi/5)
  constexpr float S2 = 0.587785252292473129f;   // sin(4pi/5)
  const float t1r = xr[1] + xr[4], t1i = xi[1] + xi[4];
  const float t2r = xr[2] + xr[3], t2i = xi[2] + xi[3];
  const float t3r = xr[1] - xr[4], t3i = xi[1] - xi[4];
  const float t4r = xr[2] - xr[3], t4i = xi[2] - xi[3];
  yr[0] = xr[0] + t1r + t2r;
  yi[0] = xi[0] + t1i + t2i;
  const float ar = fmaf(C2, t2r, fmaf(C1, t1r, xr[0])), ai = fmaf(C2, t2i, fmaf(C1, t1i, xi[0]));
  const float br = fmaf(C1, t2r, fmaf(C2, t1r, xr[0])), bi = fmaf(C1, t2i, fmaf(C2, t1i, xi[0]));
  const float cr = fmaf(S2, t4r, S1 * t3r), ci = fmaf(S2, t4i, S1 * t3i);
  const float dr = fmaf(-S1, t4r, S2 * t3r), di = fmaf(-S1, t4i, S2 * t3i);
  // y1 = a - i c, y4 = a + i c, y2 = b - i d, y3 = b + i d
  yr[1] = ar + ci; yi[1] = ai - cr;
  yr[4] = ar - ci; yi[4] = ai + cr;
  yr[2] = br + di; yi[2] = bi - dr;
  yr[3] = br - di; yi[3] = bi + dr;
}

// In-place: (xr, xi)[n] -> (xr, xi)[k].  Everything is compile-time indexed so the arrays live in registers.
__device__ __forceinline__ void dft20(float (&xr)[20], float (&xi)[20]) {
  float ur[4][5], ui[4][5];
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    float yr[4], yi[4];
    const int n0 = (4 * b) % 20, n1 = (5 + 4 * b) % 20, n2 = (10 + 4 * b) % 20, n3 = (15 + 4 * b) % 20;
    radix4(xr[n0], xi[n0], xr[n1], xi[n1], xr[n2], xi[n2], xr[n3], xi[n3], yr, yi);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
      ur[ka][b] = yr[ka];
      ui[ka][b] = yi[ka];
    }
  }
#pragma unroll
  for (int ka = 0; ka < 4; ++ka) {
    float yr[5], yi[5];
    radix5(ur[ka], ui[ka], yr, yi);
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) {
      const int k = (5 * ka + 16 * kb) % 20;
      xr[k] = yr[kb];
      xi[k] = yi[kb];
    }
  }
}

}  // namespace wft
