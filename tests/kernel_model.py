"""numpy float32 model of the fused kernel's algorithm (TEST INFRASTRUCTURE): frame pairs packed as one complex
400-point FFT, 20 x 20 Cooley-Tukey with prime-factor 4 x 5 DFT20s, mirror split of the two real spectra, sparse mel
projection with the 0.25 factor folded in, log2-based log10, per-clip floor.  Used by tests/test_kernel_model_cpu.py to
show on the CPU that the ALGORITHM the CUDA kernel implements meets the parity tolerance against the oracle."""
import numpy as np, torch, sys, time
from oracle.logmel import log_mel_spectrogram
from oracle.mel_filters import mel_filters
f32 = np.float32

def r4(x0,x1,x2,x3):
    t0 = x0 + x2; t1 = x0 - x2; t2 = x1 + x3; t3 = x1 - x3
    # -i*t3 = (t3.im, -t3.re)
    mi = (t3.imag - 1j*t3.real).astype(np.complex64)
    return t0 + t2, t1 + mi, t0 - t2, t1 - mi
C1 = f32(np.cos(2*np.pi/5)); C2 = f32(np.cos(4*np.pi/5)); S1 = f32(np.sin(2*np.pi/5)); S2 = f32(np.sin(4*np.pi/5))
def r5(x0,x1,x2,x3,x4):
    t1 = x1 + x4; t2 = x2 + x3; t3 = x1 - x4; t4 = x2 - x3
    y0 = x0 + t1 + t2
    a = x0 + C1*t1 + C2*t2
    b = x0 + C2*t1 + C1*t2
    c = S1*t3 + S2*t4
    d = S2*t3 - S1*t4
    ic = (-c.imag + 1j*c.real).astype(np.complex64); idd = (-d.imag + 1j*d.real).astype(np.complex64)
    return y0, a - ic, b - idd, b + idd, a + ic

def dft20(x):  # x: [..., 20] complex64 -> [..., 20]
    x = [x[..., n] for n in range(20)]
    u = [[None]*5 for _ in range(4)]  # u[ka][b]
    for b in range(5):
        ins = [x[(5*a + 4*b) % 20] for a in range(4)]
        o = r4(*ins)
        for ka in range(4): u[ka][b] = o[ka]
    out = [None]*20
    for ka in range(4):
        o = r5(*u[ka])
        for kb in range(5): out[(5*ka + 16*kb) % 20] = o[kb]
    return np.stack(out, -1).astype(np.complex64)


WIN = (0.5 - 0.5*np.cos(2*np.pi*np.arange(400)/400)).astype(f32)
WIN_T = torch.hann_window(400).numpy()
TW = np.exp(-2j*np.pi*np.outer(np.arange(20), np.arange(20))/400).astype(np.complex64)  # [k1][n2]

def frames_of(x):
    N = x.shape[0]; nf = N//160
    p = np.concatenate([x[200:0:-1], x, x[N-2:N-202:-1]])
    idx = 160*np.arange(nf)[:,None] + np.arange(400)[None,:]
    return p[idx]  # [nf,400]

def model_logmel(x, n_mels, win=WIN_T):
    x = x.astype(f32)
    fr = frames_of(x) * win[None,:]
    nf = fr.shape[0]
    if nf % 2: fr = np.concatenate([fr, np.zeros((1,400), f32)])
    z = (fr[0::2] + 1j*fr[1::2]).astype(np.complex64)      # [pairs,400]
    # stage A: for n2: x[n1] = z[20 n1 + n2]
    zz = z.reshape(-1, 20, 20)            # [pair, n1, n2]
    A = dft20(np.swapaxes(zz, 1, 2))      # [pair, n2, k1]
    A = (A * TW.T[None]).astype(np.complex64)   # TW.T[n2][k1]
    # stage B: for k1: y[n2] = A[pair, n2, k1]
    Zk = dft20(np.swapaxes(A, 1, 2))      # [pair, k1, k2]
    Z = np.swapaxes(Zk, 1, 2).reshape(-1, 400)  # index k1 + 20 k2
    k = np.arange(1, 200)
    Zm = np.conj(Z[:, 400 - k]); Zp = Z[:, k]
    sa = (Zp + Zm).astype(np.complex64); sb = (Zp - Zm).astype(np.complex64)
    Pa = (sa.real*sa.real + sa.imag*sa.imag).astype(f32); Pb = (sb.real*sb.real + sb.imag*sb.imag).astype(f32)
    P = np.empty((2*Pa.shape[0], 199), f32); P[0::2] = Pa; P[1::2] = Pb
    P = P[:nf]
    W = (mel_filters(n_mels)[:, 1:200] * f32(0.25)).astype(f32)
    mel = np.zeros((nf, n_mels), f32)
    for kk in range(199):
        nz = np.nonzero(W[:, kk])[0]
        for m in nz: mel[:, m] = (mel[:, m] + W[m, kk]*P[:, kk]).astype(f32)
    L = (np.log2(np.maximum(mel, f32(1e-10))).astype(f32) * f32(np.log10(2.0))).astype(f32)
    L = np.maximum(L, L.max() - f32(8.0))
    return ((L + f32(4.0)) * f32(0.25)).astype(f32).T

def signals(N=480000):
    t = np.arange(N)/16000.0
    g = torch.Generator().manual_seed(123)
    out = {}
    out['white'] = (0.1*torch.randn(N, generator=g)).clamp(-1,1).numpy()
    hdr = 0.5*np.sin(2*np.pi*220*t) + 0.2*np.sin(2*np.pi*1375.3*t) + 1e-4*torch.randn(N, generator=g).numpy()
    hdr[t >= 12] = 0; out['hdr'] = hdr.astype(f32)
    sp = np.round(32768*(0.3*np.sin(2*np.pi*313.7*t)*np.exp(-(t % 1)) + 3e-4*torch.randn(N, generator=g).numpy()))
    out['int16'] = (np.clip(sp, -32768, 32767).astype(np.int16).astype(f32)/f32(32768.0))
    out['zeros'] = np.zeros(N, f32)
    imp = np.zeros(N, f32); imp[12345] = 1.0; out['impulse'] = imp
    return out

if __name__ == '__main__':
    for name, x in signals().items():
        for n_mels in (128,):
            ref = log_mel_spectrogram(x, n_mels).numpy()
            tru = log_mel_spectrogram(x, n_mels, dtype=torch.float64).numpy()
            t0 = time.time(); got = model_logmel(x, n_mels); dt = time.time()-t0
            def m(a, b): return np.abs(a-b).max(), np.linalg.norm((a-b).ravel())/max(np.linalg.norm(b.ravel()),1e-30)
            print(f'{name:8s} model-vs-oracle maxabs {m(got,ref)[0]:.2e} relL2 {m(got,ref)[1]:.2e} | oracle-vs-f64 {m(ref,tru)[0]:.2e} {m(ref,tru)[1]:.2e} | model-vs-f64 {m(got,tru)[0]:.2e} {m(got,tru)[1]:.2e}  ({dt:.1f}s)')
