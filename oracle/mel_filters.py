"""ORACLE (test infrastructure, never shipped): Slaney mel filter bank.

Restates ``librosa.filters.mel(sr=16000, n_fft=400, n_mels=n)`` -- the recipe that produced
``whisper/assets/mel_filters.npz`` loaded by ``whisper.audio.mel_filters`` (third-party, un-vendored;
reference call site ``/root/reference/src/whisper_finetune/data/data_loader.py:278``).

Steps (librosa 0.9/0.10 ``filters.mel`` with ``htk=False, norm="slaney", dtype=float32``):
  1. ``fftfreqs = linspace(0, sr/2, 1 + n_fft//2)``                          (float64)
  2. ``mel_f    = mel_to_hz(linspace(hz_to_mel(0), hz_to_mel(sr/2), n_mels+2))`` (Slaney scale, float64)
  3. triangle ``max(0, min(lower, upper))`` per row, stored into a float32 array
  4. ``weights *= 2 / (mel_f[2:] - mel_f[:-2])``  (in-place on the float32 array: f64 product, f32 store)

Parity: the ``.npz`` itself is not on disk ("parity unpinned"); pinned instead against
``transformers.audio_utils.mel_filter_bank(201, n, 0, 8000, 16000, "slaney", "slaney")`` (<= 1 f32 ulp).
"""
import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
N_BINS = N_FFT // 2 + 1

_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mel = f / _F_SP
    log_t = f >= _MIN_LOG_HZ
    return np.where(log_t, _MIN_LOG_MEL + np.log(np.maximum(f, 1e-300) / _MIN_LOG_HZ) / _LOGSTEP, mel)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f = _F_SP * m
    log_t = m >= _MIN_LOG_MEL
    return np.where(log_t, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), f)


def mel_filters(n_mels: int) -> np.ndarray:
    """float32 [n_mels, 201] filter bank; upstream asserts n_mels in {80, 128}."""
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    fftfreqs = np.linspace(0.0, SAMPLE_RATE / 2.0, N_BINS)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(SAMPLE_RATE / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, N_BINS), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights
