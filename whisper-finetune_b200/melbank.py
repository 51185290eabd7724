"""Slaney-scale triangular mel filter bank for the Whisper front end (host side, numpy).

Produces the same float32 ``[n_mels, 201]`` matrix that ``whisper.audio.mel_filters(device, n_mels)`` loads from
``assets/mel_filters.npz`` (``librosa.filters.mel(sr=16000, n_fft=400, n_mels)``); the reference consumes it
through ``whisper.audio.log_mel_spectrogram`` at ``src/whisper_finetune/data/data_loader.py:278``.
``csrc/gen_tables.py`` bakes the non-zero taps of this bank into the CUDA epilogue as immediates.

Properties the kernel relies on (asserted by the generator and by tests): every row's support is one
contiguous band, every FFT bin feeds at most two rows, and bins 0 and 200 carry zero weight.
"""
import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
N_BINS = N_FFT // 2 + 1  # 201


def _mel_from_hz(hz: np.ndarray) -> np.ndarray:
    hz = np.asarray(hz, dtype=np.float64)
    lin = hz * (3.0 / 200.0)
    step = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        logpart = 15.0 + np.log(hz / 1000.0) / step
    return np.where(hz >= 1000.0, logpart, lin)


def _hz_from_mel(mel: np.ndarray) -> np.ndarray:
    mel = np.asarray(mel, dtype=np.float64)
    step = np.log(6.4) / 27.0
    return np.where(mel >= 15.0, 1000.0 * np.exp(step * (mel - 15.0)), mel * (200.0 / 3.0))


def slaney_mel_bank(n_mels: int) -> np.ndarray:
    """float32 ``[n_mels, 201]``; ``n_mels`` must be 80 or 128 (the two banks Whisper ships)."""
    if n_mels not in (80, 128):
        raise ValueError(f"Unsupported n_mels: {n_mels}")
    bin_hz = np.linspace(0.0, SAMPLE_RATE / 2.0, N_BINS)
    edges = _hz_from_mel(np.linspace(_mel_from_hz(0.0), _mel_from_hz(SAMPLE_RATE / 2.0), n_mels + 2))
    width = edges[1:] - edges[:-1]
    bank = np.zeros((n_mels, N_BINS), dtype=np.float32)
    for m in range(n_mels):
        rising = (bin_hz - edges[m]) / width[m]
        falling = (edges[m + 2] - bin_hz) / width[m + 1]
        bank[m] = np.maximum(0.0, np.minimum(rising, falling))  # stored as float32
    area_norm = 2.0 / (edges[2:] - edges[:-2])
    # float32 triangle times float64 normaliser, rounded once more to float32 (librosa's in-place multiply)
    return (bank.astype(np.float64) * area_norm[:, None]).astype(np.float32)
