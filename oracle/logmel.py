"""ORACLE (test infrastructure, never shipped): PCM -> Whisper log-mel, on the CPU.

Restates the published algorithm of ``whisper.audio.log_mel_spectrogram(audio, n_mels, padding, device)``
(openai-whisper >= 20240930, third-party and NOT under /root/reference -> **parity unpinned**, see
``oracle/__init__.py``), anchored on the reference's call site
``/root/reference/src/whisper_finetune/data/data_loader.py:278`` (one 1-D clip at a time, ``device=None``
-> CPU, so the max-8 floor is per clip) and on the zero audio pad at ``data_loader.py:344-346``.

Recipe: zero right-pad by ``padding`` -> periodic Hann(400) -> ``torch.stft(400, 160,
center=True, pad_mode="reflect", onesided)`` -> drop the last frame -> ``abs()**2`` -> mel bank matmul ->
``clamp(1e-10).log10()`` -> ``maximum(x, x.max() - 8)`` -> ``(x + 4) / 4``.

``dtype=torch.float32`` is the reference's arithmetic (this is "the reference's CPU path" that bench.py
times); ``dtype=torch.float64`` is the truth used to size the oracle's own rounding noise.
"""
from typing import Optional, Sequence, Union

import numpy as np
import torch

from .mel_filters import mel_filters

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000

_FILTER_CACHE = {}


def _bank(n_mels: int, dtype: torch.dtype) -> torch.Tensor:
    key = (n_mels, dtype)
    if key not in _FILTER_CACHE:
        _FILTER_CACHE[key] = torch.from_numpy(mel_filters(n_mels)).to(dtype)
    return _FILTER_CACHE[key]


def to_float_pcm(audio: Union[np.ndarray, torch.Tensor]) -> torch.Tensor:
    """int16 PCM -> float32 in [-1, 1) (the ``/ 32768.0`` convention of whisper.audio.load_audio)."""
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio))
    if audio.dtype == torch.int16:
        audio = audio.to(torch.float32) / 32768.0
    return audio


def log_mel_spectrogram(
    audio: Union[np.ndarray, torch.Tensor],
    n_mels: int = 80,
    padding: int = 0,
    dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """One clip ``[N]`` -> ``[n_mels, N // 160]`` (N includes ``padding``)."""
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    x = to_float_pcm(audio).to(dtype)
    assert x.dim() == 1
    if padding > 0:
        x = torch.nn.functional.pad(x, (0, padding))
    win = torch.hann_window(N_FFT, dtype=dtype)
    spec = torch.stft(x, N_FFT, HOP_LENGTH, window=win, return_complex=True)
    power = spec[..., :-1].abs() ** 2
    mel = _bank(n_mels, dtype) @ power
    log_spec = torch.clamp(mel, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0


def log_mel_batch(
    pcm: Union[np.ndarray, torch.Tensor],
    n_mels: int,
    lengths: Optional[Sequence[int]] = None,
    padding: int = 0,
    dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """``[B, N]`` clips -> ``[B, n_mels, (N + padding) // 160]``, each clip on its own (per-clip max).

    ``lengths[b]`` marks the valid prefix of clip ``b``; the rest is replaced by zeros, which is what
    ``np.pad(audio, (0, N_SAMPLES - len), "constant")`` at data_loader.py:346 produces.
    """
    x = to_float_pcm(pcm)
    assert x.dim() == 2
    out = []
    for b in range(x.shape[0]):
        clip = x[b]
        if lengths is not None:
            clip = clip.clone()
            clip[int(lengths[b]) :] = 0
        out.append(log_mel_spectrogram(clip, n_mels, padding, dtype))
    return torch.stack(out)
