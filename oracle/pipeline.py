"""ORACLE (test infrastructure, never shipped): the reference's per-clip feature pipeline.

Restates the ordering of ``AudioDataset.__getitem__`` / ``_calculate_mel``
(``/root/reference/src/whisper_finetune/data/data_loader.py:344-346`` and ``:273-292``) for the rows of
SURVEY.md section 8(a) that are in scope:

    a1  zero right-pad PCM to 480000            data_loader.py:346
    a2  log_mel_spectrogram(audio, n_mels)      data_loader.py:278      (oracle/logmel.py)
    a3  mel[:, :int(start * 100)]               data_loader.py:279-280
    a4  pad_or_trim(mel, 3000) if needed        data_loader.py:281-282  (oracle/pad_or_trim.py)
    a6  time mask, a7 frequency mask            data_loader.py:286-287  (oracle/specaug.py)
    a8  stack to [B, n_mels, 3000]              data_loader.py:362-367

Time-warp (data_loader.py:285) and the extremes mask (:289-290) are "next" rows and are not
applied here.  Mask parameters are explicit so the CUDA path can be compared bit for bit.
"""
from typing import Optional, Sequence

import numpy as np
import torch

from .logmel import N_FRAMES, N_SAMPLES, log_mel_spectrogram, to_float_pcm
from .pad_or_trim import pad_or_trim
from .specaug import apply_masks


def calculate_mel(audio, n_mels: int, n_valid_frames: Optional[int] = None, mask=None,
                  dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """One clip: PCM (<= 480000 samples) -> [n_mels, 3000] features."""
    x = to_float_pcm(audio)
    if x.shape[0] < N_SAMPLES:
        x = torch.nn.functional.pad(x, (0, N_SAMPLES - x.shape[0]))
    mel = log_mel_spectrogram(x, n_mels=n_mels, dtype=dtype)
    if n_valid_frames is not None:
        mel = mel[:, : int(n_valid_frames)]
    if mel.shape[1] != N_FRAMES:
        mel = pad_or_trim(mel, N_FRAMES)
    if mask is not None:
        t0, t1, f0, f1 = (int(v) for v in mask)
        mel = apply_masks(mel, t0, t1, f0, f1, 0.0)
    return mel


def front_end_batch(pcm, n_mels: int, lengths: Optional[Sequence[int]] = None,
                    n_valid_frames: Optional[Sequence[int]] = None, masks: Optional[np.ndarray] = None,
                    dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """[B, <=480000] PCM (+ per-clip valid lengths) -> [B, n_mels, 3000], clip by clip like the reference."""
    x = to_float_pcm(pcm)
    feats = []
    for b in range(x.shape[0]):
        clip = x[b] if lengths is None else x[b, : int(lengths[b])]
        nv = None if n_valid_frames is None or int(n_valid_frames[b]) < 0 else int(n_valid_frames[b])
        mk = None if masks is None else masks[b]
        feats.append(calculate_mel(clip, n_mels, nv, mk, dtype))
    return torch.stack(feats)
