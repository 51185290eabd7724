"""ORACLE (test infrastructure, never shipped): SpecAugment time / frequency masks.

Restates ``torchaudio.functional.mask_along_axis`` (torchaudio ``functional/functional.py``, the non-iid
path that ``T.TimeMasking(time_mask_param)`` / ``T.FrequencyMasking(freq_mask_param)`` take for a 2-D mel)
as it is used by the reference at ``/root/reference/src/whisper_finetune/data/data_loader.py:115-116``
(construction) and ``:286-287`` (time mask first, then frequency mask, fill value 0.0), and the gate
``_should_apply_spec_augment`` at ``data_loader.py:294-301``.

Interval formula (all float32, like ``torch.rand(1) * mask_param``):
    value = u_a * mask_param ; min_value = u_b * (size - value)
    start = trunc(min_value) ; end = start + trunc(value)        -> zero ``[start, end)`` on the axis

PINNED: torchaudio itself is installed; ``tests/test_oracle_cpu.py`` replays the global CPU generator and
checks this restatement against the real ``T.TimeMasking`` / ``T.FrequencyMasking`` output bit for bit.

The uniforms come either from a replay of torch's generator (reference behaviour) or from the explicit
counter-based draw below, which is what the CUDA path uses: Philox4x32-10 keyed by ``seed``, counter =
global clip index, ``u = (word >> 8) * 2**-24``; words 0..3 of block 0 are (time width, time start,
freq width, freq start) in the reference's draw order, word 0 of block 1 is the ``p`` gate.
"""
from typing import Tuple

import numpy as np
import torch

_M0 = 0xD2511F53
_M1 = 0xCD9E8D57
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = 0xFFFFFFFF


def philox4x32_10(counter: Tuple[int, int, int, int], key: Tuple[int, int]) -> Tuple[int, int, int, int]:
    """Reference Philox4x32 with 10 rounds (Salmon et al., SC'11), plain Python integers."""
    c0, c1, c2, c3 = [int(c) & _MASK32 for c in counter]
    k0, k1 = [int(k) & _MASK32 for k in key]
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _MASK32
        hi1, lo1 = p1 >> 32, p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK32, lo1, (hi0 ^ c3 ^ k1) & _MASK32, lo0
        k0 = (k0 + _W0) & _MASK32
        k1 = (k1 + _W1) & _MASK32
    return c0, c1, c2, c3


def _u01(word: int) -> np.float32:
    return np.float32(word >> 8) * np.float32(2.0**-24)


def clip_uniforms(seed: int, clip_index: int) -> Tuple[np.float32, ...]:
    """(u_tw, u_ts, u_fw, u_fs, u_gate) for one global clip index."""
    key = (seed & _MASK32, (seed >> 32) & _MASK32)
    lo, hi = clip_index & _MASK32, (clip_index >> 32) & _MASK32
    b0 = philox4x32_10((lo, hi, 0, 0), key)
    b1 = philox4x32_10((lo, hi, 1, 0), key)
    return tuple(_u01(w) for w in b0) + (_u01(b1[0]),)


def interval(u_width: np.float32, u_start: np.float32, mask_param: int, size: int) -> Tuple[int, int]:
    """torchaudio's ``[start, end)`` from two float32 uniforms; empty if mask_param < 1."""
    if mask_param < 1:
        return 0, 0
    value = np.float32(u_width) * np.float32(mask_param)
    min_value = np.float32(u_start) * (np.float32(size) - value)
    start = int(min_value)
    return start, start + int(value)


def draw_mask_params(seed: int, clip_offset: int, batch: int, n_mels: int, n_frames: int,
                     time_mask_param: int, freq_mask_param: int, p: float = 1.0) -> np.ndarray:
    """int32 [batch, 4] = (t0, t1, f0, f1) per clip; all zero when the ``p`` gate rejects the clip."""
    out = np.zeros((batch, 4), dtype=np.int32)
    for b in range(batch):
        u_tw, u_ts, u_fw, u_fs, u_gate = clip_uniforms(seed, clip_offset + b)
        apply = p >= 1.0 or (p > 0.0 and float(u_gate) < np.float32(p))
        if not apply:
            continue
        t0, t1 = interval(u_tw, u_ts, time_mask_param, n_frames)
        f0, f1 = interval(u_fw, u_fs, freq_mask_param, n_mels)
        out[b] = (t0, t1, f0, f1)
    return out


def apply_masks(mel: torch.Tensor, t0: int, t1: int, f0: int, f1: int, fill: float = 0.0) -> torch.Tensor:
    """New tensor with frames ``[t0, t1)`` and mel rows ``[f0, f1)`` set to ``fill`` (masked_fill semantics)."""
    out = mel.clone()
    if t1 > t0:
        out[..., :, t0:t1] = fill
    if f1 > f0:
        out[..., f0:f1, :] = fill
    return out


def torch_rng_mask_params(n_mels: int, n_frames: int, time_mask_param: int, freq_mask_param: int):
    """Draw (t0, t1, f0, f1) from torch's *global CPU generator* in the reference's order
    (time: rand, rand; freq: rand, rand) -- used to replay torchaudio for the bit-exact pin."""
    t0 = t1 = f0 = f1 = 0
    if time_mask_param >= 1:
        v = torch.rand(1) * time_mask_param
        m = torch.rand(1) * (n_frames - v)
        t0 = int(m.long())
        t1 = t0 + int(v.long())
    if freq_mask_param >= 1:
        v = torch.rand(1) * freq_mask_param
        m = torch.rand(1) * (n_mels - v)
        f0 = int(m.long())
        f1 = f0 + int(v.long())
    return t0, t1, f0, f1
