"""SpecAugment time / frequency masks of the reference's feature path, on the GPU.

Mirrors ``torchaudio.transforms.TimeMasking(time_mask_param)`` / ``FrequencyMasking(freq_mask_param)`` exactly as
``AudioDataset`` builds and applies them (``src/whisper_finetune/data/data_loader.py:115-116`` and ``:286-287``):
one interval per call, fill value 0.0, a NEW tensor is returned, and -- for the drop-in callables -- the two
uniforms per mask are drawn from torch's global CPU generator in the same order and with the same float32
interval arithmetic as ``torchaudio.functional.mask_along_axis``, so a seeded reference run and a seeded run of
these classes mask the same cells.  The fill itself is ``wft_specaug_apply_f32`` (include/wft.h).

For batches the intervals are drawn on the device by ``draw_mask_params`` (Philox4x32-10 keyed by
``(seed, global clip index)``), which makes the masks independent of the number of GPUs.
"""
import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from .audio import _stream_ptr, resolve_device


def _interval_from_global_rng(mask_param: int, size: int) -> Tuple[int, int]:
    """torchaudio's draw (functional.py: value = rand*param; min_value = rand*(size - value)), on the host RNG."""
    value = torch.rand(1) * mask_param
    min_value = torch.rand(1) * (size - value)
    start = int(min_value.long())
    return start, start + int(value.long())


def apply_masks(mel: torch.Tensor, mask_params: torch.Tensor, mask_value: float = 0.0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``mel`` CUDA float32 ``[B, R, T]`` (or ``[R, T]``), ``mask_params`` int32 ``[B, 4]`` = (t0, t1, f0, f1)."""
    lib = _lib.load()
    if not mel.is_cuda or mel.dtype != torch.float32:
        raise ValueError("mel must be a CUDA float32 tensor")
    squeeze = mel.dim() == 2
    x = mel.unsqueeze(0) if squeeze else mel
    if x.dim() != 3:
        raise ValueError("mel must be [R, T] or [B, R, T]")
    x = x.contiguous()
    B, R, T = x.shape
    mp = torch.as_tensor(mask_params, dtype=torch.int32).to(x.device).contiguous()
    if tuple(mp.shape) != (B, 4):
        raise ValueError(f"mask_params must have shape {(B, 4)}")
    res = torch.empty_like(x) if out is None else out
    with torch.cuda.device(x.device):
        _lib.check(lib.wft_specaug_apply_f32(x.data_ptr(), res.data_ptr(), B, R, T, mp.data_ptr(),
                                             float(mask_value), _stream_ptr(x.device)))
    return res[0] if squeeze else res


def draw_mask_params(seed: int, clip_offset: int, batch: int, n_mels: int, n_frames: int, time_mask_param: int,
                     freq_mask_param: int, p: float = 1.0, device=None) -> torch.Tensor:
    """Device-side counter-based draw -> int32 ``[batch, 4]`` (``wft_specaug_draw``)."""
    if not 0.0 <= p <= 1.0:
        raise ValueError(f"spec_augment p must be between 0 and 1, got {p}")
    lib = _lib.load()
    dev = resolve_device(device)
    out = torch.empty((batch, 4), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.wft_specaug_draw(ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_uint64(clip_offset), batch,
                                        n_mels, n_frames, time_mask_param, freq_mask_param, float(p),
                                        out.data_ptr(), _stream_ptr(dev)))
    return out


class _AxisMask:
    _axis = -1  # -1 time, -2 frequency

    def __init__(self, mask_param: int):
        self.mask_param = int(mask_param)

    def __call__(self, mel: torch.Tensor, mask_value: float = 0.0) -> torch.Tensor:
        if mel.dim() < 2:
            raise ValueError(f"Spectrogram must have at least two dimensions (time and frequency) ({mel.dim()} given).")
        if self.mask_param < 1:
            return mel
        size = mel.shape[self._axis]
        a, b = _interval_from_global_rng(self.mask_param, size)
        dev = resolve_device(None, mel)
        x = mel.to(dev, torch.float32)
        lead = x.shape[:-2]
        x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
        row = [a, b, 0, 0] if self._axis == -1 else [0, 0, a, b]
        mp = torch.tensor([row] * x3.shape[0], dtype=torch.int32)
        res = apply_masks(x3, mp, mask_value).reshape(*lead, x.shape[-2], x.shape[-1])
        return res if mel.is_cuda else res.to(mel.device)


class TimeMasking(_AxisMask):
    """Drop-in for ``T.TimeMasking(time_mask_param)`` as used at data_loader.py:115,286."""

    _axis = -1

    def __init__(self, time_mask_param: int):
        super().__init__(time_mask_param)


class FrequencyMasking(_AxisMask):
    """Drop-in for ``T.FrequencyMasking(freq_mask_param)`` as used at data_loader.py:116,287."""

    _axis = -2

    def __init__(self, freq_mask_param: int):
        super().__init__(freq_mask_param)
