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
from typing import Optional, Tuple

import torch

from . import _lib
from . import ops  # noqa: F401  (registers torch.ops.wft.*)
from .audio import resolve_device


def _interval_from_global_rng(mask_param: int, size: int) -> Tuple[int, int]:
    """torchaudio's draw (functional.py: value = rand*param; min_value = rand*(size - value)), on the host RNG.
    ``mask_param < 1`` masks nothing and consumes no random numbers, like torchaudio."""
    if mask_param < 1:
        return 0, 0
    value = torch.rand(1) * mask_param
    min_value = torch.rand(1) * (size - value)
    start = int(min_value.long())
    return start, start + int(value.long())


def _checked_out(out: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """An ``out=`` buffer goes to the C ABI as a bare pointer: it has to be exactly what the kernel will write."""
    if (not torch.is_tensor(out) or out.device != like.device or out.dtype != like.dtype
            or tuple(out.shape) != tuple(like.shape) or not out.is_contiguous()):
        raise ValueError(f"out must be a contiguous {like.dtype} tensor of shape {tuple(like.shape)} on {like.device}")
    return out


def apply_masks(mel: torch.Tensor, mask_params: torch.Tensor, mask_value: float = 0.0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``mel`` CUDA float32 ``[B, R, T]`` (or ``[R, T]``), ``mask_params`` int32 ``[B, 4]`` = (t0, t1, f0, f1)."""
    _lib.load()
    if not mel.is_cuda or mel.dtype != torch.float32:
        raise ValueError("mel must be a CUDA float32 tensor")
    squeeze = mel.dim() == 2
    x = mel.unsqueeze(0) if squeeze else mel
    if x.dim() != 3:
        raise ValueError("mel must be [R, T] or [B, R, T]")
    x = x.contiguous()
    B = x.shape[0]
    mp = torch.as_tensor(mask_params, dtype=torch.int32).to(x.device).contiguous()
    if tuple(mp.shape) != (B, 4):
        raise ValueError(f"mask_params must have shape {(B, 4)}")
    if out is None:
        res = torch.ops.wft.specaug_apply(x, mp, float(mask_value))
    else:
        res = _checked_out(out, x)
        if res.data_ptr() == x.data_ptr():
            torch.ops.wft.specaug_apply_(res, mp, float(mask_value))
        else:
            torch.ops.wft.augment_out(x, None, mp, None, float(mask_value), False, res)
    return res[0] if squeeze else res


def draw_mask_params(seed: int, clip_offset: int, batch: int, n_mels: int, n_frames: int, time_mask_param: int,
                     freq_mask_param: int, p: float = 1.0, device=None) -> torch.Tensor:
    """Device-side counter-based draw -> int32 ``[batch, 4]`` (``wft_specaug_draw``)."""
    if not 0.0 <= p <= 1.0:
        raise ValueError(f"spec_augment p must be between 0 and 1, got {p}")
    _lib.load()
    if batch < 1:
        raise ValueError("batch must be >= 1")
    dev = resolve_device(device)
    like = torch.empty(0, dtype=torch.int32, device=dev)
    return torch.ops.wft.specaug_draw(like, int(seed), int(clip_offset), int(batch), int(n_mels), int(n_frames),
                                      int(time_mask_param), int(freq_mask_param), float(p))


class _AxisMask:
    """One torchaudio-style mask along ``_axis``.  The result has the input's dtype and stays on the autograd graph: float32
    spectrograms go through ``wft_specaug_apply_f32`` (any ``mask_value``); fp16 / bf16 tensors and tensors that require
    grad (the reference also applies these transforms to encoder activations, model/model_utils.py:404-405) go through the
    differentiable ``mask_activations`` (``wft_mask_bsd``, fill 0)."""

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
        x = mel.to(dev)
        lead = x.shape[:-2]
        x3 = x.reshape(-1, x.shape[-2], x.shape[-1])
        if x3.dtype == torch.float32 and not x3.requires_grad:
            row = [a, b, 0, 0] if self._axis == -1 else [0, 0, a, b]
            mp = torch.tensor([row] * x3.shape[0], dtype=torch.int32)
            res = apply_masks(x3, mp, mask_value)
        else:
            if mask_value != 0.0:
                raise ValueError("a non-zero mask_value is only supported for float32 tensors that do not require grad")
            from .deep import mask_activations   # [B, S, D]: S is this tensor's frequency axis, D its time axis

            res = mask_activations(x3, (0, 0), (a, b)) if self._axis == -1 else mask_activations(x3, (a, b), (0, 0))
        res = res.reshape(*lead, x.shape[-2], x.shape[-1])
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


# ---- "next" rows: time-warp and extremes mask -------------------------------------------------------

def time_warp(mel: torch.Tensor, warp_params: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``mel`` CUDA float32 ``[B, R, T]`` (or ``[R, T]``), ``warp_params`` int32 ``[B, 2]`` = (warp_p, warp_d) ->
    time-warped copy (cubic Hermite source map, bilinear resampling, zeros outside; ``warp_p <= 0`` copies the clip)."""
    if not mel.is_cuda or mel.dtype != torch.float32:
        raise ValueError("mel must be a CUDA float32 tensor")
    squeeze = mel.dim() == 2
    x = (mel.unsqueeze(0) if squeeze else mel).contiguous()
    if x.dim() != 3:
        raise ValueError("You sure it's a Spectrogram?")
    if x.shape[1] < 2 or x.shape[2] < 3:
        raise ValueError("time warp needs at least 2 rows and 3 frames")
    if warp_params is None:
        raise ValueError("warp_params is required")
    res = augment_epilogue(x, warp_params, None, None, 0.0, out=None if out is None else _checked_out(out, x))
    return res[0] if squeeze else res


def augment_epilogue(mel: torch.Tensor, warp_params: Optional[torch.Tensor] = None,
                     mask_params: Optional[torch.Tensor] = None, extremes: Optional[torch.Tensor] = None,
                     mask_value: float = 0.0, out: Optional[torch.Tensor] = None, spline: str = "f64") -> torch.Tensor:
    """Everything ``_calculate_mel`` does after ``pad_or_trim`` (data_loader.py:284-290), in one pass over the features:
    time-warp -> time mask -> frequency mask -> extremes mask (``wft_augment_f32``).

    ``mel`` CUDA float32 ``[B, R, T]``; ``warp_params`` int32 ``[B, 2]`` (``warp_p <= 0`` = leave the clip alone),
    ``mask_params`` int32 ``[B, 4]``, ``extremes`` int32 ``[B, 2]`` = (rows masked from the bottom, from the top); any of
    them may be ``None``.  ``spline="f32"`` restates the reference's float32 spline arithmetic, ``"f64"`` rounds once."""
    lib = _lib.load()
    if not mel.is_cuda or mel.dtype != torch.float32 or mel.dim() != 3:
        raise ValueError("mel must be a CUDA float32 tensor of shape [B, R, T]")
    if spline not in ("f32", "f64"):
        raise ValueError("spline must be 'f32' or 'f64'")
    x = mel.contiguous()
    B, R, T = x.shape

    def _i32(t, cols, name):
        if t is None:
            return None
        t = torch.as_tensor(t, dtype=torch.int32).to(x.device).contiguous()
        if tuple(t.shape) != (B, cols):
            raise ValueError(f"{name} must have shape {(B, cols)}")
        return t

    wp, mp, ex = _i32(warp_params, 2, "warp_params"), _i32(mask_params, 4, "mask_params"), _i32(extremes, 2, "extremes")
    if out is None:
        return torch.ops.wft.augment(x, wp, mp, ex, float(mask_value), spline == "f32")
    res = _checked_out(out, x)
    if res.data_ptr() == x.data_ptr():
        if wp is not None:
            raise ValueError("time warp cannot run in place")
        torch.ops.wft.augment_(res, mp, ex, float(mask_value))
    else:
        torch.ops.wft.augment_out(x, wp, mp, ex, float(mask_value), spline == "f32", res)
    return res


def draw_warp_params(seed: int, clip_offset: int, batch: int, n_frames: int, time_warp_w: int, p: float = 1.0,
                     device=None) -> torch.Tensor:
    """Device-side counter-based draw of (warp_p, warp_d) -> int32 ``[batch, 2]`` (``wft_time_warp_draw``); (-1, 0) for a
    clip the ``p`` gate rejects."""
    if not 0.0 <= p <= 1.0:
        raise ValueError(f"spec_augment p must be between 0 and 1, got {p}")
    _lib.load()
    if batch < 1:
        raise ValueError("batch must be >= 1")
    dev = resolve_device(device)
    like = torch.empty(0, dtype=torch.int32, device=dev)
    return torch.ops.wft.time_warp_draw(like, int(seed), int(clip_offset), int(batch), int(n_frames), int(time_warp_w), float(p))


class TimeWarpAugmenter:
    """Drop-in for the reference's ``TimeWarpAugmenter(W)`` (data/utils.py:41-143, used at data_loader.py:117,285).

    ``warp_p = randint(W, T - W)`` and ``warp_d = randint(-W, W)`` come from torch's global CPU generator in the
    reference's order, so a seeded run warps by the same amount; the resampling runs on the GPU.  A 3-D input
    ``[C, R, T]`` gets ONE warp for all slices, exactly like the reference (it treats the leading axis as channels)."""

    def __init__(self, W: int = 50):
        self.W = int(W)

    def __call__(self, specs):
        if not torch.is_tensor(specs):
            specs = torch.from_numpy(specs)
        if specs.dim() < 2 or specs.dim() > 3:
            raise ValueError("You sure it's a Spectrogram?")
        T = specs.shape[-1]
        warp_p = int(torch.randint(self.W, T - self.W, (1,)))
        warp_d = int(torch.randint(-self.W, self.W, (1,)))
        dev = resolve_device(None, specs)
        x = specs.to(dev, torch.float32)
        lead = x.shape[0] if x.dim() == 3 else 1
        wp = torch.tensor([[warp_p, warp_d]] * lead, dtype=torch.int32)
        res = time_warp(x, wp)
        return res if specs.is_cuda else res.to(specs.device)


class ExtremesFrequencyMasking:
    """Drop-in for the reference's ``ExtremesFrequencyMasking`` (data/utils.py:146-190, used at data_loader.py:124-130,
    289-290): per sample ONE ``torch.rand(1)`` ratio; the ``round(r * low)`` lowest and ``round(r * high)`` highest mel
    rows are zeroed.  Like the reference it works IN PLACE on tensors and returns its argument."""

    def __init__(self, low_freq_range: int = 10, high_freq_range: int = 10):
        self.low_freq_range = low_freq_range
        self.high_freq_range = high_freq_range

    def __call__(self, specs: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(specs):
            specs = torch.tensor(specs)
        x = specs.unsqueeze(0) if specs.dim() == 2 else specs
        batch, n_mels, _ = x.shape
        rows = []
        for _ in range(batch):
            r = torch.rand(1).item()
            low = min(int(round(r * self.low_freq_range)), n_mels)
            high = min(int(round(r * self.high_freq_range)), n_mels)
            rows.append((low, high))
        if not specs.is_cuda or specs.dtype != torch.float32 or not x.is_contiguous():
            raise ValueError("ExtremesFrequencyMasking expects a contiguous CUDA float32 spectrogram (in-place op)")
        augment_epilogue(x, None, None, torch.tensor(rows, dtype=torch.int32), 0.0, out=x)
        return specs
