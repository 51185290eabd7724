"""Reference-facing functions of the front end: ``log_mel_spectrogram`` and ``pad_or_trim``.

Same names, argument meaning and error behaviour as the callables the reference imports:

* ``whisper.audio.log_mel_spectrogram(audio, n_mels=80, padding=0, device=None)``
  (imported at ``src/whisper_finetune/data/data_loader.py:13``, called at ``:278``);
* ``whisper_finetune.data.utils.pad_or_trim(array, length=N_SAMPLES, *, axis=-1)``
  (``src/whisper_finetune/data/utils.py:380-404``, called at ``data_loader.py:282``).

Both are thin wrappers over the C ABI (``include/wft.h``): tensors provide device memory and the stream,
nothing else.  CUDA only -- there is no CPU path.
"""
from typing import Optional, Union

import numpy as np
import torch

from . import _lib
from . import ops  # (registers torch.ops.wft.*)

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000 samples in a 30-second chunk
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000 frames in a mel spectrogram input


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def resolve_device(device: Optional[Union[str, torch.device]], like: Optional[torch.Tensor] = None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("whisper-finetune_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    if device is not None:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"whisper-finetune_b200 computes on CUDA only, got device={device!r}")
        return torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
    if like is not None and like.is_cuda:
        return like.device
    return torch.device("cuda", torch.cuda.current_device())


def _as_pcm_tensor(audio, device: torch.device) -> torch.Tensor:
    if isinstance(audio, str):
        raise NotImplementedError("loading audio from a path (ffmpeg) is not part of the front end; pass PCM samples")
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio))
    if audio.dtype not in (torch.float32, torch.int16):
        if audio.dtype.is_floating_point:
            audio = audio.to(torch.float32)
        else:
            raise TypeError(f"PCM must be float32 or int16, got {audio.dtype}")
    audio = audio.to(device, non_blocking=True)
    if audio.dim() >= 1 and audio.stride(-1) != 1:
        audio = audio.contiguous()
    return audio


def as_i32_on(t, dev, shape, name):
    """Optional per-clip metadata (list / ndarray / tensor) -> contiguous int32 tensor on ``dev`` with ``shape``, or None."""
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t), dtype=torch.int32)
    t = t.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
    if tuple(t.shape) != shape:
        raise ValueError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
    return t


def frontend_forward(pcm: torch.Tensor, n_mels: int, padding: int = 0, lengths: Optional[torch.Tensor] = None,
                     n_frames_out: int = 0, n_valid_frames: Optional[torch.Tensor] = None,
                     mask_params: Optional[torch.Tensor] = None, mask_value: float = 0.0,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One launch of ``wft_frontend_forward`` (``wft::frontend_forward`` / ``wft::frontend_forward_out``, see ``ops.run_eager``) on a CUDA ``[B, N]`` float32 /
    int16 batch -> ``[B, n_mels, T]``."""
    _lib.load()   # fail loudly, before any tensor work, if the CUDA library is missing
    if n_mels not in (80, 128):
        raise ValueError(f"Unsupported n_mels: {n_mels}")
    if not pcm.is_cuda or pcm.dim() != 2 or pcm.stride(1) != 1:
        raise ValueError("pcm must be a CUDA tensor of shape [B, N] with unit stride along N")
    if pcm.dtype not in (torch.float32, torch.int16):
        raise ValueError(f"pcm must be float32 or int16, got {pcm.dtype}")
    B, N = pcm.shape
    if B < 1:
        raise ValueError("empty batch")
    dev = pcm.device

    def _i32(t, name, shape):
        return as_i32_on(t, dev, shape, name)

    lengths = _i32(lengths, "lengths", (B,))
    n_valid_frames = _i32(n_valid_frames, "n_valid_frames", (B,))
    mask_params = _i32(mask_params, "mask_params", (B, 4))
    T = n_frames_out if n_frames_out and n_frames_out > 0 else (N + padding) // HOP_LENGTH
    if out is None:
        return ops.run_eager(ops.frontend_forward, pcm, n_mels, padding, lengths, T, n_valid_frames, mask_params, float(mask_value))
    if not out.is_cuda or out.dtype != torch.float32 or tuple(out.shape) != (B, n_mels, T) or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous CUDA float32 tensor of shape {(B, n_mels, T)}")
    ops.run_eager(ops.frontend_forward_out, pcm, n_mels, padding, lengths, T, n_valid_frames, mask_params, float(mask_value), out)
    return out


def log_mel_spectrogram(
    audio: Union[str, np.ndarray, torch.Tensor],
    n_mels: int = 80,
    padding: int = 0,
    device: Optional[Union[str, torch.device]] = None,
) -> torch.Tensor:
    """Compute the log-Mel spectrogram of 16 kHz PCM on the GPU.

    Parameters mirror ``whisper.audio.log_mel_spectrogram``:
      audio   -- ``[N]`` samples (ndarray or tensor; float32 in [-1, 1] or int16).  A ``[B, N]`` batch is accepted
                 as an extension and every clip gets its OWN max-8 floor, exactly as if the reference had been
                 called clip by clip (data_loader.py:278 runs inside ``__getitem__``).
      n_mels  -- 80 or 128.
      padding -- zeros appended on the right before the STFT.
      device  -- CUDA device to compute on (default: the tensor's device if it is on CUDA, else the current one).

    Returns a float32 CUDA tensor ``[n_mels, (N + padding) // 160]`` (``[B, n_mels, T]`` for a batch).
    """
    if isinstance(audio, str):
        raise NotImplementedError("loading audio from a path (ffmpeg) is not part of the front end; pass PCM samples")
    if n_mels not in (80, 128):
        raise ValueError(f"Unsupported n_mels: {n_mels}")
    if padding < 0:
        raise ValueError("padding must be >= 0")
    dev = resolve_device(device, audio if torch.is_tensor(audio) else None)
    pcm = _as_pcm_tensor(audio, dev)
    if pcm.dim() == 1:
        return frontend_forward(pcm.unsqueeze(0), n_mels, padding)[0]
    if pcm.dim() == 2:
        return frontend_forward(pcm, n_mels, padding)
    raise ValueError(f"audio must be 1-D (or a 2-D batch), got shape {tuple(pcm.shape)}")


def _pad_or_trim_cuda(x: torch.Tensor, length: int, axis: int) -> torch.Tensor:
    x = x.contiguous()
    shape = list(x.shape)
    outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64)) if axis + 1 < len(shape) else 1
    shape[axis] = length
    return torch.ops.wft.pad_or_trim(x.reshape(outer, x.shape[axis], inner), length).reshape(shape)


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Pad or trim ``array`` to ``length`` along ``axis``, padding with the MINIMUM of the array.

    Mirror of the reference's ``pad_or_trim`` (data/utils.py:380-404; note: min-value pad, not whisper's zero
    pad): tensor in -> tensor out on the same device, ndarray in -> ndarray out, and the input object itself is
    returned when it already has the requested length.  The minimum is reduced on the GPU (no ``.item()`` sync
    for CUDA tensors).  The kernel works in float32: float16 / bfloat16 inputs make the round trip through float32
    exactly (min and copies are exact) and come back in their own dtype; other dtypes raise ``TypeError``.
    """
    ndim = array.ndim
    ax = axis + ndim if axis < 0 else axis
    if not 0 <= ax < ndim:
        raise IndexError(f"axis {axis} out of range for a {ndim}-D array")
    n = array.shape[ax]
    if n == length:
        return array
    is_tensor = torch.is_tensor(array)
    n_elems = array.numel() if is_tensor else array.size
    if n < length and n_elems == 0:
        if is_tensor:
            raise RuntimeError("pad_or_trim: min(): cannot take the minimum of an empty tensor")
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    src = array if is_tensor else torch.from_numpy(np.ascontiguousarray(array))
    if src.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise TypeError(f"pad_or_trim supports float32 / float16 / bfloat16, got {src.dtype}")
    dev = resolve_device(None, src)
    res = _pad_or_trim_cuda(src.to(dev, torch.float32, non_blocking=True), length, ax).to(src.dtype)
    if is_tensor:
        return res if array.is_cuda else res.to(array.device)
    return res.cpu().numpy()
