"""Loader integration: workers hand over raw PCM, the front end runs ONCE per batch on the GPU.

In the reference every DataLoader worker computes the features of one clip on the CPU inside
``AudioDataset.__getitem__`` (``src/whisper_finetune/data/data_loader.py:321-360`` -> ``_calculate_mel`` ``:273-292``),
``collate_fn`` (``:362-367``) stacks them, and ``train_step`` moves ``x`` to the GPU
(``model/model_utils.py:59-62``).  Here the same objects are re-wired so that the features never exist on the host:

* ``deferred_calculate_mel`` replaces ``AudioDataset._calculate_mel``: it still runs the CPU audio augmentations,
  still evaluates the partial-segment cut and the SpecAugment gate (consuming torch's global RNG exactly like the
  reference), but returns a **PCM record** instead of a spectrogram: the 480000 samples (float32, or int16 = half the
  H2D bytes) followed by an 8-element trailer {magic, n_valid_frames + 1, augment flag, length / 1024, length % 1024,
  extremes low rows, extremes high rows, 0}.  A record is a plain 1-D tensor, so it crosses the worker boundary through
  the DataLoader's shared-memory fast path.
* ``pcm_collate_fn`` replaces ``collate_fn``: ``(PcmBatch, y_in, y_out)`` with ``y_in`` / ``y_out`` padded as before.
* ``DeviceFrontEndLoader`` wraps the DataLoader and yields ``(x, y_in, y_out)`` with ``x`` = ``[B, n_mels, 3000]`` already
  on the device: H2D of the PCM on a copy stream, one fused launch (+ the optional warp / extremes launches).

What changes with respect to the reference's seeds: the gate, the audio augmentations and the extremes ratio still come
from the worker's torch RNG, but the time-warp and mask intervals are drawn on the device by Philox keyed by
``(seed, running clip index)`` -- same distributions, different numbers.
"""
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence

from .audio import N_FRAMES, N_SAMPLES
from .frontend import FrontEnd

TRAILER = 8
MAGIC = 23131
_LEN_RADIX = 1024


def encode_pcm_record(audio, n_valid_frames: Optional[int] = None, augment: bool = False, pcm_dtype=torch.float32,
                      extremes: Tuple[int, int] = (0, 0)) -> torch.Tensor:
    """``audio``: 1-D float samples in [-1, 1] (<= 480000) -> record ``[480000 + 8]`` of ``pcm_dtype``."""
    if pcm_dtype not in (torch.float32, torch.int16):
        raise TypeError("pcm_dtype must be torch.float32 or torch.int16")
    a = np.asarray(audio.cpu() if torch.is_tensor(audio) else audio).reshape(-1)
    # an augmentation (time stretch) can leave a few samples more than 30 s: the reference trims the SPECTROGRAM to 3000
    # frames (data_loader.py:281-282), here the audio is cut at 480000 samples -- same frames except the last one, whose
    # window would have seen 40 samples beyond the cut
    a = a[:N_SAMPLES]
    nz = np.flatnonzero(a)
    length = int(nz[-1]) + 1 if nz.size else 0          # trailing zeros are padding: the kernel skips those tiles
    nv = -1 if n_valid_frames is None else min(int(n_valid_frames), N_FRAMES)
    if nv == 0:
        # the reference reaches pad_or_trim with an empty spectrogram and torch.min raises (data/utils.py:380-404)
        raise RuntimeError("min(): cannot pad an empty spectrogram (partial segment starts at 0)")
    rec = torch.zeros(N_SAMPLES + TRAILER, dtype=pcm_dtype)
    if pcm_dtype == torch.float32:
        rec[: a.shape[0]] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    else:
        q = np.clip(np.rint(a.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
        rec[: a.shape[0]] = torch.from_numpy(q)
    trailer = [MAGIC, nv + 1, 1 if augment else 0, length // _LEN_RADIX, length % _LEN_RADIX, int(extremes[0]),
               int(extremes[1]), 0]
    rec[N_SAMPLES:] = torch.tensor(trailer, dtype=pcm_dtype)
    return rec


class PcmBatch:
    """What ``pcm_collate_fn`` puts where the reference's ``x`` was: PCM ``[B, 480000]`` + per-clip int32 metadata."""

    def __init__(self, pcm: torch.Tensor, lengths: torch.Tensor, n_valid_frames: torch.Tensor, augment: torch.Tensor,
                 extremes: torch.Tensor):
        self.pcm, self.lengths, self.n_valid_frames, self.augment, self.extremes = pcm, lengths, n_valid_frames, augment, extremes

    def pin_memory(self):   # DataLoader(pin_memory=True) calls this on custom batch types
        return PcmBatch(self.pcm.pin_memory(), self.lengths.pin_memory(), self.n_valid_frames.pin_memory(),
                        self.augment.pin_memory(), self.extremes.pin_memory())

    def __len__(self):
        return self.pcm.shape[0]


def decode_pcm_records(records: Sequence[torch.Tensor]) -> PcmBatch:
    recs = torch.stack(list(records))
    if recs.dim() != 2 or recs.shape[1] != N_SAMPLES + TRAILER:
        raise ValueError("not a batch of PCM records")
    tr = recs[:, N_SAMPLES:].to(torch.int64)
    if not bool((tr[:, 0] == MAGIC).all()):
        raise ValueError("PCM record trailer is corrupt (magic mismatch)")
    lengths = (tr[:, 3] * _LEN_RADIX + tr[:, 4]).to(torch.int32)
    n_valid = (tr[:, 1] - 1).to(torch.int32)                 # -1 = keep every frame
    return PcmBatch(recs[:, :N_SAMPLES], lengths, n_valid, tr[:, 2].to(torch.int32), tr[:, 5:7].to(torch.int32).contiguous())


_PCM_DTYPE = {"dtype": torch.float32}


def deferred_calculate_mel(self, audio_array, next_partial_segment_start, no_timestamps):
    """Replacement for ``AudioDataset._calculate_mel`` (data_loader.py:273-292): same order of decisions, no features."""
    if self.aud_augment is not None:
        audio_array = self.aud_augment(audio_array, sample_rate=16000)
    n_valid = None
    if no_timestamps and next_partial_segment_start is not None:
        n_valid = int(next_partial_segment_start * self.num_frames_per_second)
    augment = self._should_apply_spec_augment()
    extremes = (0, 0)
    efm = getattr(self, "extreme_freq_masking", None)
    if efm:
        r = torch.rand(1).item()   # the one draw ExtremesFrequencyMasking makes per sample (data/utils.py:173-186)
        extremes = (min(int(round(r * efm.low_freq_range)), self.n_mels), min(int(round(r * efm.high_freq_range)), self.n_mels))
    return encode_pcm_record(audio_array, n_valid, augment, _PCM_DTYPE["dtype"], extremes)


def pcm_collate_fn(data):
    """Replacement for ``collate_fn`` (data_loader.py:362-367) when the dataset yields PCM records."""
    x, y_in, y_out = zip(*data)
    y_in = pad_sequence(y_in, batch_first=True, padding_value=0)
    y_out = pad_sequence(y_out, batch_first=True, padding_value=-100)
    return decode_pcm_records(x), y_in, y_out


class DeviceFrontEndLoader:
    """Iterate ``loader`` (built with ``pcm_collate_fn``) and yield ``(x, y_in, y_out)`` with ``x`` on the GPU.

    The mask / warp draws of a clip are keyed by a GLOBAL clip index so that ranks do not repeat each other: batch ``k`` of
    rank ``r`` owns indices ``clip_offset + (k * world_size + r) * batch .. + batch`` (``rank`` / ``world_size`` default to
    the initialised ``torch.distributed`` group, else 0 / 1)."""

    def __init__(self, loader: Iterable, front_end: FrontEnd, clip_offset: int = 0, rank: Optional[int] = None,
                 world_size: Optional[int] = None):
        import torch.distributed as dist

        self.loader = loader
        self.fe = front_end
        self.clip_offset = int(clip_offset)
        ddp = dist.is_available() and dist.is_initialized()
        self.rank = int(rank) if rank is not None else (dist.get_rank() if ddp else 0)
        self.world_size = int(world_size) if world_size is not None else (dist.get_world_size() if ddp else 1)
        if not 0 <= self.rank < self.world_size:
            raise ValueError(f"Invalid rank {self.rank}, rank should be in the interval [0, {self.world_size - 1}]")
        self._batches = 0
        self._copy_stream = torch.cuda.Stream(device=front_end.device)

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch: PcmBatch):
        dev = self.fe.device
        with torch.cuda.stream(self._copy_stream):
            staged = tuple(t.to(dev, non_blocking=True) for t in
                           (batch.pcm, batch.lengths, batch.n_valid_frames, batch.augment, batch.extremes))
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return staged, ev

    def features(self, batch: PcmBatch) -> torch.Tensor:
        (pcm, lengths, n_valid, augment, extremes), ev = self._stage(batch)
        cur = torch.cuda.current_stream(self.fe.device)
        cur.wait_event(ev)
        for t in (pcm, lengths, n_valid, augment, extremes):
            t.record_stream(cur)
        B = pcm.shape[0]
        offset = self.clip_offset + (self._batches * self.world_size + self.rank) * B
        self._batches += 1
        has_extremes = bool(batch.extremes.any())   # host copy: no device sync
        return self.fe(pcm, lengths=lengths, n_valid_frames=n_valid, clip_offset=offset, augment=augment,
                       extremes=extremes if has_extremes else None)

    def __iter__(self):
        for batch, y_in, y_out in self.loader:
            yield self.features(batch), y_in, y_out


def install_loader(data_loader_module=None, pcm_dtype=torch.float32) -> None:
    """Re-wire the reference's data loader module: ``AudioDataset._calculate_mel`` -> PCM records, ``collate_fn`` ->
    ``pcm_collate_fn``.  ``get_dataloader(...)`` then returns a loader to wrap in ``DeviceFrontEndLoader``."""
    import sys

    if pcm_dtype not in (torch.float32, torch.int16):
        raise TypeError("pcm_dtype must be torch.float32 or torch.int16")
    dl = data_loader_module or sys.modules.get("whisper_finetune.data.data_loader")
    if dl is None:
        raise RuntimeError("whisper_finetune.data.data_loader is not imported; pass the module explicitly")
    _PCM_DTYPE["dtype"] = pcm_dtype
    dl.AudioDataset._calculate_mel = deferred_calculate_mel
    dl.collate_fn = pcm_collate_fn
