"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's deep SpecAugment hook body.

Follows ``_norm_hook`` in ``register_deep_spec_augment_hooks`` (src/whisper_finetune/model/model_utils.py:412-421):
``output`` is ``[batch, seq, dim]``; permute to ``[batch, dim, seq]``, ``T.TimeMasking`` (axis -1 = seq), then
``T.FrequencyMasking`` (axis -2 = dim), permute back.  For a 3-D input torchaudio draws ONE interval per mask for the
whole batch (``iid_masks=False``, ``functional.mask_along_axis``): ``value = rand(1) * param``,
``min_value = rand(1) * (size - value)``, cells ``[floor(min_value), floor(min_value) + floor(value))`` := 0.

Pinned against the real torchaudio transforms in tests/test_oracle_cpu.py (same seed -> identical tensors).
"""
import torch


def _draw(mask_param: int, size: int):
    if mask_param < 1:   # torchaudio returns the input untouched and draws nothing (functional.mask_along_axis)
        return 0, 0
    value = torch.rand(1) * mask_param
    min_value = torch.rand(1) * (size - value)
    start = int(min_value.long())
    return start, start + int(value.long())


def deep_spec_augment(output: torch.Tensor, time_mask_param: int, freq_mask_param: int):
    """-> (masked copy of ``output`` [batch, seq, dim], (t0, t1), (f0, f1)); consumes 4 draws of the global RNG."""
    _, seq, dim = output.shape
    t0, t1 = _draw(time_mask_param, seq)
    f0, f1 = _draw(freq_mask_param, dim)
    # index comparison like torchaudio (a span may start below 0 or end beyond the axis when param > size)
    s = torch.arange(seq)
    d = torch.arange(dim)
    masked = ((s >= t0) & (s < t1))[:, None] | ((d >= f0) & (d < f1))[None, :]
    return output.masked_fill(masked[None], 0), (t0, t1), (f0, f1)
