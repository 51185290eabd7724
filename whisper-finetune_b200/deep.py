"""Deep SpecAugment on encoder activations, on the GPU and without the permutes.

Mirror of ``register_deep_spec_augment_hooks`` (``src/whisper_finetune/model/model_utils.py:382-437``): a forward hook
on ``encoder.blocks[i].attn_ln`` that, while training, masks a time span and a feature span of the normalised
activations ``[batch, seq, dim]``.  The reference does ``permute(0, 2, 1)`` -> ``T.TimeMasking`` ->
``T.FrequencyMasking`` -> ``permute(0, 2, 1)``: two full ``masked_fill`` copies plus a non-contiguous result on every
hooked layer.  Here it is ONE pass of ``wft_mask_bsd`` (include/wft.h) over the tensor in its own layout, in
fp32 / fp16 / bf16, and the backward is the same pass over the gradient.

The two intervals are drawn exactly like torchaudio does for a 3-D input (one mask for the whole batch, two
``torch.rand(1)`` per mask from the global CPU generator, float32 interval arithmetic), time first, then frequency,
so a seeded reference run and a seeded run of these hooks mask the same cells.
"""
from typing import Iterable, Optional, Tuple

import torch

from . import _lib
from . import ops  # noqa: F401  (registers torch.ops.wft.*)
from .augment import _interval_from_global_rng

def mask_activations(x: torch.Tensor, time_span: Tuple[int, int], feature_span: Tuple[int, int]) -> torch.Tensor:
    """``x[b, s, d] := 0`` for ``s`` in ``time_span`` or ``d`` in ``feature_span``; ``x`` is ``[batch, seq, dim]``
    (CUDA, fp32 / fp16 / bf16).  Returns a new tensor; differentiable (``torch.ops.wft.mask_bsd`` registers its backward:
    the same mask over the gradient)."""
    _lib.load()
    if not x.is_cuda:
        raise RuntimeError("whisper-finetune_b200 computes on CUDA only; the activations must be a CUDA tensor")
    return torch.ops.wft.mask_bsd(x, int(time_span[0]), int(time_span[1]), int(feature_span[0]), int(feature_span[1]))


def draw_deep_spans(seq: int, dim: int, time_mask_param: int, freq_mask_param: int) -> Tuple[Tuple[int, int], Tuple[int, int]]:
    """The reference's draws for one hooked layer: TimeMasking on the ``seq`` axis first, FrequencyMasking on ``dim``."""
    t = _interval_from_global_rng(time_mask_param, seq)
    f = _interval_from_global_rng(freq_mask_param, dim)
    return t, f


class DeepSpecAugment:
    """Deep SpecAugment for ONE model: the gate, the two mask widths and the hooks that apply them.

    ``attach(model)`` installs a forward pre-hook on the encoder (rolls the ``p`` gate once per encoder forward, so that a
    gradient-checkpoint recomputation of the same forward sees the same decision) and a forward hook on ``attn_ln`` of the
    selected blocks (masks the normalised activations while training and the gate is open)."""

    def __init__(self, time_mask_param: int, freq_mask_param: int, p: float = 1.0):
        p = float(p)
        if not 0.0 <= p <= 1.0:
            raise ValueError(f"deep_spec_augment p must be between 0 and 1, got {p}")
        self.time_mask_param = int(time_mask_param)
        self.freq_mask_param = int(freq_mask_param)
        self.p = p
        self.gate_open = False
        self.handles = []

    def roll_gate(self) -> bool:
        """p >= 1 / p <= 0 decide without touching the RNG, like the reference; otherwise one ``torch.rand(1)``."""
        if self.p >= 1.0 or self.p <= 0.0:
            return self.p >= 1.0
        return torch.rand(1).item() < self.p

    def on_encoder_forward(self, encoder, args):
        self.gate_open = self.roll_gate()

    def on_attn_ln(self, layer_norm, args, normed):
        if not (layer_norm.training and self.gate_open):
            return normed
        time_span, feature_span = draw_deep_spans(normed.shape[1], normed.shape[2], self.time_mask_param, self.freq_mask_param)
        return mask_activations(normed, time_span, feature_span)

    @staticmethod
    def hooked_blocks(n_blocks: int, layer_indices: Optional[Iterable[int]]):
        """Blocks that get the hook: every block but the last by default (the model needs one clean block to recover);
        an explicit list is taken as given, minus the last block, and must stay in range."""
        wanted = range(n_blocks - 1) if layer_indices is None else list(layer_indices)
        for idx in wanted:
            if idx >= n_blocks:
                raise ValueError(f"Layer index {idx} out of range")
        return [idx for idx in wanted if idx != n_blocks - 1]

    def attach(self, model, layer_indices: Optional[Iterable[int]] = None) -> "DeepSpecAugment":
        blocks = model.encoder.blocks
        for idx in self.hooked_blocks(len(blocks), layer_indices):
            self.handles.append(blocks[idx].attn_ln.register_forward_hook(self.on_attn_ln))
        self.handles.append(model.encoder.register_forward_pre_hook(self.on_encoder_forward))
        return self

    def detach(self) -> None:
        for h in self.handles:
            h.remove()
        self.handles = []


def register_deep_spec_augment_hooks(model, time_mask_param: int, freq_mask_param: int, p: float = 1.0,
                                     layer_indices: Optional[Iterable[int]] = None) -> None:
    """Drop-in for the reference function of this name (model/model_utils.py:382-437): same arguments, same gate, same
    layer selection, same error messages."""
    DeepSpecAugment(time_mask_param, freq_mask_param, p).attach(model, layer_indices)
