"""whisper-finetune_b200: the B200-native (sm_100a) Whisper audio front end.

One hot path of i4Ds/whisper-finetune, rebuilt as hand-written CUDA behind the reference's own call surface:
PCM -> log-mel -> cut / min-pad -> SpecAugment masks -> ``x[B, n_mels, 3000]``
(``src/whisper_finetune/data/data_loader.py:273-292, 344-346, 362-367``).  See DESIGN.md / INTEGRATION.md.
"""
from .audio import (CHUNK_LENGTH, HOP_LENGTH, N_FFT, N_FRAMES, N_SAMPLES, SAMPLE_RATE, frontend_forward,
                    log_mel_spectrogram, pad_or_trim)
from .augment import (ExtremesFrequencyMasking, FrequencyMasking, TimeMasking, TimeWarpAugmenter, apply_masks,
                      augment_epilogue, draw_mask_params, draw_warp_params, time_warp)
from .deep import DeepSpecAugment, draw_deep_spans, mask_activations, register_deep_spec_augment_hooks
from .frontend import FrontEnd, HostPipeline
from .install import install
from .loader import (DeviceFrontEndLoader, PcmBatch, decode_pcm_records, deferred_calculate_mel, encode_pcm_record,
                     install_loader, pcm_collate_fn)
from .melbank import slaney_mel_bank
from .ops import set_overlap, set_programmatic_launch
from .sharding import all_gather_features, shard_indices

__all__ = [
    "SAMPLE_RATE", "N_FFT", "HOP_LENGTH", "CHUNK_LENGTH", "N_SAMPLES", "N_FRAMES",
    "log_mel_spectrogram", "pad_or_trim", "frontend_forward", "FrontEnd", "HostPipeline",
    "TimeMasking", "FrequencyMasking", "apply_masks", "draw_mask_params",
    "TimeWarpAugmenter", "ExtremesFrequencyMasking", "time_warp", "draw_warp_params", "augment_epilogue",
    "mask_activations", "draw_deep_spans", "register_deep_spec_augment_hooks", "DeepSpecAugment",
    "encode_pcm_record", "decode_pcm_records", "PcmBatch", "pcm_collate_fn", "deferred_calculate_mel",
    "DeviceFrontEndLoader", "install_loader",
    "shard_indices", "all_gather_features", "slaney_mel_bank", "install", "set_programmatic_launch", "set_overlap",
]
