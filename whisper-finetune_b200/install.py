"""Install the GPU front end under the names the reference imports.

The reference binds the path by module-level names -- ``from whisper.audio import ... log_mel_spectrogram``
(``src/whisper_finetune/data/data_loader.py:13``) and ``from whisper_finetune.data.utils import pad_or_trim``
(``:16-20``) -- and its tests patch ``data_loader_module.log_mel_spectrogram`` by name
(``tests/test_data_loader.py:203-207``).  ``install()`` rebinds exactly those attributes.
"""
import sys
from types import ModuleType
from typing import Optional


class _TransformsNamespace:
    """What ``data_loader.T`` is rebound to by ``install(patch_masks=True)``: ``torchaudio.transforms`` with the two mask
    classes swapped.  The real module is NOT touched -- ``model/model_utils.py`` builds its deep-SpecAugment hooks from the
    same ``torchaudio.transforms`` (``T.TimeMasking`` on activations that require grad), and every other importer of
    torchaudio keeps the stock classes."""

    def __init__(self, transforms_module, time_masking, frequency_masking):
        self._module = transforms_module
        self.TimeMasking = time_masking
        self.FrequencyMasking = frequency_masking

    def __getattr__(self, name):
        return getattr(self._module, name)


def install(data_loader_module: Optional[ModuleType] = None, patch_whisper_audio: bool = True,
            patch_masks: bool = False) -> None:
    from . import audio, augment

    if patch_whisper_audio:
        wa = sys.modules.get("whisper.audio")
        if wa is not None:
            wa.log_mel_spectrogram = audio.log_mel_spectrogram
    dl = data_loader_module or sys.modules.get("whisper_finetune.data.data_loader")
    if dl is not None:
        dl.log_mel_spectrogram = audio.log_mel_spectrogram
        dl.pad_or_trim = audio.pad_or_trim
        if patch_masks and hasattr(dl, "T") and not isinstance(dl.T, _TransformsNamespace):
            dl.T = _TransformsNamespace(dl.T, augment.TimeMasking, augment.FrequencyMasking)
        if patch_masks:
            if hasattr(dl, "TimeWarpAugmenter"):
                dl.TimeWarpAugmenter = augment.TimeWarpAugmenter
            if hasattr(dl, "ExtremesFrequencyMasking"):
                dl.ExtremesFrequencyMasking = augment.ExtremesFrequencyMasking
    du = sys.modules.get("whisper_finetune.data.utils")
    if du is not None:
        du.pad_or_trim = audio.pad_or_trim
