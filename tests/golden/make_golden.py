#!/usr/bin/env python
"""Regenerate the golden fixtures in this directory.  RUN IN THE BUILD CONTAINER ONLY (it reads /root/reference).

What gets pinned, and by what:

* ``pad_or_trim.npz``      -- outputs of the REFERENCE'S OWN ``pad_or_trim`` (src/whisper_finetune/data/utils.py:380-404),
                              imported from /root/reference/src with the same ``whisper`` stub its tests install
                              (tests/test_data_loader.py:12-47).  Tensor and ndarray branches, trim / pad / no-op.
* ``calculate_mel.npz``    -- outputs of the REFERENCE'S OWN ``AudioDataset._calculate_mel`` (data_loader.py:273-292) driven
                              with real ``torchaudio`` masks under fixed torch seeds: pins the a2 -> a3 -> a4 -> a6 -> a7
                              ordering, the min-value pad and the mask draw.  ``whisper.audio.log_mel_spectrogram`` is NOT
                              importable here (third-party, un-vendored: "parity unpinned"), so the oracle restatement
                              stands in for it and time-warp (a "next" row) is the identity.  Stored sub-sampled
                              (every 16th frame) plus whole-tensor sums to keep the fixture small.
* ``logmel_hf.npz``        -- log-mel of short seeded clips from the independent ``transformers`` Whisper feature extractor
                              (feature_extraction_whisper.py ``_np_extract_fbank_features``), the closest published
                              implementation available offline; cross-checks the oracle's STFT / mel / log recipe.
* ``logmel_hf_torch.npz``  -- full 30-s clips (all six signal kinds, 80 and 128 mel) through the float32 torch.stft path of the
                              same extractor (``_torch_extract_fbank_features``): agrees with the oracle to ONE float32 ulp
                              (max-abs <= 2e-7), three orders tighter than the numpy variant above.
* ``timewarp.npz``         -- outputs of the REFERENCE'S OWN ``TimeWarpAugmenter`` and ``ExtremesFrequencyMasking``
                              (data/utils.py:41-190) on oracle log-mels under fixed torch seeds, with the replayed
                              (warp_p, warp_d) / (low, high) parameters; sub-sampled rows to stay small.
* ``mel_filters.npz``      -- ``transformers.audio_utils.mel_filter_bank`` (slaney / slaney) for 80 and 128 rows.

Inputs are regenerated from seeds by ``tests/signals.py``; only outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/src"

from oracle import logmel as O  # noqa: E402
from tests import signals as S  # noqa: E402


def install_stubs():
    """whisper / audiomentations stand-ins so that the reference modules import (same idea as the reference's tests)."""
    w = types.ModuleType("whisper")
    wa = types.ModuleType("whisper.audio")
    wa.CHUNK_LENGTH, wa.HOP_LENGTH, wa.N_FFT, wa.N_FRAMES, wa.N_SAMPLES = 30, 160, 400, 3000, 480000
    wa.log_mel_spectrogram = lambda audio, n_mels=80, padding=0, device=None: O.log_mel_spectrogram(audio, n_mels, padding)
    wt = types.ModuleType("whisper.tokenizer")
    wt.LANGUAGES, wt.TO_LANGUAGE_CODE, wt.Tokenizer = {"de": "german"}, {"german": "de"}, object
    w.audio, w.tokenizer = wa, wt
    sys.modules.update({"whisper": w, "whisper.audio": wa, "whisper.tokenizer": wt})

    class _Any(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, x, **k: x})

    sys.modules["audiomentations"] = _Any("audiomentations")
    sys.path.insert(0, REF_SRC)


def golden_pad_or_trim():
    from whisper_finetune.data.utils import pad_or_trim

    g = torch.Generator().manual_seed(11)
    cases = {}
    specs = [((16, 123), -1, 300), ((16, 300), -1, 150), ((128, 1), 1, 40), ((7, 50, 3), 1, 64),
             ((7, 50, 3), 0, 4), ((500,), 0, 600), ((16, 300), -1, 300), ((3, 9), 0, 5)]
    for k, (shape, axis, length) in enumerate(specs):
        x = torch.randn(*shape, generator=g) - 0.3
        yt = pad_or_trim(x, length, axis=axis)
        yn = pad_or_trim(x.numpy(), length, axis=axis)
        assert np.array_equal(yt.numpy(), yn)
        cases[f"in{k}"] = x.numpy()
        cases[f"out{k}"] = yt.numpy()
        cases[f"meta{k}"] = np.array([axis, length], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "pad_or_trim.npz"), n=len(specs), **cases)


def golden_calculate_mel():
    import torchaudio.transforms as T
    from whisper_finetune.data import data_loader as dl
    from oracle.specaug import torch_rng_mask_params

    out = {}
    cases = [  # (kind, n_samples, n_mels, partial_start_s, time_param, freq_param, seed)
        ("white", 480000, 80, None, 100, 43, 3),
        ("white", 200000, 128, 7.31, 100, 27, 4),
        ("hdr", 480000, 80, 0.02, 100, 43, 5),
        ("int16", 333333, 128, None, 100, 27, 6),
        ("chirp", 480000, 128, 29.99, 0, 0, 7),
    ]
    for k, (kind, n, n_mels, start, tp, fp, seed) in enumerate(cases):
        x = S.make(kind, n=n, seed=seed)
        if x.dtype == torch.int16:
            x = x.float() / 32768.0
        audio = np.pad(x.numpy(), (0, 480000 - n), "constant")  # data_loader.py:346
        ds = dl.AudioDataset.__new__(dl.AudioDataset)
        ds.aud_augment, ds.n_mels, ds.device = None, n_mels, None
        ds.num_frames_per_second = dl.N_FRAMES / dl.CHUNK_LENGTH
        ds.spec_augment, ds.spec_augment_p = tp > 0, 1.0
        ds.time_warping = lambda mel: mel  # time-warp is a "next" row
        ds.time_masking = T.TimeMasking(tp) if tp > 0 else None
        ds.freq_masking = T.FrequencyMasking(fp) if tp > 0 else None
        ds.extreme_freq_masking = None
        torch.manual_seed(seed)
        mel = ds._calculate_mel(audio, start, no_timestamps=start is not None)
        torch.manual_seed(seed)
        params = torch_rng_mask_params(n_mels, 3000, tp, fp) if tp > 0 else (0, 0, 0, 0)
        assert mel.shape == (n_mels, 3000)
        out[f"sub{k}"] = mel[:, ::16].numpy().copy()
        out[f"sum{k}"] = np.array([mel.double().sum().item(), mel.double().abs().sum().item(),
                                   float((mel == 0).sum())])
        out[f"mask{k}"] = np.array(params, dtype=np.int32)
        out[f"meta{k}"] = np.array([n, n_mels, -1 if start is None else int(start * 100), tp, fp, seed], dtype=np.int64)
        out[f"kind{k}"] = np.array(kind)
    np.savez_compressed(os.path.join(HERE, "calculate_mel.npz"), n=len(cases), **out)


def golden_logmel_hf():
    from transformers import WhisperFeatureExtractor

    out = {}
    cases = [("white", 16000, 80, 21), ("white", 24000, 128, 22), ("hdr", 32000, 128, 23), ("int16", 16000, 80, 24),
             ("impulse", 16000, 128, 25), ("zeros", 8000, 80, 26)]
    for k, (kind, n, n_mels, seed) in enumerate(cases):
        x = S.make(kind, n=n, seed=seed)
        if x.dtype == torch.int16:
            x = x.float() / 32768.0
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        feats = fe._np_extract_fbank_features(x.numpy()[None, :], "cpu")[0]
        out[f"out{k}"] = feats.astype(np.float32)
        out[f"meta{k}"] = np.array([n, n_mels, seed], dtype=np.int64)
        out[f"kind{k}"] = np.array(kind)
    np.savez_compressed(os.path.join(HERE, "logmel_hf.npz"), n=len(cases), **out)


def golden_logmel_hf_torch():
    """Full 30-s clips through ``WhisperFeatureExtractor._torch_extract_fbank_features`` -- the float32 ``torch.stft`` path
    of the independent ``transformers`` implementation, which agrees with the oracle to one float32 ulp.
    Stored: every 24th frame, the first and last 32 frames, and float64 per-frame column sums (so EVERY frame is pinned)."""
    from transformers import WhisperFeatureExtractor

    out, k = {}, 0
    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        for kind in ("white", "hdr", "int16", "zeros", "impulse", "chirp"):
            x = S.make(kind)
            if x.dtype == torch.int16:
                x = x.float() / 32768.0
            feats = np.asarray(fe._torch_extract_fbank_features(x.numpy()[None, :], "cpu"))[0].astype(np.float32)
            assert feats.shape == (n_mels, 3000)
            out[f"sub{k}"] = feats[:, ::24].copy()
            out[f"head{k}"] = feats[:, :32].copy()
            out[f"tail{k}"] = feats[:, -32:].copy()
            out[f"colsum{k}"] = feats.astype(np.float64).sum(axis=0)
            out[f"meta{k}"] = np.array([n_mels], dtype=np.int64)
            out[f"kind{k}"] = np.array(kind)
            k += 1
    np.savez_compressed(os.path.join(HERE, "logmel_hf_torch.npz"), n=k, **out)


def golden_timewarp():
    from whisper_finetune.data.utils import ExtremesFrequencyMasking, TimeWarpAugmenter

    out = {}
    cases = [("white", 128, 80, 3), ("hdr", 80, 80, 4), ("chirp", 128, 50, 5), ("int16", 80, 20, 6)]
    for k, (kind, n_mels, W, seed) in enumerate(cases):
        x = S.make(kind, seed=seed)
        if x.dtype == torch.int16:
            x = x.float() / 32768.0
        mel = O.log_mel_spectrogram(x, n_mels)
        torch.manual_seed(seed)
        warped = TimeWarpAugmenter(W=W)(mel)
        torch.manual_seed(seed)
        wp = int(torch.randint(W, 3000 - W, (1,)))
        wd = int(torch.randint(-W, W, (1,)))
        torch.manual_seed(seed + 100)
        ext = ExtremesFrequencyMasking(low_freq_range=10, high_freq_range=15)(mel.clone())
        torch.manual_seed(seed + 100)
        r = torch.rand(1).item()
        out[f"warp{k}"] = warped[::8].numpy().copy()
        out[f"ext_zero_rows{k}"] = (ext == 0).all(dim=1).numpy()
        out[f"meta{k}"] = np.array([n_mels, W, seed, wp, wd, int(round(r * 10)), int(round(r * 15))], dtype=np.int64)
        out[f"kind{k}"] = np.array(kind)
    np.savez_compressed(os.path.join(HERE, "timewarp.npz"), n=len(cases), **out)


def golden_mel_filters():
    from transformers.audio_utils import mel_filter_bank

    np.savez_compressed(os.path.join(HERE, "mel_filters.npz"),
                        **{f"hf{n}": mel_filter_bank(201, n, 0.0, 8000.0, 16000, "slaney", "slaney").T for n in (80, 128)})


if __name__ == "__main__":
    install_stubs()
    golden_pad_or_trim()
    golden_calculate_mel()
    golden_logmel_hf()
    golden_logmel_hf_torch()
    golden_timewarp()
    golden_mel_filters()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
