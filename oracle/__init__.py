"""CPU oracle for the Whisper audio front end -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it, and only as the checker or as the timed *baseline*; the product path
(``whisper-finetune_b200``) never imports it and has no CPU fallback.

What it restates (reference paths are relative to ``/root/reference``):

* ``logmel.py``      -- ``whisper.audio.log_mel_spectrogram`` (third-party ``openai-whisper>=20240930``,
                        pinned only by a floor in ``pyproject.toml:12`` and NOT vendored in the reference;
                        call site ``src/whisper_finetune/data/data_loader.py:13,278``).
                        **parity unpinned** at this boundary: the reference's own tests stub the function
                        (``tests/test_data_loader.py:32``) and ship no golden vector.  The restatement follows
                        the published upstream algorithm (torch.stft recipe) and is cross-checked against the
                        independent ``transformers`` Whisper feature extractor available in this image.
* ``mel_filters.py`` -- ``librosa.filters.mel(sr=16000, n_fft=400, n_mels)`` (the content of upstream's
                        ``assets/mel_filters.npz``), cross-checked against ``transformers.audio_utils``.
* ``pad_or_trim.py`` -- ``src/whisper_finetune/data/utils.py:380-404`` (reference-owned; PINNED by golden
                        vectors produced by importing the reference function, ``tests/golden/make_golden.py``).
* ``specaug.py``     -- ``torchaudio.functional.mask_along_axis`` as used at ``data_loader.py:115-116,286-287``
                        (the real torchaudio is installed here and is the pin) + the counter-based draw.
* ``pipeline.py``    -- ``AudioDataset._calculate_mel`` ordering, ``data_loader.py:273-292,344-346``.
* ``sampler.py``     -- ``torch.utils.data.DistributedSampler`` index partition used at
                        ``src/whisper_finetune/scripts/finetune.py:619-629``.
"""
