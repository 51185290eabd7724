"""Runs the reference's OWN tests/test_data_loader.py with this package bound under the names the reference imports
(``whisper.audio.log_mel_spectrogram``, ``data_loader.pad_or_trim``).  Only possible where /root/reference exists (the
build container); skipped on the GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

_DRIVER = r"""
import sys, types
root, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [root, ref + "/src"]
import whisper_finetune_b200 as wft

w = types.ModuleType("whisper"); wa = types.ModuleType("whisper.audio"); wt = types.ModuleType("whisper.tokenizer")
wa.CHUNK_LENGTH, wa.HOP_LENGTH, wa.N_FFT, wa.N_FRAMES, wa.N_SAMPLES = (wft.CHUNK_LENGTH, wft.HOP_LENGTH, wft.N_FFT,
                                                                        wft.N_FRAMES, wft.N_SAMPLES)
wa.log_mel_spectrogram = wft.log_mel_spectrogram          # the drop-in, under the reference's import name
wt.LANGUAGES, wt.TO_LANGUAGE_CODE, wt.Tokenizer = {"de": "german"}, {"german": "de"}, object
w.audio, w.tokenizer = wa, wt
sys.modules.update({"whisper": w, "whisper.audio": wa, "whisper.tokenizer": wt})

class _Any(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"): raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, x, **k: x})
sys.modules["audiomentations"] = _Any("audiomentations")

from whisper_finetune.data import data_loader as dl
assert dl.log_mel_spectrogram is wft.log_mel_spectrogram
wft.install(dl)
assert dl.pad_or_trim is wft.pad_or_trim
import pytest
sys.exit(pytest.main(["-q", "-p", "no:cacheprovider", ref + "/tests/test_data_loader.py"]))
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_data_loader_tests_pass_with_drop_in(tmp_path):
    script = tmp_path / "drive.py"
    script.write_text(_DRIVER)
    res = subprocess.run([sys.executable, str(script), ROOT, REF], capture_output=True, text=True, timeout=600,
                         cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "10 passed" in res.stdout
