"""Import alias: ``import whisper_finetune_b200`` -> the package in ``whisper-finetune_b200/`` (a hyphen is not
importable with a plain ``import`` statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("whisper-finetune_b200")
sys.modules[__name__] = _pkg
