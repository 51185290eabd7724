"""ORACLE (test infrastructure, never shipped): the reference's min-value ``pad_or_trim``.

Behavioural restatement of ``/root/reference/src/whisper_finetune/data/utils.py:380-404`` (NOT whisper's
zero-padding variant): along ``axis`` keep the first ``length`` entries, or right-pad up to ``length`` with
the minimum of the *whole input array*.  Tensor in -> tensor out, ndarray in -> ndarray out, equal length ->
the input object itself.  An empty input makes the min reduction raise, exactly like the reference.

PINNED: ``tests/golden/pad_or_trim_*.npz`` were produced by importing the reference function
(``tests/golden/make_golden.py``) and this restatement is checked against them bit for bit.
"""
import numpy as np
import torch

N_SAMPLES = 480000


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    n = array.shape[axis]
    if n == length:
        return array
    is_t = torch.is_tensor(array)
    if n > length:
        if is_t:
            return array.narrow(axis, 0, length).clone()
        return np.take(array, np.arange(length), axis=axis)
    # n < length : fill value is the global minimum (utils.py:393 / :401)
    if is_t:
        fill = torch.min(array).item()
        shape = list(array.shape)
        shape[axis] = length - n
        tail = torch.full(shape, fill, dtype=array.dtype, device=array.device)
        return torch.cat([array, tail], dim=axis)
    fill = np.min(array)
    shape = list(array.shape)
    shape[axis] = length - n
    tail = np.full(shape, fill, dtype=array.dtype)
    return np.concatenate([array, tail], axis=axis)
