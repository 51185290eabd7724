"""CPU check of the kernel's ALGORITHM (not of the CUDA code): the numpy float32 model in tests/kernel_model.py follows the
same decomposition as csrc/frontend_kernel.cuh and must meet the north-star tolerance against the oracle."""
import numpy as np
import pytest
import torch

from oracle import logmel as O
from tests import kernel_model as KM
from tests import signals as S


def test_dft20_prime_factor_maps():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((5, 20)) + 1j * rng.standard_normal((5, 20))).astype(np.complex64)
    assert np.abs(KM.dft20(x) - np.fft.fft(x.astype(np.complex128))).max() < 5e-6


@pytest.mark.parametrize("kind", ["white", "hdr", "int16", "zeros", "impulse"])
def test_algorithm_meets_tolerance(kind):
    x = S.make(kind)
    xf = x.float() / 32768.0 if x.dtype == torch.int16 else x
    got = torch.from_numpy(KM.model_logmel(xf.numpy(), 128))
    ref = O.log_mel_spectrogram(x, 128)
    ma, rl = S.metrics(got, ref)
    assert ma <= S.MAX_ABS and rl <= S.REL_L2, (kind, ma, rl)
