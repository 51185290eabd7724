"""CPU tests: the oracle against every pin we have (reference-generated goldens, torchaudio, transformers,
torch's DistributedSampler, Philox known answers).  No GPU, no /root/reference at run time."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel as O
from oracle import mel_filters as OM
from oracle import pipeline as OP
from oracle import sampler as OSamp
from oracle import specaug as OS
from oracle.pad_or_trim import pad_or_trim as o_pad_or_trim
from tests import signals as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _npz(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def test_constants_match_reference_test_stub():
    # /root/reference/tests/test_data_loader.py:27-31 pins exactly these
    assert (O.CHUNK_LENGTH, O.HOP_LENGTH, O.N_FFT, O.N_FRAMES, O.N_SAMPLES) == (30, 160, 400, 3000, 480000)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_mel_bank_known_answers(n_mels):
    f = OM.mel_filters(n_mels)
    assert f.shape == (n_mels, 201) and f.dtype == np.float32
    assert int((f != 0).sum()) == {80: 391, 128: 394}[n_mels]
    assert not f[:, 0].any() and not f[:, 200].any()
    assert abs(float(f[0, 1]) - {80: 0.024862594902515, 128: 0.012373986653984}[n_mels]) < 1e-12
    assert ((f != 0).sum(0) <= 2).all()
    hf = _npz("mel_filters.npz")[f"hf{n_mels}"]
    assert np.abs(f - hf.astype(np.float32)).max() <= 4e-9  # <= 1 float32 ulp of the largest weight


def test_pad_or_trim_against_reference_goldens():
    z = _npz("pad_or_trim.npz")
    for k in range(int(z["n"])):
        x, want = z[f"in{k}"], z[f"out{k}"]
        axis, length = (int(v) for v in z[f"meta{k}"])
        got_t = o_pad_or_trim(torch.from_numpy(x), length, axis=axis)
        got_n = o_pad_or_trim(x, length, axis=axis)
        assert torch.is_tensor(got_t) and isinstance(got_n, np.ndarray)
        assert np.array_equal(got_t.numpy(), want) and np.array_equal(got_n, want)
    x = torch.randn(4, 7)
    assert o_pad_or_trim(x, 7) is x  # no-op returns the input object, like the reference
    with pytest.raises(RuntimeError):
        o_pad_or_trim(torch.zeros(80, 0), 3000)
    with pytest.raises(ValueError):
        o_pad_or_trim(np.zeros((80, 0), dtype=np.float32), 3000)


def test_calculate_mel_against_reference_goldens():
    """oracle/pipeline.py reproduces AudioDataset._calculate_mel run in the build container (ordering, cut, min pad,
    masks): sub-sampled frames and whole-tensor sums match bit for bit."""
    z = _npz("calculate_mel.npz")
    for k in range(int(z["n"])):
        n, n_mels, nv, tp, fp, seed = (int(v) for v in z[f"meta{k}"])
        x = S.make(str(z[f"kind{k}"]), n=n, seed=seed)
        mask = z[f"mask{k}"] if tp > 0 else None
        got = OP.calculate_mel(x, n_mels, None if nv < 0 else nv, mask)
        assert got.shape == (n_mels, 3000)
        assert np.array_equal(got[:, ::16].numpy(), z[f"sub{k}"])
        sums = z[f"sum{k}"]
        assert got.double().sum().item() == sums[0] and got.double().abs().sum().item() == sums[1]
        assert float((got == 0).sum()) == sums[2]


def test_logmel_against_transformers_goldens():
    z = _npz("logmel_hf.npz")
    for k in range(int(z["n"])):
        n, n_mels, seed = (int(v) for v in z[f"meta{k}"])
        x = S.make(str(z[f"kind{k}"]), n=n, seed=seed)
        got = O.log_mel_spectrogram(x, n_mels).numpy()
        want = z[f"out{k}"]
        assert got.shape == want.shape == (n_mels, n // 160)
        assert np.abs(got - want).max() <= 2e-4, (k, np.abs(got - want).max())


def test_logmel_against_transformers_torch_extractor_full_clips():
    """The tight independent anchor: full 30-s clips, all signal kinds, 80 / 128 mel, against the float32
    torch.stft path of the ``transformers`` Whisper feature extractor -- one float32 ulp."""
    z = _npz("logmel_hf_torch.npz")
    assert int(z["n"]) == 12
    for k in range(int(z["n"])):
        n_mels = int(z[f"meta{k}"][0])
        got = O.log_mel_spectrogram(S.make(str(z[f"kind{k}"])), n_mels).numpy()
        assert got.shape == (n_mels, 3000)
        for name, view in (("sub", got[:, ::24]), ("head", got[:, :32]), ("tail", got[:, -32:])):
            d = np.abs(view - z[f"{name}{k}"]).max()
            assert d <= 2e-7, (k, str(z[f"kind{k}"]), name, d)
        # every frame: float64 column sums agree to n_mels x 2e-7
        assert np.abs(got.astype(np.float64).sum(axis=0) - z[f"colsum{k}"]).max() <= n_mels * 2e-7


def test_logmel_known_answers():
    z = O.log_mel_spectrogram(torch.zeros(480000), 80)
    assert z.shape == (80, 3000) and torch.all(z == -1.5)
    x = S.make("hdr")
    m = O.log_mel_spectrogram(x, 128)
    assert (m.max() - m.min()).item() <= 2.0 + 1e-6
    # float32 oracle vs float64 truth: the oracle's own rounding noise stays inside the parity budget
    ma, rl = S.metrics(m, O.log_mel_spectrogram(x, 128, dtype=torch.float64))
    assert ma <= 1e-3 and rl <= 1e-5
    # explicit reflect-pad / frame / DFT restatement == torch.stft (float64)
    xs = S.make("white", n=4000, seed=9).double()
    p = torch.cat([xs[1:201].flip(0), xs, xs[-201:-1].flip(0)])
    frames = p.unfold(0, 400, 160)[:-1]
    w = torch.hann_window(400, dtype=torch.float64)
    k = torch.arange(201, dtype=torch.float64)[:, None] * torch.arange(400, dtype=torch.float64)[None, :]
    dft = torch.exp(-2j * np.pi * k / 400)
    spec = (frames * w).to(torch.complex128) @ dft.T
    ref = torch.stft(xs, 400, 160, window=w, return_complex=True)[..., :-1]
    assert torch.allclose(spec.T, ref, atol=1e-9, rtol=0)


def test_batch_is_clip_by_clip():
    xs = torch.stack([S.make("white", n=16000, seed=1), 1e-3 * S.make("white", n=16000, seed=2)])
    b = O.log_mel_batch(xs, 80)
    for i in range(2):
        assert torch.equal(b[i], O.log_mel_spectrogram(xs[i], 80))
    lens = [16000, 5000]
    b2 = O.log_mel_batch(xs, 80, lengths=lens)
    y = xs[1].clone()
    y[5000:] = 0
    assert torch.equal(b2[1], O.log_mel_spectrogram(y, 80))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert OS.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert OS.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert OS.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_mask_restatement_replays_torchaudio_bit_exact():
    import torchaudio.transforms as T

    mel = torch.randn(128, 3000)
    for seed in range(200):
        torch.manual_seed(seed)
        ref = T.FrequencyMasking(27)(T.TimeMasking(100)(mel))
        torch.manual_seed(seed)
        t0, t1, f0, f1 = OS.torch_rng_mask_params(128, 3000, 100, 27)
        assert torch.equal(OS.apply_masks(mel, t0, t1, f0, f1), ref)
        assert 0 <= t1 - t0 < 100 and 0 <= f1 - f0 < 27 and t1 <= 3000 and f1 <= 128


def test_counter_based_draw_properties():
    m = OS.draw_mask_params(42, 0, 512, 128, 3000, 100, 27, 1.0)
    assert m.shape == (512, 4) and m.dtype == np.int32
    assert ((m[:, 1] - m[:, 0]) < 100).all() and ((m[:, 3] - m[:, 2]) < 27).all() and (m >= 0).all()
    assert (m[:, 1] <= 3000).all() and (m[:, 3] <= 128).all()
    # shard independence: rank r of 4 draws exactly its rows of the global table
    for r in range(4):
        part = np.concatenate([OS.draw_mask_params(42, i, 1, 128, 3000, 100, 27, 1.0) for i in range(r, 512, 4)])
        assert np.array_equal(part, m[r::4])
    assert not OS.draw_mask_params(42, 0, 64, 128, 3000, 100, 27, 0.0).any()
    half = OS.draw_mask_params(42, 0, 2000, 128, 3000, 100, 27, 0.5)
    frac = (half.any(axis=1)).mean()
    assert 0.4 < frac < 0.6
    assert not OS.draw_mask_params(1, 0, 8, 128, 3000, 0, 0, 1.0).any()


@pytest.mark.parametrize("n,world,drop_last,shuffle", [(100, 4, False, True), (101, 4, True, True), (7, 8, False, True),
                                                         (64, 2, True, False), (1000, 8, False, True), (9, 4, True, True)])
def test_sampler_restatement_matches_torch(n, world, drop_last, shuffle):
    from torch.utils.data import DistributedSampler

    for epoch in (0, 3):
        for rank in range(world):
            ds = DistributedSampler(range(n), num_replicas=world, rank=rank, shuffle=shuffle, seed=42, drop_last=drop_last)
            ds.set_epoch(epoch)
            assert list(iter(ds)) == OSamp.rank_indices(n, world, rank, epoch, 42, shuffle, drop_last)


def test_timewarp_and_extremes_against_reference_goldens():
    """oracle/timewarp.py vs the reference's TimeWarpAugmenter / ExtremesFrequencyMasking outputs.  The reference evaluates
    its spline in float32 (pow + matmul), which moves source coordinates by ~1e-4 frames: tolerance max-abs 1e-3."""
    from oracle import timewarp as OT

    z = _npz("timewarp.npz")
    for k in range(int(z["n"])):
        n_mels, W, seed, wp, wd, low, high = (int(v) for v in z[f"meta{k}"])
        x = S.make(str(z[f"kind{k}"]), seed=seed)
        mel = O.log_mel_spectrogram(x, n_mels)
        got = OT.time_warp(mel, wp, wd)
        assert np.abs(got[::8].numpy() - z[f"warp{k}"]).max() <= 1e-3
        assert W <= wp < 3000 - W and -W <= wd < W
        ext = OT.extremes_mask(mel, low, high)
        assert np.array_equal((ext == 0).all(dim=1).numpy(), z[f"ext_zero_rows{k}"])
    ident = OT.time_warp(mel, 1500, 0)  # warp_d = 0: the spline is the identity map
    assert (ident - mel).abs().max() <= 2e-4


def test_deep_spec_augment_oracle_matches_torchaudio():
    """oracle.deep_specaug == the reference hook body (permute -> torchaudio masks -> permute), same seed."""
    import torchaudio.transforms as T

    from oracle.deep_specaug import deep_spec_augment

    x = torch.randn(3, 150, 64)
    for seed, (tp, fp) in enumerate([(100, 27), (40, 10), (1, 1), (0, 43), (400, 27), (400, 200), (149, 63)]):
        torch.manual_seed(seed)
        ref = T.FrequencyMasking(freq_mask_param=fp)(T.TimeMasking(time_mask_param=tp)(x.permute(0, 2, 1))).permute(0, 2, 1)
        torch.manual_seed(seed)
        got, t, f = deep_spec_augment(x, tp, fp)
        assert torch.equal(got, ref.contiguous()), (seed, tp, fp, t, f)
        assert t[1] - t[0] <= max(tp - 1, 0) and f[1] - f[0] <= max(fp - 1, 0)
