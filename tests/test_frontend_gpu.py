"""GPU parity: the CUDA front end (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): max-abs <= 1e-3 and rel-L2 <= 1e-5 in float32 for the features; mask cells
and pad_or_trim bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import logmel as O
from oracle import pipeline as OP
from oracle import specaug as OS
from tests import signals as S

pytestmark = pytest.mark.gpu


def _check(got, ref, what=""):
    ma, rl = S.metrics(got, ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert ma <= S.MAX_ABS and rl <= S.REL_L2, f"{what}: max-abs {ma:.3e} rel-L2 {rl:.3e}"
    return ma, rl


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("kind", ["white", "hdr", "int16", "zeros", "impulse", "chirp"])
def test_logmel_full_clip_matches_oracle(wft, cuda, kind, n_mels):
    x = S.make(kind)
    got = wft.log_mel_spectrogram(x.to(cuda), n_mels=n_mels).cpu()
    ref = O.log_mel_spectrogram(x, n_mels)
    assert got.shape == (n_mels, 3000)
    _check(got, ref, f"{kind}/{n_mels}")


def test_zeros_known_answer(wft, cuda):
    got = wft.log_mel_spectrogram(torch.zeros(480000, device=cuda), n_mels=80)
    assert torch.allclose(got, torch.full_like(got, -1.5), atol=2e-6, rtol=0)


def test_dynamic_range_property(wft, cuda):
    x = S.make("hdr")
    got = wft.log_mel_spectrogram(x.to(cuda), n_mels=128)
    assert (got.max() - got.min()).item() <= 2.0 + 1e-6


@pytest.mark.parametrize("n", [201, 400, 1599, 1600, 1601, 5120, 16000, 48001, 479999])
def test_short_and_ragged_lengths(wft, cuda, n):
    x = S.make("white", n=n, seed=n)
    got = wft.log_mel_spectrogram(x.to(cuda), n_mels=80).cpu()
    ref = O.log_mel_spectrogram(x, 80)
    assert got.shape == (80, n // 160)
    _check(got, ref, f"n={n}")


@pytest.mark.parametrize("padding", [1, 160, 4000])
def test_padding_argument(wft, cuda, padding):
    x = S.make("white", n=32000, seed=5)
    got = wft.log_mel_spectrogram(x.to(cuda), n_mels=128, padding=padding).cpu()
    ref = O.log_mel_spectrogram(x, 128, padding=padding)
    _check(got, ref, f"padding={padding}")


def test_numpy_and_int16_inputs(wft, cuda):
    xi = S.make("int16", n=64000)
    got = wft.log_mel_spectrogram(xi.numpy(), n_mels=80, device="cuda").cpu()
    ref = O.log_mel_spectrogram(xi, 80)
    _check(got, ref, "int16 ndarray")
    xf = xi.float() / 32768.0
    got2 = wft.log_mel_spectrogram(xf.numpy(), n_mels=80, device="cuda").cpu()
    assert torch.equal(got, got2), "int16 and the equivalent float32 PCM must give identical features"


def test_batch_is_per_clip(wft, cuda):
    """A batch must equal clip-by-clip calls (per-clip max, F6), not whisper's global max over the batch."""
    xs = torch.stack([S.make("white", n=48000, seed=1), 1e-3 * S.make("white", n=48000, seed=2),
                      S.make("zeros", n=48000), S.make("chirp", n=48000)])
    got = wft.log_mel_spectrogram(xs.to(cuda), n_mels=128)
    for b in range(xs.shape[0]):
        one = wft.log_mel_spectrogram(xs[b].to(cuda), n_mels=128)
        assert torch.equal(got[b], one)
    _check(got.cpu(), O.log_mel_batch(xs, 128), "batch")


def test_errors(wft, cuda):
    with pytest.raises(ValueError):
        wft.log_mel_spectrogram(torch.zeros(16000, device=cuda), n_mels=64)
    with pytest.raises(NotImplementedError):
        wft.log_mel_spectrogram("clip.wav")
    with pytest.raises(ValueError):
        wft.log_mel_spectrogram(torch.zeros(100, device=cuda), n_mels=80)


@pytest.mark.parametrize("dtype", ["f32", "i16"])
def test_front_end_variable_length_cut_and_masks(wft, cuda, dtype):
    """config 3 shape: ragged clips zero-padded in-kernel, 25% partial-segment cuts with min-pad, masks."""
    rng = np.random.default_rng(42)
    B = 12
    lengths = rng.integers(16000, 480001, size=B).astype(np.int32)
    lengths[0] = 480000
    lengths[1] = 16000
    pcm = torch.zeros(B, 480000)
    for b in range(B):
        pcm[b, : lengths[b]] = S.make("white", n=int(lengths[b]), seed=100 + b)
    if dtype == "i16":
        pcm = torch.round(pcm * 32767).to(torch.int16)
    n_valid = np.full(B, -1, dtype=np.int32)
    for b in range(0, B, 4):
        n_valid[b] = int(rng.uniform(0.02, 30.0) * 100)
    n_valid[4] = 1
    masks = OS.draw_mask_params(42, 1000, B, 128, 3000, 100, 27, 1.0)
    fe = wft.FrontEnd(n_mels=128, spec_augment=True,
                      spec_augment_params={"time_mask_param": 100, "freq_mask_param": 27, "p": 1.0}, seed=42)
    got = fe(pcm.to(cuda), lengths=lengths, n_valid_frames=n_valid, clip_offset=1000).cpu()
    ref = OP.front_end_batch(pcm, 128, lengths=lengths, n_valid_frames=n_valid, masks=masks)
    assert got.shape == (B, 128, 3000)
    for b in range(B):
        keep = 3000 if n_valid[b] < 0 else int(n_valid[b])
        _check(got[b, :, :keep], ref[b, :, :keep], f"front end clip {b} (kept frames)")
        assert (got[b] - ref[b]).abs().max() <= S.MAX_ABS
    for b in range(B):
        t0, t1, f0, f1 = masks[b]
        m = torch.zeros(128, 3000, dtype=torch.bool)
        m[:, t0:t1] = True
        m[f0:f1, :] = True
        assert torch.equal(got[b][m], torch.zeros(int(m.sum()))), "masked cells must be exactly 0.0"
        if n_valid[b] >= 0:  # min-value pad: constant, equal to the minimum of the kept part (bit-exact property)
            keep = int(n_valid[b])
            unmasked_pad = got[b][:, keep:][~m[:, keep:]]
            kept_min = wft.log_mel_spectrogram(pcm[b, : lengths[b]].to(cuda), n_mels=128,
                                               padding=480000 - int(lengths[b]))[:, :keep].min().item()
            assert unmasked_pad.numel() == 0 or (unmasked_pad == kept_min).all()


def test_device_draw_matches_oracle_draw(wft, cuda):
    for (seed, off, B, nm, T, F, p) in [(42, 0, 257, 128, 100, 27, 1.0), (7, 2**33 + 5, 64, 80, 100, 43, 0.5),
                                        (0, 0, 5, 128, 0, 43, 1.0), (1, 9, 33, 80, 100, 43, 0.0)]:
        got = wft.draw_mask_params(seed, off, B, nm, 3000, T, F, p).cpu().numpy()
        ref = OS.draw_mask_params(seed, off, B, nm, 3000, T, F, p)
        assert np.array_equal(got, ref)


def test_mask_callables_replay_torchaudio(wft, cuda):
    """Same torch seed -> same cells masked as the real torchaudio transforms (data_loader.py:286-287)."""
    import torchaudio.transforms as T

    mel = torch.randn(128, 3000)
    for seed in range(20):
        torch.manual_seed(seed)
        ref = T.FrequencyMasking(27)(T.TimeMasking(100)(mel))
        torch.manual_seed(seed)
        got = wft.FrequencyMasking(27)(wft.TimeMasking(100)(mel.to(cuda)))
        assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("shape,axis,length", [((80, 1234), -1, 3000), ((80, 3000), -1, 1500), ((128, 1), 1, 3000),
                                               ((7, 50, 3), 1, 64), ((7, 50, 3), 0, 4), ((5000,), 0, 480000)])
def test_pad_or_trim_bit_exact(wft, cuda, shape, axis, length):
    from oracle.pad_or_trim import pad_or_trim as ref_pad

    x = torch.randn(*shape) - 0.3
    got = wft.pad_or_trim(x.to(cuda), length, axis=axis)
    ref = ref_pad(x, length, axis=axis)
    assert got.is_cuda and torch.equal(got.cpu(), ref)
    got_np = wft.pad_or_trim(x.numpy(), length, axis=axis)
    assert isinstance(got_np, np.ndarray) and np.array_equal(got_np, ref.numpy())
    same = wft.pad_or_trim(x.to(cuda), x.shape[axis], axis=axis)
    assert same.data_ptr() == x.to(cuda).data_ptr() or torch.equal(same.cpu(), x)


def test_pad_or_trim_nonfinite_minimum_like_torch_min(wft, cuda):
    """torch.min propagates NaN and an all-(+inf) input has minimum +inf (data/utils.py:380-404 pads with array.min())."""
    x = torch.randn(4, 10)
    x[1, 3] = float("nan")
    got = wft.pad_or_trim(x.to(cuda), 20, axis=1).cpu()
    assert torch.isnan(got[:, 10:]).all()
    assert torch.equal(got[:, :10].nan_to_num(7.0), x.nan_to_num(7.0))
    y = torch.full((2, 5), float("inf"))
    got = wft.pad_or_trim(y.to(cuda), 8, axis=1).cpu()
    assert torch.isinf(got).all() and (got > 0).all()


def test_pad_or_trim_empty_raises(wft, cuda):
    with pytest.raises(RuntimeError):
        wft.pad_or_trim(torch.zeros(80, 0, device=cuda), 3000)


def test_large_batch_roundtrip_properties(wft, cuda):
    """BASELINE size (B=64, 128 mel): size-independent properties instead of a full oracle run."""
    g = torch.Generator().manual_seed(0)
    pcm = (0.1 * torch.randn(64, 480000, generator=g)).clamp(-1, 1)
    pcm[5] *= 1e-3
    pcm[9, 100000:] = 0
    d = pcm.to(cuda)
    a = wft.log_mel_spectrogram(d, n_mels=128)
    b = wft.log_mel_spectrogram(d, n_mels=128)
    assert torch.equal(a, b), "deterministic"
    assert a.shape == (64, 128, 3000) and torch.isfinite(a).all()
    rng = a.amax(dim=(1, 2)) - a.amin(dim=(1, 2))
    assert (rng <= 2.0 + 1e-6).all()
    perm = torch.randperm(64, generator=g)
    c = wft.log_mel_spectrogram(d[perm.to(cuda)], n_mels=128)
    assert torch.equal(c, a[perm.to(cuda)]), "clips are independent: permuting the batch permutes the output"
    ref = O.log_mel_batch(pcm, 128)          # every clip of the batch against the oracle (not a spot check)
    host = a.cpu()
    for bidx in range(64):
        _check(host[bidx], ref[bidx], f"clip {bidx}")


def _config3_batch(B, seed):
    """config 3 of BASELINE.json: ragged U(1 s, 30 s) clips, 25 % partial-segment cuts, SpecAugment masks."""
    rng = np.random.default_rng(seed)
    lengths = rng.integers(16000, 480001, size=B).astype(np.int32)
    lengths[0], lengths[1] = 480000, 16000
    g = torch.Generator().manual_seed(seed)
    pcm = (0.1 * torch.randn(B, 480000, generator=g)).clamp(-1, 1)
    pcm *= torch.from_numpy(10.0 ** rng.uniform(-3, 0, size=(B, 1))).float()      # clips of very different loudness
    pcm[torch.arange(480000)[None, :] >= torch.from_numpy(lengths)[:, None]] = 0.0
    n_valid = np.full(B, -1, dtype=np.int32)
    cut = rng.random(B) < 0.25
    n_valid[cut] = (rng.uniform(0.02, 30.0, size=int(cut.sum())) * 100).astype(np.int32)
    return pcm, lengths, n_valid


def test_config3_batch_256_matches_oracle(wft, cuda):
    """BASELINE config 3 at its full size (B = 256, 128 mel, ragged + cuts + masks), every clip against the oracle."""
    B = 256
    pcm, lengths, n_valid = _config3_batch(B, 2024)
    masks = OS.draw_mask_params(11, 5000, B, 128, 3000, 100, 27, 1.0)
    fe = wft.FrontEnd(n_mels=128, spec_augment=True,
                      spec_augment_params={"time_mask_param": 100, "freq_mask_param": 27, "p": 1.0}, seed=11)
    got = fe(pcm.to(cuda), lengths=lengths, n_valid_frames=n_valid, clip_offset=5000).cpu()
    ref = OP.front_end_batch(pcm, 128, lengths=lengths, n_valid_frames=n_valid, masks=masks)
    for b in range(B):
        keep = 3000 if n_valid[b] < 0 else int(n_valid[b])
        if keep:
            _check(got[b, :, :keep], ref[b, :, :keep], f"config 3 clip {b} (kept frames)")
        assert (got[b] - ref[b]).abs().max() <= S.MAX_ABS
        t0, t1, f0, f1 = masks[b]
        m = torch.zeros(128, 3000, dtype=torch.bool)
        m[:, t0:t1] = True
        m[f0:f1, :] = True
        assert torch.equal(got[b][m], torch.zeros(int(m.sum()))), "masked cells must be exactly 0.0"
        assert int((got[b][~m] == 0).sum()) <= 2, "an unmasked feature is 0.0 only by coincidence"


@pytest.mark.parametrize("max_ctas", [1, 3, 40])
def test_result_does_not_depend_on_the_grid(wft, cuda, max_ctas):
    """The result must not depend on the persistent grids: with 1-40 CTAs both the front-end kernel (tile claims, TMA
    prefetch chain) and the fix-up kernel (grid-stride scan; floor binds on the hdr clip, min pad, silent tiles, masks)
    walk many more tiles per CTA than on the full grid."""
    lib = wft._lib.load()
    pcm = torch.stack([S.make("hdr"), S.make("white", seed=3), S.make("int16").float() / 32768.0, S.make("chirp")])
    lengths = np.asarray([480000, 300000, 480000, 123456], dtype=np.int32)
    pcm[torch.arange(480000)[None, :] >= torch.from_numpy(lengths)[:, None]] = 0.0
    n_valid = np.asarray([-1, 1000, 2999, -1], dtype=np.int32)
    masks = OS.draw_mask_params(5, 77, 4, 128, 3000, 100, 27, 1.0)
    fe = wft.FrontEnd(n_mels=128)
    full = fe(pcm.to(cuda), lengths=lengths, n_valid_frames=n_valid, mask_params=masks)
    try:
        lib.wft_debug_set_max_ctas(max_ctas)
        capped = fe(pcm.to(cuda), lengths=lengths, n_valid_frames=n_valid, mask_params=masks)
        torch.cuda.synchronize()
    finally:
        lib.wft_debug_set_max_ctas(0)
    assert torch.equal(capped, full)
    ref = OP.front_end_batch(pcm, 128, lengths=lengths, n_valid_frames=n_valid, masks=masks)
    for b in range(4):
        keep = 3000 if n_valid[b] < 0 else int(n_valid[b])
        _check(capped[b, :, :keep].cpu(), ref[b, :, :keep], f"capped grid clip {b}")
        assert (capped[b].cpu() - ref[b]).abs().max() <= S.MAX_ABS


def test_forty_minute_clip_floor_binds_almost_everywhere(wft, cuda):
    """One 40-minute clip = 15 000 tiles of ONE clip: the floor binds on almost all of them (2 s of loud tone, then faint
    noise 14 decades down), so the fix-up grid rewrites nearly the whole output."""
    n = 40 * 60 * 16000
    g = torch.Generator().manual_seed(40)
    x = 1e-7 * torch.randn(n, generator=g)
    t = torch.arange(32000, dtype=torch.float64) / 16000.0
    x[:32000] += (0.5 * torch.sin(2 * np.pi * 440.0 * t)).float()
    got = wft.log_mel_spectrogram(x.to(cuda), n_mels=128).cpu()
    ref = O.log_mel_spectrogram(x, 128)
    assert got.shape == (128, n // 160)
    _check(got, ref, "40-minute clip")
    floor = got.min().item()
    assert (got == floor).float().mean().item() > 0.9, "the floor must bind on the quiet part"
    assert abs((got.max().item() - floor) - 2.0) <= 1e-5


def test_overlapping_batches_equal_ordinary_launches(wft, cuda):
    """WFT_LAUNCH_OVERLAP (independent batches do not wait for the grids in front of them): 40 consecutive calls on rotating
    buffers -- more than two trips round the workspace ring -- must reproduce the ordinary launches bit for bit: full-length
    batches (nothing to fix up), a ragged batch with cuts and masks (the fix-up grid rewrites a third of the tiles while the
    next batch already runs), and a call that re-uses the previous call's output buffer (must NOT be overlapped)."""
    B = 12
    pcm, lengths, n_valid = _config3_batch(B, 7)
    masks = OS.draw_mask_params(3, 100, B, 128, 3000, 100, 27, 1.0)
    g = torch.Generator().manual_seed(11)
    sets = []
    for s in range(3):
        x = (0.1 * torch.randn(B, 480000, generator=g)).clamp(-1, 1)
        x[s] *= 1e-4                      # one quiet clip per batch
        x[s + 3, 200000:] = 0.0           # exact zeros without `lengths`: the floor binds, nothing is "silent"
        sets.append(x.to(cuda))
    ragged = pcm.to(cuda)
    # device-side metadata: no host-to-device copy sits between two calls, so the ragged batch overlaps its neighbours too
    lengths, n_valid = torch.from_numpy(lengths).to(cuda), torch.from_numpy(n_valid).to(cuda)
    masks = torch.as_tensor(np.asarray(masks), dtype=torch.int32).to(cuda)
    fe = wft.FrontEnd(n_mels=128)
    want = [fe(x) for x in sets]
    want_ragged = fe(ragged, lengths=lengths, n_valid_frames=n_valid, mask_params=masks)
    torch.cuda.synchronize()
    outs = [torch.empty_like(want[0]) for _ in range(4)]
    old = wft.set_overlap(True)
    try:
        got = []
        for i in range(40):
            k = i % 4
            if k == 3:
                fe(ragged, lengths=lengths, n_valid_frames=n_valid, mask_params=masks, out=outs[3])
            else:
                fe(sets[k], out=outs[k])
            if i % 7 == 6:   # same output buffer twice in a row: the second call must wait for the first
                fe(sets[(k + 1) % 3], out=outs[k])
                fe(sets[k] if k < 3 else sets[0], out=outs[k])
                if k == 3:
                    fe(ragged, lengths=lengths, n_valid_frames=n_valid, mask_params=masks, out=outs[3])
        torch.cuda.synchronize()
    finally:
        wft.set_overlap(old)
    for k in range(3):
        assert torch.equal(outs[k], want[k]), f"overlapped batch {k}"
    assert torch.equal(outs[3], want_ragged), "overlapped ragged batch"
    # and with the intervals drawn inside the call (draw grid -> front-end grid -> fix-up grid per batch)
    fe_aug = wft.FrontEnd(n_mels=128, spec_augment=True, spec_augment_params={"time_mask_param": 100, "freq_mask_param": 27, "p": 1.0},
                          seed=5)
    want_aug = [fe_aug(sets[k], clip_offset=64 * k).clone() for k in range(3)]
    old = wft.set_overlap(True)
    try:
        for i in range(36):
            fe_aug(sets[i % 3], clip_offset=64 * (i % 3), out=outs[i % 3])
        torch.cuda.synchronize()
    finally:
        wft.set_overlap(old)
    for k in range(3):
        assert torch.equal(outs[k], want_aug[k]), f"overlapped augmented batch {k}"


def test_production_batch_with_drawn_epilogue_equals_explicit_composition(wft, cuda):
    """FrontEnd with time_warp_w > 0 takes front-end grid -> fix-up grid -> ONE epilogue grid that draws the warp point and the
    mask intervals itself (wft_augment_drawn_f32).  It must equal, bit for bit, the explicit composition draw -> draw ->
    front end -> epilogue, also when consecutive batches overlap (scratch buffers alternate)."""
    B = 6
    g = torch.Generator().manual_seed(21)
    sets = [(0.1 * torch.randn(B, 480000, generator=g)).clamp(-1, 1).to(cuda) for _ in range(3)]
    lengths = torch.tensor([480000, 400000, 480000, 16000, 480000, 250000], dtype=torch.int32, device=cuda)
    for x in sets:
        x[torch.arange(480000, device=cuda)[None, :] >= lengths[:, None]] = 0.0
    params = {"time_mask_param": 100, "freq_mask_param": 27, "time_warp_w": 80, "p": 0.6}
    fe = wft.FrontEnd(n_mels=128, spec_augment=True, spec_augment_params=params, seed=17)
    want = []
    for k, x in enumerate(sets):
        masks = wft.draw_mask_params(17, 1000 * k, B, 128, 3000, 100, 27, 0.6, cuda)
        warps = wft.draw_warp_params(17, 1000 * k, B, 3000, 80, 0.6, cuda)
        plain = wft.frontend_forward(x, 128, lengths=lengths)
        want.append(wft.augment_epilogue(plain, warps, masks, None, 0.0))
    torch.cuda.synchronize()
    assert any(int(w[0]) > 0 for w in wft.draw_warp_params(17, 0, B, 3000, 80, 0.6, cuda).cpu()), "some clip must be warped"
    outs = [torch.empty_like(want[0]) for _ in range(3)]
    for overlap in (False, True):
        old = wft.set_overlap(overlap)
        try:
            for i in range(9):
                fe(sets[i % 3], lengths=lengths, clip_offset=1000 * (i % 3), out=outs[i % 3])
            torch.cuda.synchronize()
        finally:
            wft.set_overlap(old)
        for k in range(3):
            assert torch.equal(outs[k], want[k]), f"batch {k}, overlap={overlap}"
            outs[k].zero_()


@pytest.mark.parametrize("n_mels,dtype,spline", [(128, "f32", "f64"), (80, "i16", "f32")])
def test_fused_production_call_finishes_cells_like_the_fixup_grid(wft, cuda, n_mels, dtype, spline):
    """wft_frontend_augment_forward (front-end grid -> epilogue that finishes every cell as it loads it) against
    wft_frontend_forward (front-end grid + fix-up grid) followed by wft_augment_drawn_f32, bit for bit, on a batch that needs
    every kind of fix-up: a quiet clip and a clip with exact zeros (the max-8 floor binds), short clips (silent tiles that were
    never computed), partial-segment cuts (min-value pad), plus the full-length batch without `lengths` (floor-only instance)."""
    B = 8
    g = torch.Generator().manual_seed(33)
    x = (0.1 * torch.randn(B, 480000, generator=g)).clamp(-1, 1)
    x[0] *= 1e-4
    x[1, 150000:] = 0.0
    x[2, :240000] *= 1e-3                      # the floor binds on the first half only
    lengths = torch.tensor([480000, 480000, 480000, 16000, 300001, 479999, 1000, 480000], dtype=torch.int32)
    n_valid = torch.tensor([-1, 1200, -1, 90, 3000, 1, -1, 2999], dtype=torch.int32)
    x[torch.arange(480000)[None, :] >= lengths[:, None]] = 0.0
    if dtype == "i16":
        x = (x * 32767).round().to(torch.int16)
    x, lengths, n_valid = x.to(cuda), lengths.to(cuda), n_valid.to(cuda)
    ext = torch.tensor([[0, 0], [3, 0], [0, 5], [2, 2], [0, 0], [0, 0], [1, 0], [0, 0]], dtype=torch.int32, device=cuda)
    for ragged in (True, False):
        L, NV = (lengths, n_valid) if ragged else (None, None)
        plain = wft.frontend_forward(x, n_mels, lengths=L, n_valid_frames=NV)
        want = torch.empty_like(plain)
        torch.ops.wft.augment_drawn_out(plain, 9, 500, 100, 27, 80, 0.7, ext, 0.0, spline == "f32", want)
        scratch = torch.full_like(plain, float("nan"))      # whatever the scratch held must not matter
        got = torch.empty_like(plain)
        torch.ops.wft.frontend_augment_drawn_out(x, n_mels, 0, L, 3000, NV, 9, 500, 100, 27, 80, 0.7, ext, 0.0, spline == "f32",
                                                 scratch, got)
        torch.cuda.synchronize()
        assert torch.isfinite(got).all()
        assert torch.equal(got, want), f"ragged={ragged}: {(got != want).sum().item()} cells differ"
        wft._lib.load().wft_debug_set_augment_generic(1)      # and the generic instance of the finishing epilogue
        try:
            got2 = torch.empty_like(plain)
            torch.ops.wft.frontend_augment_drawn_out(x, n_mels, 0, L, 3000, NV, 9, 500, 100, 27, 80, 0.7, ext, 0.0, spline == "f32",
                                                     scratch.fill_(float("nan")), got2)
            torch.cuda.synchronize()
        finally:
            wft._lib.load().wft_debug_set_augment_generic(0)
        assert torch.equal(got2, want), f"generic instance, ragged={ragged}"
    # the floor really bound somewhere (the test would be vacuous otherwise): clip 1 is exact zeros after sample 150000
    assert (plain[1] == plain[1].min()).float().mean().item() > 0.5


@pytest.mark.parametrize("R,T", [(128, 3000), (80, 3000), (128, 1500), (5, 40), (33, 516)])
def test_staged_and_generic_epilogue_agree_bit_for_bit(wft, cuda, R, T):
    """The epilogue has two kernels: the staged one (source windows by bulk copy through shared memory, n_frames % 4 == 0) and
    the generic one (taps straight from global memory).  Same arithmetic in the same order: equal bit for bit, for explicit and
    drawn parameters, both splines, steep warps (windows wider than the staging buffer fall back inside the staged kernel),
    clips without a warp, masks at the edges, extremes masks."""
    lib = wft._lib.load()
    B = 9
    g = torch.Generator().manual_seed(R * 10000 + T)
    mel = torch.randn(B, R, T, generator=g).to(cuda)
    W = min(80, (T - 1) // 2 - 1)
    # steepest maps the draw can produce (warp point next to an edge, full displacement), the identity, a rejected clip
    wp = [W, W, T - W - 1, T - W - 1, T // 2, T // 3, -1, 1, T - 2]
    wd = [W - 1, -W, W - 1, -W, 0, W // 2, 0, -W, W - 1]
    warps = torch.tensor(list(zip(wp, wd)), dtype=torch.int32, device=cuda)
    masks = torch.tensor([[0, 0, 0, 0], [0, min(100, T), 0, 1], [T - 7, T, R - 2, R], [T // 2, T // 2 + 50, 1, 3], [0, 0, 0, R],
                          [3, 4, 0, 0], [10, 20, 1, 2], [0, T, 0, 0], [5, 5, 2, 2]], dtype=torch.int32, device=cuda)
    ext = torch.tensor([[0, 0], [1, 0], [0, 2], [1, 1], [0, 0], [0, 0], [2, 0], [0, 0], [0, 1]], dtype=torch.int32, device=cuda)

    def run():
        outs = []
        for f32 in (False, True):
            outs.append(torch.ops.wft.augment(mel, warps, masks, ext, 0.25, f32))
            outs.append(torch.ops.wft.augment(mel, warps, None, None, 0.0, f32))
            o = torch.empty_like(mel)
            torch.ops.wft.augment_drawn_out(mel, 77, 123, min(100, T // 4), min(27, R // 2), W, 0.8, ext, 0.0, f32, o)
            outs.append(o)
        outs.append(torch.ops.wft.augment(mel, None, masks, ext, -1.0, False))
        torch.cuda.synchronize()
        return outs

    staged = run()
    lib.wft_debug_set_augment_generic(1)
    try:
        generic = run()
    finally:
        lib.wft_debug_set_augment_generic(0)
    for k, (a, b) in enumerate(zip(staged, generic)):
        assert torch.equal(a, b), f"variant {k}: {(a != b).sum().item()} cells differ"


def test_small_batches_in_flight_never_share_a_scratch_buffer(wft, cuda):
    """B = 1: whole calls fit on the GPU side by side, so an independent launch may run while calls further back than its
    predecessor are still in flight.  The launch bookkeeping (ops._LAST_CALL) must keep every buffer of every call since the
    last waiting launch apart: 50 production calls and 50 masks-only calls on rotating outputs equal the ordinary launches."""
    params = {"time_mask_param": 100, "freq_mask_param": 27, "time_warp_w": 80, "p": 1.0}
    fe = wft.FrontEnd(n_mels=128, spec_augment=True, spec_augment_params=params, seed=3)
    fe_plain = wft.FrontEnd(n_mels=128)
    g = torch.Generator().manual_seed(5)
    clips = [(0.1 * torch.randn(1, 480000, generator=g)).clamp(-1, 1).to(cuda) for _ in range(5)]
    want = [fe(c, clip_offset=k).clone() for k, c in enumerate(clips)]
    want_plain = [fe_plain(c).clone() for c in clips]
    torch.cuda.synchronize()
    outs = [torch.empty_like(want[0]) for _ in range(5)]
    old = wft.set_overlap(True)
    try:
        for i in range(50):
            fe(clips[i % 5], clip_offset=i % 5, out=outs[i % 5])
        torch.cuda.synchronize()
        for k in range(5):
            assert torch.equal(outs[k], want[k]), f"production call {k}"
        for i in range(50):
            fe_plain(clips[i % 5], out=outs[i % 5])
        torch.cuda.synchronize()
        for k in range(5):
            assert torch.equal(outs[k], want_plain[k]), f"plain call {k}"
    finally:
        wft.set_overlap(old)


def test_front_end_calls_survive_cuda_graph_capture_and_replay(wft, cuda):
    """A captured front-end step is replayed verbatim: the call must not bake a host-side ring position or an overlap decision
    into the graph (ops._launch_frontend switches to a private workspace that is zeroed by a memset node of the graph).
    Replays with new PCM in the captured input buffer equal eager calls, for the masks-only and the production (time-warp) path."""
    B = 4
    g = torch.Generator().manual_seed(8)
    batches = [(0.1 * torch.randn(B, 480000, generator=g)).clamp(-1, 1).to(cuda) for _ in range(3)]
    p_masks = {"time_mask_param": 100, "freq_mask_param": 43, "p": 1.0}
    p_warp = dict(p_masks, time_warp_w=80)
    old = wft.set_overlap(True)    # must be ignored while capturing
    try:
        for params in (p_masks, p_warp):
            fe = wft.FrontEnd(n_mels=128, spec_augment=True, spec_augment_params=params, seed=12)
            want = [fe(x, clip_offset=40).clone() for x in batches]
            static_in = batches[0].clone()
            static_out = torch.empty_like(want[0])
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fe(static_in, clip_offset=40, out=static_out)      # warm-up outside the capture
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fe(static_in, clip_offset=40, out=static_out)
                fe(static_in, clip_offset=40, out=static_out)      # twice: two calls of one graph must not share counters either
            for k in (1, 2, 0, 1):
                static_in.copy_(batches[k])
                graph.replay()
                torch.cuda.synchronize()
                assert torch.equal(static_out, want[k]), f"replay of batch {k} ({'warp' if 'time_warp_w' in params else 'masks'})"
    finally:
        wft.set_overlap(old)


def _gold(name):
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name), allow_pickle=False)


def test_front_end_against_reference_calculate_mel_goldens(wft, cuda):
    """The fused kernel against outputs of the reference's own AudioDataset._calculate_mel (tests/golden/make_golden.py):
    features within tolerance, masked cells and min-pad structure exact."""
    z = _gold("calculate_mel.npz")
    for k in range(int(z["n"])):
        n, n_mels, nv, tp, fp, seed = (int(v) for v in z[f"meta{k}"])
        x = S.make(str(z[f"kind{k}"]), n=n, seed=seed)
        fe = wft.FrontEnd(n_mels=n_mels)
        got = fe(x.unsqueeze(0).to(cuda), n_valid_frames=None if nv < 0 else [nv],
                 mask_params=None if tp == 0 else z[f"mask{k}"][None, :])[0].cpu()
        want_sub = torch.from_numpy(z[f"sub{k}"])
        got_sub = got[:, ::16]
        # tolerance on the frames that are features; the min-value pad (one value replicated over the tail) is
        # checked for what pad_or_trim guarantees: constant, equal to OUR kept minimum, and close to the reference's
        n_feat = 3000 if nv < 0 else nv
        cols = torch.arange(0, 3000, 16) < n_feat
        ma, rl = S.metrics(got_sub[:, cols], want_sub[:, cols])
        assert ma <= S.MAX_ABS and rl <= S.REL_L2, (k, ma, rl)
        if n_feat < 3000:
            t0, t1, f0, f1 = (int(v) for v in z[f"mask{k}"]) if tp > 0 else (0, 0, 0, 0)
            pad = got[:, n_feat:]
            keepmask = torch.ones_like(pad, dtype=torch.bool)
            keepmask[f0:f1, :] = False
            keepmask[:, max(t0 - n_feat, 0):max(t1 - n_feat, 0)] = False
            unmasked = fe(x.unsqueeze(0).to(cuda), n_valid_frames=[nv])[0].cpu()  # same call without masks
            assert (pad[keepmask] == unmasked[:, :n_feat].min()).all()
            if (~cols).any():
                assert (got_sub[:, ~cols] - want_sub[:, ~cols]).abs().max() <= S.MAX_ABS
        assert torch.equal(got_sub == 0, want_sub == 0)
        assert float((got == 0).sum()) == float(z[f"sum{k}"][2])
        assert abs(got.double().sum().item() - z[f"sum{k}"][0]) <= 1e-4 * z[f"sum{k}"][1]


def test_logmel_against_transformers_goldens(wft, cuda):
    z = _gold("logmel_hf.npz")
    for k in range(int(z["n"])):
        n, n_mels, seed = (int(v) for v in z[f"meta{k}"])
        x = S.make(str(z[f"kind{k}"]), n=n, seed=seed)
        got = wft.log_mel_spectrogram(x.to(cuda), n_mels=n_mels).cpu().numpy()
        assert np.abs(got - z[f"out{k}"]).max() <= 3e-4


def test_n_frames_out_longer_and_shorter_than_clip(wft, cuda):
    """pad_or_trim fused both ways: output longer than the clip (min pad) and shorter (trim)."""
    from oracle.pad_or_trim import pad_or_trim as ref_pad

    x = S.make("white", n=40000, seed=77)
    ref = O.log_mel_spectrogram(x, 80)  # [80, 250]
    for T in (96, 250, 251, 300, 1000):
        got = wft.frontend_forward(x.unsqueeze(0).to(cuda), 80, n_frames_out=T)[0].cpu()
        want = ref_pad(ref, T)
        assert got.shape == want.shape
        ma, rl = S.metrics(got, want)
        assert ma <= S.MAX_ABS and rl <= S.REL_L2
        if T > 250:
            assert (got[:, 250:] == got[:, :250].min()).all()


def test_repeated_launches_and_streams_are_deterministic(wft, cuda):
    x = torch.stack([S.make("white", n=480000, seed=s) for s in range(8)]).to(cuda)
    ref = wft.log_mel_spectrogram(x, n_mels=128)
    s1 = torch.cuda.Stream()
    with torch.cuda.stream(s1):
        for _ in range(20):
            again = wft.log_mel_spectrogram(x, n_mels=128)
            assert torch.equal(again, ref)
    s1.synchronize()


def test_host_pipeline_matches_direct_call(wft, cuda):
    fe = wft.FrontEnd(n_mels=128, spec_augment=True,
                      spec_augment_params={"time_mask_param": 100, "freq_mask_param": 27, "p": 1.0}, seed=3)
    host = torch.stack([S.make("white", n=480000, seed=40 + s) for s in range(6)]).pin_memory()
    out = torch.empty(6, 128, 3000).pin_memory()
    pipe = wft.HostPipeline(fe, 6, n_chunks=3, n_streams=2)
    pipe(host, out, clip_offset=100)
    pipe.synchronize()
    direct = fe(host.to(cuda), clip_offset=100).cpu()
    assert torch.equal(out, direct)
    # two batches back to back without a synchronisation in between (they overlap on the side streams), then join():
    # the current stream sees both results
    host_b = torch.stack([S.make("white", n=480000, seed=60 + s) for s in range(6)]).pin_memory()
    out_b = torch.empty(6, 128, 3000).pin_memory()
    out.zero_()
    pipe(host, out, clip_offset=100)
    pipe(host_b, out_b, clip_offset=106)
    pipe.join()
    torch.cuda.current_stream().synchronize()
    assert torch.equal(out, direct)
    assert torch.equal(out_b, fe(host_b.to(cuda), clip_offset=106).cpu())
    host16 = (host * 32767).round().to(torch.int16).pin_memory()
    pipe16 = wft.HostPipeline(fe, 6, pcm_dtype=torch.int16, n_chunks=2, n_streams=2)
    pipe16(host16, out, clip_offset=100)
    pipe16.synchronize()
    assert torch.equal(out, fe(host16.to(cuda), clip_offset=100).cpu())


@pytest.mark.parametrize("dtype", ["f32", "i16"])
def test_zero_padded_tails_take_the_silent_path(wft, cuda, dtype):
    """Clips much shorter than 30 s: every tile past the valid samples is all-zero PCM (skipped FFT, constant rows written
    by the fix-up).  Lengths straddle the tile geometry (tile = 16 frames = 2560 samples, halo 200/240)."""
    lens = [0, 1, 199, 200, 201, 2359, 2360, 2361, 2559, 2560, 2561, 5000, 16000, 239999, 240000, 476999, 479759, 479760, 480000]
    B = len(lens)
    pcm = torch.zeros(B, 480000)
    for b, n in enumerate(lens):
        if n:
            pcm[b, :n] = S.make("white", n=n, seed=500 + b)
    pcm[:, 300000:] += 0.0  # keep the tail exactly zero even where lengths would hide it
    if dtype == "i16":
        pcm = torch.round(pcm * 32767).to(torch.int16)
    lengths = np.asarray(lens, dtype=np.int32)
    masks = OS.draw_mask_params(9, 0, B, 80, 3000, 100, 43, 1.0)
    n_valid = np.full(B, -1, dtype=np.int32)
    n_valid[5], n_valid[11], n_valid[13] = 7, 2000, 1499
    fe = wft.FrontEnd(n_mels=80)
    got = fe(pcm.to(cuda), lengths=lengths, n_valid_frames=n_valid, mask_params=masks).cpu()
    ref = OP.front_end_batch(pcm, 80, lengths=lengths, n_valid_frames=n_valid, masks=masks)
    for b in range(B):
        keep = 3000 if n_valid[b] < 0 else int(n_valid[b])
        _check(got[b, :, :keep], ref[b, :, :keep], f"len={lens[b]} (kept frames)")
        assert (got[b] - ref[b]).abs().max() <= S.MAX_ABS
        assert torch.equal(got[b] == 0, ref[b] == 0)
    # garbage beyond `lengths` must not leak in: same result when the hidden samples are non-zero
    noisy = pcm.clone().float()
    for b, n in enumerate(lens):
        noisy[b, n:] = 0.25
    if dtype == "i16":
        noisy = torch.round(noisy).to(torch.int16) if pcm.dtype == torch.int16 else noisy
        noisy = torch.where(torch.arange(480000)[None, :] < torch.as_tensor(lens)[:, None], pcm, torch.full_like(pcm, 8191))
    got2 = fe(noisy.to(cuda), lengths=lengths, n_valid_frames=n_valid, mask_params=masks).cpu()
    assert torch.equal(got2, got)


def test_time_warp_matches_oracle_and_reference_golden(wft, cuda):
    from oracle import timewarp as OT

    z = _gold("timewarp.npz")
    for k in range(int(z["n"])):
        n_mels, W, seed, wp, wd, low, high = (int(v) for v in z[f"meta{k}"])
        x = S.make(str(z[f"kind{k}"]), seed=seed)
        mel = O.log_mel_spectrogram(x, n_mels)
        got = wft.time_warp(mel.to(cuda), torch.tensor([[wp, wd]], dtype=torch.int32)).cpu()
        ref = OT.time_warp(mel, wp, wd)
        assert (got - ref).abs().max() <= 2e-5, "CUDA vs oracle (same float64 spline)"
        assert np.abs(got[::8].numpy() - z[f"warp{k}"]).max() <= 1e-3, "CUDA vs the reference class (float32 spline)"
        # the float32 restatement of the reference's own spline arithmetic (data/utils.py:65-93): the source coordinate
        # lands on the reference's value wherever torch.pow is correctly rounded, so all but a few frames agree to float32
        # rounding of the bilinear blend; a frame whose coordinate differs by one ulp moves by ~1e-4 of a frame
        got32 = wft.augment_epilogue(mel.unsqueeze(0).to(cuda), torch.tensor([[wp, wd]], dtype=torch.int32), spline="f32")[0].cpu()
        d32 = np.abs(got32[::8].numpy() - z[f"warp{k}"])
        assert d32.max() <= 1e-3 and np.mean(d32 > 2e-6) <= 0.02, (d32.max(), np.mean(d32 > 2e-6))
        # drop-in class: same torch seed -> same (warp_p, warp_d) as the reference
        torch.manual_seed(seed)
        via_class = wft.TimeWarpAugmenter(W=W)(mel.to(cuda)).cpu()
        assert torch.equal(via_class, got)
    batch = torch.stack([mel, mel.flip(1)]).to(cuda)
    wps = torch.tensor([[700, 33], [2500, -41]], dtype=torch.int32)
    out = wft.time_warp(batch, wps).cpu()
    for b in range(2):
        assert (out[b] - OT.time_warp(batch[b].cpu(), int(wps[b, 0]), int(wps[b, 1]))).abs().max() <= 2e-5
    draws = wft.draw_warp_params(42, 0, 4096, 3000, 80).cpu()
    assert (draws[:, 0] >= 80).all() and (draws[:, 0] < 2920).all() and (draws[:, 1] >= -80).all() and (draws[:, 1] < 80).all()
    assert torch.equal(draws[100:200], wft.draw_warp_params(42, 100, 100, 3000, 80).cpu())
    assert (wft.draw_warp_params(42, 0, 16, 3000, 80, p=0.0).cpu() == torch.tensor([-1, 0])).all(), "(-1, 0) = no warp"
    same = wft.time_warp(batch, torch.tensor([[-1, 0], [2500, -41]], dtype=torch.int32))
    assert torch.equal(same[0], batch[0]) and torch.equal(same[1].cpu(), out[1]), "warp_p = -1 copies the clip"
    with pytest.raises(ValueError):
        wft.time_warp(batch, wps, out=torch.empty(2, 80, 2999, device=cuda))


def test_extremes_mask_matches_reference_golden(wft, cuda):
    z = _gold("timewarp.npz")
    for k in range(int(z["n"])):
        n_mels, W, seed, wp, wd, low, high = (int(v) for v in z[f"meta{k}"])
        x = S.make(str(z[f"kind{k}"]), seed=seed)
        mel = O.log_mel_spectrogram(x, n_mels).to(cuda)
        keep = mel.clone()
        torch.manual_seed(seed + 100)
        res = wft.ExtremesFrequencyMasking(10, 15)(mel)
        assert res.data_ptr() == mel.data_ptr(), "in place, like the reference"
        zero_rows = (res == 0).all(dim=1).cpu().numpy()
        assert np.array_equal(zero_rows, z[f"ext_zero_rows{k}"])
        assert torch.equal(res[~torch.from_numpy(zero_rows).to(cuda)], keep[~torch.from_numpy(zero_rows).to(cuda)])


def test_front_end_with_time_warp_follows_reference_order(wft, cuda):
    """warp -> time mask -> frequency mask (data_loader.py:285-287)."""
    from oracle import timewarp as OT

    x = torch.stack([S.make("white", seed=61), S.make("chirp", seed=62)])
    fe = wft.FrontEnd(n_mels=80, spec_augment=True, seed=5,
                      spec_augment_params={"time_mask_param": 100, "freq_mask_param": 43, "time_warp_w": 80, "p": 1.0})
    assert fe.time_warp_w == 80, "a positive time_warp_w turns the warp on, like the reference (data_loader.py:117)"
    got = fe(x.to(cuda), clip_offset=10).cpu()
    warps = wft.draw_warp_params(5, 10, 2, 3000, 80).cpu().numpy()
    masks = OS.draw_mask_params(5, 10, 2, 80, 3000, 100, 43, 1.0)
    for b in range(2):
        ref = OS.apply_masks(OT.time_warp(O.log_mel_spectrogram(x[b], 80), int(warps[b, 0]), int(warps[b, 1])), *masks[b])
        assert (got[b] - ref).abs().max() <= 1e-3
        assert torch.equal(got[b] == 0, ref == 0)


def test_torch_custom_ops_pass_opcheck(wft, cuda):
    """torch.library.opcheck: schema, fake kernel, autograd registration and AOT dispatch of every torch.ops.wft op."""
    from torch.library import opcheck

    g = torch.Generator().manual_seed(1)
    pcm = (0.1 * torch.randn(3, 48000, generator=g)).to(cuda)
    lengths = torch.tensor([48000, 20000, 333], dtype=torch.int32, device=cuda)
    nv = torch.tensor([-1, 100, 3], dtype=torch.int32, device=cuda)
    masks = torch.tensor([[10, 60, 3, 9], [0, 0, 0, 0], [290, 300, 70, 80]], dtype=torch.int32, device=cuda)
    opcheck(torch.ops.wft.frontend_forward, (pcm, 80, 0, lengths, 300, nv, masks, 0.0))
    opcheck(torch.ops.wft.frontend_forward, ((pcm * 32767).to(torch.int16), 128, 160, None, 0, None, None, 0.0))
    opcheck(torch.ops.wft.frontend_forward_out, (pcm, 80, 0, None, 300, None, masks, 0.0, torch.empty(3, 80, 300, device=cuda)))
    opcheck(torch.ops.wft.frontend_forward_drawn_out,
            (pcm, 80, 0, lengths, 300, nv, 42, 7, 100, 27, 0.5, 0.0, torch.empty(3, 80, 300, device=cuda)))
    opcheck(torch.ops.wft.frontend_augment_drawn_out,
            (pcm, 80, 0, lengths, 300, nv, 42, 7, 100, 27, 20, 0.5, None, 0.0, False, torch.empty(3, 80, 300, device=cuda),
             torch.empty(3, 80, 300, device=cuda)))
    drawn = torch.empty(3, 80, 300, device=cuda)
    torch.ops.wft.frontend_forward_drawn_out(pcm, 80, 0, lengths, 300, nv, 42, 7, 100, 27, 1.0, 0.0, drawn)
    assert torch.equal(drawn, torch.ops.wft.frontend_forward(pcm, 80, 0, lengths, 300, nv,
                                                               wft.draw_mask_params(42, 7, 3, 80, 300, 100, 27, 1.0), 0.0))
    opcheck(torch.ops.wft.pad_or_trim, (torch.randn(2, 70, 5, device=cuda), 300))
    opcheck(torch.ops.wft.pad_or_trim, (torch.randn(2, 70, 5, device=cuda), 30))
    mel = torch.randn(3, 80, 300, device=cuda)
    warps = torch.tensor([[100, 7], [-1, 0], [200, -9]], dtype=torch.int32, device=cuda)
    ext = torch.tensor([[2, 3], [0, 0], [5, 0]], dtype=torch.int32, device=cuda)
    opcheck(torch.ops.wft.augment, (mel, warps, masks, ext, 0.0, False))
    opcheck(torch.ops.wft.augment, (mel, warps, None, None, 0.0, True))
    opcheck(torch.ops.wft.augment_out, (mel, warps, masks, ext, 0.0, False, torch.empty_like(mel)))
    opcheck(torch.ops.wft.augment_, (mel.clone(), masks, ext, 0.0))
    opcheck(torch.ops.wft.specaug_apply, (mel, masks, 0.0))
    opcheck(torch.ops.wft.specaug_apply_, (mel.clone(), masks, 0.0))
    opcheck(torch.ops.wft.specaug_draw, (mel, 42, 7, 5, 80, 300, 100, 27, 0.5))
    opcheck(torch.ops.wft.time_warp_draw, (mel, 42, 7, 5, 300, 20, 0.5))
    for dtype in (torch.float32, torch.bfloat16):
        act = torch.randn(2, 50, 64, device=cuda, dtype=dtype, requires_grad=True)
        opcheck(torch.ops.wft.mask_bsd, (act, 3, 11, 8, 20))
    # the ops are what the public functions run on
    assert torch.equal(wft.frontend_forward(pcm, 80, lengths=lengths, n_frames_out=300, n_valid_frames=nv, mask_params=masks),
                       torch.ops.wft.frontend_forward(pcm, 80, 0, lengths, 300, nv, masks, 0.0))


def test_logmel_against_transformers_torch_extractor_full_clips(wft, cuda):
    """CUDA against the float32 torch.stft path of the transformers extractor (tests/golden/logmel_hf_torch.npz), which the
    oracle matches to one ulp: full 30-s clips, every signal kind, 80 / 128 mel."""
    z = _gold("logmel_hf_torch.npz")
    for k in range(int(z["n"])):
        n_mels = int(z[f"meta{k}"][0])
        got = wft.log_mel_spectrogram(S.make(str(z[f"kind{k}"])).to(cuda), n_mels=n_mels).cpu()
        for name, view in (("sub", got[:, ::24]), ("head", got[:, :32]), ("tail", got[:, -32:])):
            ma, rl = S.metrics(view, torch.from_numpy(z[f"{name}{k}"]))
            assert ma <= S.MAX_ABS and rl <= S.REL_L2, (k, name, ma, rl)
        assert np.abs(got.double().sum(dim=0).numpy() - z[f"colsum{k}"]).max() <= n_mels * 1e-4


def test_pad_or_trim_keeps_half_precision_dtypes(wft, cuda):
    from oracle.pad_or_trim import pad_or_trim as ref_pad

    for dtype in (torch.float16, torch.bfloat16):
        x = (torch.randn(8, 70) - 0.3).to(dtype)
        got = wft.pad_or_trim(x.to(cuda), 300)
        assert got.dtype == dtype and torch.equal(got.cpu(), ref_pad(x, 300))
    with pytest.raises(TypeError):
        wft.pad_or_trim(torch.zeros(4, 7, dtype=torch.float64, device=cuda), 30)


def test_augment_epilogue_is_warp_then_masks_then_extremes(wft, cuda):
    """One pass == the reference's sequence time_warping -> time_masking -> freq_masking -> extreme_freq_masking
    (data_loader.py:284-290), and a front end built from the reference's config block runs exactly that."""
    from oracle import timewarp as OT

    x = torch.stack([S.make("white", seed=71), S.make("hdr", seed=72), S.make("chirp", seed=73)])
    mel = O.log_mel_batch(x, 128)
    warps = torch.tensor([[700, 33], [-1, 0], [2500, -41]], dtype=torch.int32)
    masks = OS.draw_mask_params(3, 0, 3, 128, 3000, 100, 27, 1.0)
    ext = torch.tensor([[4, 9], [0, 0], [10, 0]], dtype=torch.int32)
    got = wft.augment_epilogue(mel.to(cuda), warps, masks, ext).cpu()
    for b in range(3):
        ref = mel[b] if warps[b, 0] < 0 else OT.time_warp(mel[b], int(warps[b, 0]), int(warps[b, 1]))
        ref = OT.extremes_mask(OS.apply_masks(ref, *masks[b]), int(ext[b, 0]), int(ext[b, 1]))
        assert (got[b] - ref).abs().max() <= 2e-5
        assert torch.equal(got[b] == 0, ref == 0)
    # masks + extremes only: in place, bit exact
    buf = mel.to(cuda).clone()
    res = wft.augment_epilogue(buf, None, masks, ext, out=buf)
    assert res.data_ptr() == buf.data_ptr()
    for b in range(3):
        assert torch.equal(res[b].cpu(), OT.extremes_mask(OS.apply_masks(mel[b], *masks[b]), int(ext[b, 0]), int(ext[b, 1])))
    with pytest.raises(ValueError):
        wft.augment_epilogue(buf, warps, out=buf)      # a warp cannot run in place
    # ragged frame count (scalar stores) and a tiny tensor
    small = torch.randn(2, 5, 37)
    got = wft.augment_epilogue(small.to(cuda), torch.tensor([[10, 3], [20, -4]], dtype=torch.int32),
                               torch.tensor([[3, 9, 1, 2], [0, 0, 0, 5]], dtype=torch.int32)).cpu()
    for b, (wp, wd, mk) in enumerate([(10, 3, (3, 9, 1, 2)), (20, -4, (0, 0, 0, 5))]):
        assert (got[b] - OS.apply_masks(OT.time_warp(small[b], wp, wd), *mk)).abs().max() <= 2e-5


def test_upstream_gate_is_not_rolled_twice(wft, cuda):
    """with the gate decided upstream (`augment`), p = 0.5 must not thin the augmented clips a second time."""
    B = 64
    x = torch.stack([S.make("white", n=16000, seed=900 + b) for b in range(B)])
    fe = wft.FrontEnd(n_mels=80, spec_augment=True, seed=9,
                      spec_augment_params={"time_mask_param": 100, "freq_mask_param": 43, "time_warp_w": 80, "p": 0.5})
    flags = (torch.arange(B) % 3 != 0).to(torch.int32)
    got = fe(x.to(cuda), clip_offset=0, augment=flags).cpu()
    plain = wft.FrontEnd(n_mels=80)(x.to(cuda)).cpu()
    masks = OS.draw_mask_params(9, 0, B, 80, 3000, 100, 43, 1.0)
    for b in range(B):
        if flags[b]:
            t0, t1, f0, f1 = masks[b]
            assert (got[b][f0:f1] == 0).all() and (got[b][:, t0:t1] == 0).all(), "every gated-in clip carries its masks"
            assert not torch.equal(got[b], plain[b])
        else:
            assert torch.equal(got[b], plain[b]), "a gated-out clip gets neither warp nor masks"
    # without an upstream gate the device rolls p itself: about half of the clips, the same ones for warp and masks
    own = fe(x.to(cuda), clip_offset=0).cpu()
    changed = torch.tensor([not torch.equal(own[b], plain[b]) for b in range(B)])
    assert 16 <= int(changed.sum()) <= 48
    gate_ref = OS.draw_mask_params(9, 0, B, 80, 3000, 100, 43, 0.5).any(axis=1)
    assert np.array_equal(changed.numpy(), gate_ref)


def test_axis_masks_keep_dtype_and_gradients(wft, cuda):
    """the drop-in mask classes may meet bf16 activations that require grad (model_utils.py:404-405)."""
    for dtype in (torch.bfloat16, torch.float16, torch.float32):
        x = torch.randn(2, 64, 50, device=cuda, dtype=dtype, requires_grad=True)
        torch.manual_seed(4)
        y = wft.FrequencyMasking(9)(wft.TimeMasking(11)(x))
        assert y.dtype == dtype and y.requires_grad
        y.float().sum().backward()
        assert x.grad is not None and x.grad.dtype == dtype
        torch.manual_seed(4)
        import torchaudio.transforms as T

        ref = T.FrequencyMasking(9)(T.TimeMasking(11)(x.detach()))
        assert torch.equal(y.detach(), ref)
        assert torch.equal(x.grad == 0, ref == 0) or (x.detach() == 0).any()


# ---- deep SpecAugment on activations (model/model_utils.py:382-437) ----------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(4, 1500, 1280), (2, 75, 384), (3, 33, 100), (1, 7, 5)])
def test_mask_activations_matches_oracle_bit_exact(wft, cuda, dtype, shape):
    from oracle.deep_specaug import deep_spec_augment

    g = torch.Generator().manual_seed(3)
    x = torch.randn(shape, generator=g).to(dtype)
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        want, t, f = deep_spec_augment(x, 100, 43)
        torch.manual_seed(seed)
        t2, f2 = wft.draw_deep_spans(shape[1], shape[2], 100, 43)
        assert (t2, f2) == (t, f)
        got = wft.mask_activations(x.to(cuda), t2, f2)
        assert got.dtype == dtype and torch.equal(got.cpu(), want)


@pytest.mark.gpu
def test_mask_activations_backward_is_the_same_mask(wft, cuda):
    x = torch.randn(2, 40, 64, device=cuda, dtype=torch.bfloat16, requires_grad=True)
    y = wft.mask_activations(x, (5, 17), (60, 64))
    gy = torch.randn_like(y)
    y.backward(gy)
    want = gy.clone()
    want[:, 5:17, :] = 0
    want[:, :, 60:64] = 0
    assert torch.equal(x.grad, want)
    # unaligned views take the scalar path and give the same answer
    base = torch.randn(2, 40, 65, device=cuda)
    v = base[:, :, 1:]
    out = wft.mask_activations(v, (0, 3), (10, 20))
    ref = v.clone()
    ref[:, 0:3, :] = 0
    ref[:, :, 10:20] = 0
    assert torch.equal(out, ref)


@pytest.mark.gpu
def test_deep_spec_augment_hooks_match_reference_hook(wft, cuda):
    """Same seed -> the hooked toy encoder produces what the reference hook body (torchaudio masks) produces."""
    import torch.nn as nn
    import torchaudio.transforms as T

    class Block(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.attn_ln = nn.LayerNorm(d)

        def forward(self, x):
            return x + self.attn_ln(x)

    class Enc(nn.Module):
        def __init__(self, d, n):
            super().__init__()
            self.blocks = nn.ModuleList([Block(d) for _ in range(n)])

        def forward(self, x):
            for b in self.blocks:
                x = b(x)
            return x

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = Enc(64, 3)

    torch.manual_seed(0)
    ours, ref = Model().to(cuda).train(), Model().to(cuda).train()
    ref.load_state_dict(ours.state_dict())
    wft.register_deep_spec_augment_hooks(ours, 20, 9, p=1.0)
    tm, fm = T.TimeMasking(time_mask_param=20), T.FrequencyMasking(freq_mask_param=9)

    def ref_hook(module, inp, out):   # body of the reference's _norm_hook (model_utils.py:412-421)
        if module.training:
            return fm(tm(out.permute(0, 2, 1))).permute(0, 2, 1)
        return out

    for i in range(2):  # the last block is skipped, like the reference
        ref.encoder.blocks[i].attn_ln.register_forward_hook(ref_hook)
    x = torch.randn(2, 50, 64, device=cuda)
    torch.manual_seed(11)
    a = ours.encoder(x)
    torch.manual_seed(11)
    b = ref.encoder(x)
    assert torch.equal(a, b)
    ours.eval()
    assert torch.equal(ours.encoder(x), ref.eval().encoder(x))
    with pytest.raises(ValueError):
        wft.register_deep_spec_augment_hooks(ours, 20, 9, p=1.5)


# ---- loader integration: PCM records -> pcm_collate_fn -> DeviceFrontEndLoader ---------------------------
class _StandInDataset:
    """The attribute surface ``deferred_calculate_mel`` uses of the reference's AudioDataset (data_loader.py:60-150)."""

    def __init__(self, n_mels, p, extremes=None):
        self.aud_augment = None
        self.n_mels = n_mels
        self.num_frames_per_second = 3000 / 30
        self.spec_augment = p > 0
        self.spec_augment_p = p
        self.extreme_freq_masking = extremes

    def _should_apply_spec_augment(self):  # data_loader.py:294-301
        if not self.spec_augment:
            return False
        if self.spec_augment_p >= 1.0:
            return True
        if self.spec_augment_p <= 0.0:
            return False
        return torch.rand(1).item() < self.spec_augment_p


@pytest.mark.gpu
@pytest.mark.parametrize("pcm_dtype", [torch.float32, torch.int16])
def test_device_front_end_loader_matches_oracle(wft, cuda, pcm_dtype):
    from types import SimpleNamespace

    from whisper_finetune_b200 import loader as L

    rng = np.random.default_rng(7)
    B, n_mels = 6, 80
    ds = _StandInDataset(n_mels, 0.5, extremes=SimpleNamespace(low_freq_range=10, high_freq_range=6))
    L._PCM_DTYPE["dtype"] = pcm_dtype
    clips, starts, items = [], [], []
    torch.manual_seed(5)
    for b in range(B):
        n = int(rng.integers(16000, 480001)) if b else 480000
        a = S.make("int16" if b % 2 else "white", n=n, seed=300 + b)
        a = (a.float() / 32768.0 if a.dtype == torch.int16 else a).numpy()
        if pcm_dtype == torch.int16:
            a = np.round(a * 32768.0).clip(-32768, 32767).astype(np.float32) / 32768.0   # what a 16-bit source decodes to
        start = None if b % 3 else float(rng.uniform(0.5, 29.0))
        clips.append(a)
        starts.append(start)
        rec = wft.deferred_calculate_mel(ds, np.pad(a, (0, 480000 - n)), start, True)
        items.append((rec, torch.arange(3 + b), torch.arange(2 + b)))
    L._PCM_DTYPE["dtype"] = torch.float32
    batch, y_in, y_out = wft.pcm_collate_fn(items)
    assert batch.pcm.dtype == pcm_dtype and tuple(batch.pcm.shape) == (B, 480000)
    assert y_in.shape == (B, 3 + B - 1) and y_out[0, -1] == -100
    assert batch.lengths.tolist() == [int(np.flatnonzero(c)[-1]) + 1 for c in clips]
    want_nv = [-1 if s is None else int(s * 100) for s in starts]
    assert batch.n_valid_frames.tolist() == want_nv
    assert 0 < int(batch.augment.sum()) < B, "p = 0.5 with this seed should gate some clips on and some off"

    fe = wft.FrontEnd(n_mels=n_mels, spec_augment=True,
                      spec_augment_params={"time_mask_param": 100, "freq_mask_param": 27, "p": 1.0}, seed=9)
    dl = wft.DeviceFrontEndLoader([(batch, y_in, y_out)], fe, clip_offset=40)
    out = list(dl)
    assert len(out) == 1 and out[0][1] is y_in and out[0][2] is y_out
    x = out[0][0]
    assert x.is_cuda and tuple(x.shape) == (B, n_mels, 3000) and dl._batches == 1
    masks = OS.draw_mask_params(9, 40, B, n_mels, 3000, 100, 27, 1.0)
    # ranks of a DDP job draw from disjoint global clip indices: batch k of rank r starts at offset + (k * world + r) * B
    dl2 = wft.DeviceFrontEndLoader([(batch, y_in, y_out)] * 2, fe, clip_offset=40, rank=1, world_size=4)
    xs = [o[0].cpu() for o in dl2]
    for k, xk in enumerate(xs):
        mk = OS.draw_mask_params(9, 40 + (k * 4 + 1) * B, B, n_mels, 3000, 100, 27, 1.0)
        for b in range(B):
            if batch.augment[b]:
                t0, t1, f0, f1 = mk[b]
                assert (xk[b][:, t0:t1] == 0).all() and (xk[b][f0:f1] == 0).all()
    x = x.cpu()
    for b in range(B):
        nv = None if want_nv[b] < 0 else want_nv[b]
        ref = OP.calculate_mel(torch.from_numpy(clips[b]), n_mels, nv, masks[b] if batch.augment[b] else None)
        lo, hi = batch.extremes[b].tolist()
        ref[:lo] = 0
        if hi:
            ref[n_mels - hi:] = 0
        keep = 3000 if nv is None else nv
        _check(x[b, :, :keep], ref[:, :keep], f"loader clip {b}")
        assert (x[b] - ref).abs().max() <= S.MAX_ABS
        assert torch.equal(x[b] == 0, ref == 0), "masked / extreme cells are exactly 0.0, nothing else is"


@pytest.mark.gpu
def test_pcm_record_errors(wft, cuda):
    with pytest.raises(RuntimeError):
        wft.encode_pcm_record(np.ones(100, np.float32), n_valid_frames=0)       # empty spectrogram: torch.min raises upstream
    # a clip a few samples over 30 s (time-stretch augmentation) is cut at 480000 samples, not rejected
    long = wft.encode_pcm_record(np.full(480123, 0.5, np.float32))
    assert long.shape == (480008,) and wft.decode_pcm_records([long]).lengths.tolist() == [480000]
    bad = wft.encode_pcm_record(np.ones(100, np.float32))
    bad[480000] = 1
    with pytest.raises(ValueError):
        wft.decode_pcm_records([bad])
