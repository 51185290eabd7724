"""ORACLE (test infrastructure, never shipped): SpecAugment time-warp and extremes mask ("next" rows).

Restates ``TimeWarpAugmenter`` (``/root/reference/src/whisper_finetune/data/utils.py:41-143``, applied at
``data_loader.py:285``) for explicit warp parameters: a cubic Hermite spline through the three knots
``(0, -1), (warp_p, (warp_p - warp_d) * 2 / (T - 1) - 1), (T - 1, 1)`` (end slopes = secant slopes, middle slope = their mean)
gives the normalised source coordinate of every output frame; the mel is resampled along time with
``grid_sample(bilinear, zeros padding, align_corners=True)`` -- the same library op the reference calls.  The reference
draws ``warp_p = randint(W, T - W)`` then ``warp_d = randint(-W, W)`` from torch's global generator.

``ExtremesFrequencyMasking`` (``data/utils.py:146-190``, applied at ``data_loader.py:289-290``): one ``torch.rand(1)`` per
clip, ``round(r * range)`` lowest and highest rows set to zero.

PINNED by ``tests/golden/timewarp.npz`` (outputs of the reference classes, see make_golden.py).
"""
import torch


def warp_source_coords(n_frames: int, warp_p: int, warp_d: int) -> torch.Tensor:
    """float32 [T]: normalised (align_corners) source x-coordinate of each output frame.

    Evaluated in float64 and rounded once: the reference evaluates the same spline in float32 (pow / matmul), which moves
    the coordinate by up to ~1e-4 frames; parity for this row is therefore tolerance based."""
    T = n_frames
    x = torch.tensor([0.0, float(warp_p), float(T - 1)], dtype=torch.float64)
    y = torch.tensor([-1.0, (warp_p - warp_d) * 2 / (T - 1.0) - 1.0, 1.0], dtype=torch.float64)
    s = (y[1:] - y[:-1]) / (x[1:] - x[:-1])               # secant slopes of the two segments
    m = torch.stack([s[0], (s[0] + s[1]) / 2, s[1]])     # knot slopes
    xs = torch.arange(T, dtype=torch.float64)
    seg = (xs > x[1]).long()                               # searchsorted(x[1:], xs): 0 for xs <= warp_p, else 1
    x0, x1 = x[seg], x[seg + 1]
    dx = x1 - x0
    t = (xs - x0) / dx
    t2, t3 = t * t, t * t * t
    h00 = 1 - 3 * t2 + 2 * t3
    h10 = t - 2 * t2 + t3
    h01 = 3 * t2 - 2 * t3
    h11 = -t2 + t3
    return (h00 * y[seg] + h10 * m[seg] * dx + h01 * y[seg + 1] + h11 * m[seg + 1] * dx).to(torch.float32)


def time_warp(mel: torch.Tensor, warp_p: int, warp_d: int) -> torch.Tensor:
    """``mel`` [R, T] (or [C, R, T]: the same warp for every leading slice) -> warped copy."""
    squeeze = mel.dim() == 2
    x = mel.unsqueeze(0) if squeeze else mel
    C, R, T = x.shape
    ys = warp_source_coords(T, warp_p, warp_d)
    grid = torch.stack([ys.view(1, T).expand(R, T), torch.linspace(-1, 1, R).view(R, 1).expand(R, T)], dim=-1)
    out = torch.nn.functional.grid_sample(x.unsqueeze(0).float(), grid.unsqueeze(0), align_corners=True)[0]
    return out[0] if squeeze else out


def extremes_mask_lengths(r: float, low_freq_range: int, high_freq_range: int):
    return int(round(r * low_freq_range)), int(round(r * high_freq_range))


def extremes_mask(mel: torch.Tensor, low_len: int, high_len: int) -> torch.Tensor:
    out = mel.clone()
    n = out.shape[-2]
    if low_len > 0:
        out[..., : min(low_len, n), :] = 0
    if high_len > 0:
        out[..., max(n - high_len, 0):, :] = 0
    return out
