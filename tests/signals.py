"""Seeded synthetic PCM used by both CPU and GPU tests (SURVEY.md section 8d value distributions)."""
import numpy as np
import torch

SR = 16000


def make(kind: str, n: int = 480000, seed: int = 123) -> torch.Tensor:
    """float32 [n] (or int16 for kind='int16')."""
    t = np.arange(n, dtype=np.float64) / SR
    g = torch.Generator().manual_seed(seed)
    if kind == "white":
        return (0.1 * torch.randn(n, generator=g)).clamp(-1, 1)
    if kind == "hdr":  # loud tones + faint noise for 12 s, then digital silence: exposes precision shortcuts
        x = 0.5 * np.sin(2 * np.pi * 220.0 * t) + 0.2 * np.sin(2 * np.pi * 1375.3 * t)
        x = x + 1e-4 * torch.randn(n, generator=g).numpy()
        x[t >= 12.0] = 0.0
        return torch.from_numpy(x.astype(np.float32))
    if kind == "int16":  # decaying tone bursts + noise, quantised to 16 bit
        x = 0.3 * np.sin(2 * np.pi * 313.7 * t) * np.exp(-(t % 1.0)) + 3e-4 * torch.randn(n, generator=g).numpy()
        return torch.from_numpy(np.clip(np.round(32768.0 * x), -32768, 32767).astype(np.int16))
    if kind == "zeros":
        return torch.zeros(n)
    if kind == "impulse":
        x = torch.zeros(n)
        x[min(12345, n - 1)] = 1.0
        return x
    if kind == "chirp":
        f = 50.0 + (7900.0 - 50.0) * t / max(t[-1], 1e-9)
        return torch.from_numpy((0.25 * np.sin(2 * np.pi * np.cumsum(f) / SR)).astype(np.float32))
    raise KeyError(kind)


def metrics(got: torch.Tensor, ref: torch.Tensor):
    d = (got.double() - ref.double())
    return d.abs().max().item(), (d.norm() / ref.double().norm().clamp_min(1e-30)).item()


# tolerance of BASELINE.json north_star for float32 features
MAX_ABS = 1e-3
REL_L2 = 1e-5
