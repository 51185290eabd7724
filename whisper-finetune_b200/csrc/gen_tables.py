#!/usr/bin/env python
"""Generate ``wft_tables.inc``: the constant tables of the fused front-end kernel.

Everything here is a pure function of the Whisper front-end constants (``N_FFT=400``, ``HOP=160``,
Slaney mel bank for 80 / 128 rows) that the reference uses through ``whisper.audio`` at
``src/whisper_finetune/data/data_loader.py:13,278``:

* ``WFT_WINDOW_TABLE``  -- periodic Hann(400) exactly as ``torch.hann_window(400)`` rounds it in float32,
  permuted to the kernel's stage-A layout ``[n2][n1]`` = ``w[20*n1 + n2]`` (20-float rows: consecutive threads
  land in distinct 16-byte bank groups for LDS.128).
* ``WFT_TWIDDLE_TABLE`` -- ``W400^(n2 * k1)`` as (re, im), layout ``[n2][k1]``, rows padded from 40 to 44 floats
  (same bank argument).
* ``WFT_MEL<NM>_PROG`` -- the sparse triangular mel projection as a "row program".  Every mel row is one entry
  ``(first_bin * 4, row, w[0..C-1])`` whose tap count is padded to a class C in {2, 6, 10, 14} (entry = 4, 8, 12
  or 16 words, always 16-byte aligned); rows are dealt to the 10 warps of a CTA so that every warp carries the
  same cost, and inside a warp they are grouped by class so that the kernel runs four tiny loops instead of
  2000 unrolled instructions (the hot loop has to fit the SM's instruction cache).  Weights are 0.25 * bank
  weight because the packed two-frame FFT yields 4*|X|^2.  ``WFT_MEL<NM>_INDEX`` lists, per warp and class,
  (first_uint4, n_rows).

Run ``python gen_tables.py`` to rewrite the file next to this script (``__graft_entry__.build()`` does).
"""
import importlib.util
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
N_GROUPS = 10
TWIDDLE_ROW = 44  # floats per n2 twiddle row: 20 complex + 4 pad


def _load_melbank():
    spec = importlib.util.spec_from_file_location("_wft_melbank", os.path.join(HERE, "..", "melbank.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def flit(x) -> str:
    """Exact float32 literal."""
    v = np.float32(x)
    s = "%.9g" % float(v)
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    assert np.float32(float(s)) == v
    return s + "f"


def fbits(x) -> str:
    """float32 bit pattern as an unsigned literal (tables are stored as uint4)."""
    return "0x%08xu" % struct.unpack("<I", struct.pack("<f", float(np.float32(x))))[0]


def hann_window_f32() -> np.ndarray:
    import torch

    w = torch.hann_window(400, dtype=torch.float32).numpy().copy()
    ref = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(400) / 400.0)
    assert np.abs(w - ref).max() < 5e-7 and w[0] == 0.0
    return w


def window_table() -> np.ndarray:
    w = hann_window_f32()
    t = np.zeros((20, 20), dtype=np.float32)
    for n2 in range(20):
        for n1 in range(20):
            t[n2, n1] = w[20 * n1 + n2]
    return t


def twiddle_table() -> np.ndarray:
    t = np.zeros((20, TWIDDLE_ROW), dtype=np.float32)
    for n2 in range(20):
        for k1 in range(20):
            ang = -2.0 * np.pi * ((n2 * k1) % 400) / 400.0
            t[n2, 2 * k1] = np.float32(np.cos(ang))
            t[n2, 2 * k1 + 1] = np.float32(np.sin(ang))
    return t


MEL_CLASSES = (2, 6, 10, 14)


def row_support(bank: np.ndarray):
    supp = []
    for m in range(bank.shape[0]):
        nz = np.nonzero(bank[m])[0]
        assert len(nz) > 0 and nz[-1] - nz[0] + 1 == len(nz), "mel row support must be one contiguous band"
        assert nz[0] >= 1 and nz[-1] <= 199, "bins 0 and 200 must carry no weight"
        supp.append((int(nz[0]), int(nz[-1])))
    return supp


def row_class(n_taps: int) -> int:
    for c in MEL_CLASSES:
        if n_taps <= c:
            return c
    raise AssertionError(f"mel row with {n_taps} taps")


def row_program(bank: np.ndarray, n_groups: int = N_GROUPS):
    """-> (words: list of uint32 literals, index[g][class] = (first_uint4, n_rows), assignment[g] = rows)."""
    supp = row_support(bank)
    n_mels = bank.shape[0]
    cost = {m: 10.0 + 2.25 * row_class(supp[m][1] - supp[m][0] + 1) for m in range(n_mels)}
    load = [0.0] * n_groups
    rows_of = [[] for _ in range(n_groups)]
    for m in sorted(range(n_mels), key=lambda r: (-cost[r], r)):  # longest-processing-time first
        g = min(range(n_groups), key=lambda j: (load[j], j))
        rows_of[g].append(m)
        load[g] += cost[m]
    words, index = [], []
    for g in range(n_groups):
        idx_g = []
        for c in MEL_CLASSES:
            rows = sorted(m for m in rows_of[g] if row_class(supp[m][1] - supp[m][0] + 1) == c)
            assert len(words) % 4 == 0
            idx_g.append((len(words) // 4, len(rows)))
            for m in rows:
                k0, k1 = supp[m]
                if k0 + c - 1 > 199:  # padded taps must stay on bins 1..199 (rewritten every tile, always finite)
                    k0 = 199 - (c - 1)
                assert k0 >= 1
                entry = ["%du" % (4 * k0), "%du" % m]
                for j in range(c):
                    k = k0 + j
                    w = np.float32(bank[m, k]) * np.float32(0.25)
                    entry.append(fbits(w))
                assert len(entry) == c + 2
                words.extend(entry)
        index.append(idx_g)
    return words, index, rows_of


def simulate(bank: np.ndarray, words, index):
    """Replay the row program on a random power spectrum and compare with the dense product (float64)."""
    rng = np.random.default_rng(0)
    P = rng.random(202)
    vals = []
    for w in words:
        w = w.rstrip("u")
        vals.append(int(w, 16) if w.startswith("0x") else int(w))
    out = np.full(bank.shape[0], np.nan)
    for idx_g in index:
        for c, (first, n) in zip(MEL_CLASSES, idx_g):
            pos = first * 4
            for _ in range(n):
                k0, m = vals[pos] // 4, vals[pos + 1]
                acc = 0.0
                for j in range(c):
                    wgt = struct.unpack("<f", struct.pack("<I", vals[pos + 2 + j]))[0]
                    acc += wgt * P[k0 + j]
                assert np.isnan(out[m])
                out[m] = acc
                pos += c + 2
    ref = (bank.astype(np.float64) * 0.25) @ P[:201]
    assert np.allclose(out, ref, rtol=1e-12, atol=0), np.abs(out - ref).max()


def generate() -> str:
    mb = _load_melbank()
    lines = [
        "// GENERATED by gen_tables.py -- do not edit.  Constants of the Whisper front end",
        "// (Hann(400) window, W400 twiddles, Slaney mel banks for 80 / 128 rows) in kernel layout.",
        "#pragma once",
        "",
    ]
    wt = window_table().reshape(-1)
    lines.append(f"#define WFT_WINDOW_TABLE_LEN {wt.size}")
    lines.append("#define WFT_WINDOW_TABLE_INIT { \\")
    for r in range(0, wt.size, 8):
        lines.append("  " + ", ".join(flit(v) for v in wt[r : r + 8]) + ", \\")
    lines.append("}")
    tt = twiddle_table().reshape(-1)
    lines.append(f"#define WFT_TWIDDLE_ROW {TWIDDLE_ROW}")
    lines.append(f"#define WFT_TWIDDLE_TABLE_LEN {tt.size}")
    lines.append("#define WFT_TWIDDLE_TABLE_INIT { \\")
    for r in range(0, tt.size, 8):
        lines.append("  " + ", ".join(flit(v) for v in tt[r : r + 8]) + ", \\")
    lines.append("}")
    lines.append("")
    lines.append(f"#define WFT_MEL_GROUPS {N_GROUPS}")
    lines.append(f"#define WFT_MEL_CLASSES {len(MEL_CLASSES)}")
    for n_mels in (80, 128):
        bank = mb.slaney_mel_bank(n_mels)
        assert not bank[:, 0].any() and not bank[:, 200].any()
        words, index, _ = row_program(bank)
        simulate(bank, words, index)
        lines.append(f"#define WFT_MEL{n_mels}_PROG_VEC {len(words) // 4}")
        lines.append(f"#define WFT_MEL{n_mels}_PROG_INIT {{ \\")
        for r in range(0, len(words), 4):
            lines.append("  {" + ", ".join(words[r : r + 4]) + "}, \\")
        lines.append("}")
        lines.append(f"// per warp: (first_uint4, n_rows) for tap classes {MEL_CLASSES}, n_mels={n_mels}")
        lines.append(f"#define WFT_MEL{n_mels}_INDEX_INIT {{ \\")
        for idx_g in index:
            lines.append("  {" + ", ".join("{%d, %d}" % fc for fc in idx_g) + "}, \\")
        lines.append("}")
        lines.append("")
    return "\n".join(lines) + "\n"


def main(path=None):
    path = path or os.path.join(HERE, "wft_tables.inc")
    text = generate()
    old = None
    if os.path.exists(path):
        with open(path) as f:
            old = f.read()
    if old != text:
        with open(path, "w") as f:
            f.write(text)
    return path


if __name__ == "__main__":
    print(main(sys.argv[1] if len(sys.argv) > 1 else None))
