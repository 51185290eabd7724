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


N_THREADS = 160
N_WARPS = 5
TILE_FRAMES = 16
P_STRIDE = 20  # floats per bin row of the power tile [bin][frame]; stride/4 odd -> consecutive bins hit distinct 16-byte bank groups


def row_support(bank: np.ndarray):
    supp = []
    for m in range(bank.shape[0]):
        nz = np.nonzero(bank[m])[0]
        assert len(nz) > 0 and nz[-1] - nz[0] + 1 == len(nz), "mel row support must be one contiguous band"
        assert nz[0] >= 1 and nz[-1] <= 199, "bins 0 and 200 must carry no weight"
        supp.append((int(nz[0]), int(nz[-1])))
    return supp


# ---- mel phase plan -------------------------------------------------------------------------------------------------
# A thread owns ONE mel row for 8 consecutive frames of the 16-frame tile: lane l of a warp plays row slot l % 16 and frame
# half l / 16, so lanes l and l + 16 write the two 32-byte halves of the SAME 64-byte row segment and a warp-wide 32-byte
# store touches 16 lines instead of 32 (the L1 charges a store per line it touches: ncu counted 16 wavefronts for a store
# whose lanes sit on 32 lines and 8 for one on 16).  The rows are sorted by tap count and cut into groups of 16; a warp
# takes one or two groups (its "passes"), each with its own fully unrolled tap loop of T = (largest tap count in the group)
# taps.  The assignment of groups to warps minimises the slowest warp (every phase ends on a CTA barrier).
def _group_cost(taps):
    """Issue-slot estimate of one pass of one thread: per tap 2 LDS.128 + 4 FFMA2 + 1 weight load; per frame clamp, lg2, fma and
    its share of the running max / min; the store and fixed overhead."""
    return taps * 7.0 + 4.0 * 8 + 24


def _conflict_cost(lanes):
    """lanes = [(start_bin, f0)] of one warp in lane order.  An LDS.128 is served a quarter-warp (8 lanes) at a time;
    two lanes of a quarter collide when they touch different addresses in the same 16-byte bank group.  The power tile is
    [bin][frame] with a row stride of 20 floats, so the bank group of (bin, frame f) is (5 * bin + f / 4) mod 8."""
    cost = 0
    for g in range(0, len(lanes), 8):
        by_group = {}
        for s, f0 in set(lanes[g : g + 8]):
            key = (5 * s + f0 // 4) % 8
            by_group[key] = by_group.get(key, 0) + 1
        cost += max(by_group.values()) - 1
    return cost


def _lanes_of(order, start, nf):
    """lane -> (row, f0) for one pass: 16 rows, lane l plays row slot l % 16 and frames 8 * (l / 16) .."""
    n = len(order)
    assert n == 16 and nf == 8
    return [(order[lane % n], nf * (lane // n)) for lane in range(32)]


def _order_rows(rows, start_range, nf):
    """Choose every row's first tap bin (anywhere its padded window still covers the support) and permute the rows over
    the lanes so that the 8 lanes of a quarter-warp start on distinct 16-byte bank groups.

    -> (lane order, {row: start bin}, residual conflicts)."""
    import random

    rng = random.Random(1234)
    n = len(rows)

    def cost_of(order, start):
        return _conflict_cost([(start[m], f0) for m, f0 in _lanes_of(order, start, nf)])

    best = None
    for _restart in range(8):
        order = list(rows)
        rng.shuffle(order) if _restart else None
        start = {m: start_range[m][1] for m in rows}
        cost = cost_of(order, start)
        for _ in range(6000):
            if cost == 0:
                break
            if rng.random() < 0.5 and n > 1:
                i, j = rng.randrange(n), rng.randrange(n)
                if i == j:
                    continue
                order[i], order[j] = order[j], order[i]
                c = cost_of(order, start)
                if c <= cost:
                    cost = c
                else:
                    order[i], order[j] = order[j], order[i]
            else:
                m = order[rng.randrange(n)]
                lo, hi = start_range[m]
                if lo == hi:
                    continue
                old = start[m]
                start[m] = rng.randint(lo, hi)
                c = cost_of(order, start)
                if c <= cost:
                    cost = c
                else:
                    start[m] = old
        if best is None or cost < best[2]:
            best = (list(order), dict(start), cost)
        if cost == 0:
            break
    return best


def mel_plan(bank: np.ndarray):
    """Thread-level plan of the mel phase.

    -> dict(weights=[float32...], thread=[[(row, start_bin) per pass] * 160], warp_t=[[taps per pass] per warp],
            warp_wbase=[weight offset of warp w], conflicts=int)
    Weight of (warp w, pass p, tap j, row slot i) = weights[wbase[w] + (sum of the taps of the earlier passes + j) * 16 + i].
    """
    import itertools

    n_mels = bank.shape[0]
    assert n_mels % 16 == 0
    supp = row_support(bank)
    ntaps = [supp[m][1] - supp[m][0] + 1 for m in range(n_mels)]
    by_taps = sorted(range(n_mels), key=lambda m: (ntaps[m], m))
    groups = [by_taps[g : g + 16] for g in range(0, n_mels, 16)]
    gt = [max(ntaps[m] for m in rows) for rows in groups]
    n_groups = len(groups)
    assert N_WARPS <= n_groups <= 2 * N_WARPS
    # choose which groups share a warp: n_groups - N_WARPS warps take two groups
    best = None
    n_double = n_groups - N_WARPS
    idx = list(range(n_groups))
    for pairs in itertools.combinations(itertools.combinations(idx, 2), n_double):
        used = [g for pr in pairs for g in pr]
        if len(set(used)) != len(used):
            continue
        singles = [g for g in idx if g not in used]
        assign = [list(pr) for pr in pairs] + [[g] for g in singles]
        costs = [sum(_group_cost(gt[g]) for g in a) for a in assign]
        key = (max(costs), sum(c * c for c in costs))
        if best is None or key < best[0]:
            best = (key, assign)
    _, assign = best
    assign.sort(key=lambda a: (-len(a), a))   # two-pass warps first (warp-uniform dispatch in the kernel)
    weights, thread, warp_t, warp_wbase = [], [], [], []
    conflicts = 0
    for a in assign:
        warp_wbase.append(len(weights))
        lanes = [[] for _ in range(32)]
        taps_of = []
        for g in a:
            rows, taps = groups[g], gt[g]
            # the padded window [start, start + taps) must cover the support and stay on bins 1..199 (rewritten every
            # tile, always finite); whatever freedom is left goes into avoiding bank conflicts
            # (a group whose rows leave no freedom may not be arrangeable without conflicts: one padded tap more buys it)
            cands = []
            for extra in (0, 1, 2):
                t_try = taps + extra
                start_range = {m: (max(1, supp[m][1] - t_try + 1), min(supp[m][0], 200 - t_try)) for m in rows}
                assert all(lo <= hi for lo, hi in start_range.values())
                order, start_of, cost = _order_rows(rows, start_range, 8)
                # wavefronts per tile: every tap is 2 LDS.128 of 4 wavefronts, a conflict adds one to each of them
                cands.append((2 * t_try * (4 + cost), t_try, order, start_of, cost))
                if cost == 0:
                    break
            _, taps, order, start_of, cost = min(cands, key=lambda c: c[0])
            conflicts += cost
            for j in range(taps):
                for m in order:
                    # 0.25 * bank weight: the packed two-frame FFT yields 4 |X|^2 (exact scaling)
                    weights.append(np.float32(bank[m, start_of[m] + j]) * np.float32(0.25))
            for lane in range(32):
                m = order[lane % 16]
                lanes[lane].append((m, start_of[m]))
            taps_of.append(taps)
        thread.extend(lanes)
        warp_t.append(taps_of)
    assert len(thread) == N_THREADS
    return dict(weights=weights, thread=thread, warp_t=warp_t, warp_wbase=warp_wbase, conflicts=conflicts)


def simulate(bank: np.ndarray, plan):
    """Replay the plan on a random power tile and compare with the dense product (float64); every cell exactly once."""
    rng = np.random.default_rng(0)
    P = rng.random((201, TILE_FRAMES))
    out = np.full((bank.shape[0], TILE_FRAMES), np.nan)
    w = np.asarray(plan["weights"], dtype=np.float64)
    for t, passes in enumerate(plan["thread"]):
        wi, lane = t // 32, t % 32
        f0, f1 = 8 * (lane // 16), 8 * (lane // 16) + 8
        ofs = plan["warp_wbase"][wi] + lane % 16
        for (row, start), taps in zip(passes, plan["warp_t"][wi]):
            assert 1 <= start and start + taps - 1 <= 199 and row < 128 and start < 256
            acc = np.zeros(8)
            for j in range(taps):
                acc += w[ofs + j * 16] * P[start + j, f0:f1]
            ofs += taps * 16
            assert np.isnan(out[row, f0:f1]).all(), "cell computed twice"
            out[row, f0:f1] = acc
    assert not np.isnan(out).any(), "cell never computed"
    ref = (bank.astype(np.float64) * 0.25) @ P
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
    lines.append(f"#define WFT_MEL_P_STRIDE {P_STRIDE}")
    for n_mels in (80, 128):
        bank = mb.slaney_mel_bank(n_mels)
        assert not bank[:, 0].any() and not bank[:, 200].any()
        plan = mel_plan(bank)
        simulate(bank, plan)
        w = list(plan["weights"])
        while len(w) % 4:
            w.append(np.float32(0.0))
        assert len(w) < 4096
        lines.append(f"// n_mels={n_mels}: taps per warp and pass = {plan['warp_t']}, residual LDS.128 conflicts = {plan['conflicts']}")
        lines.append(f"#define WFT_MEL{n_mels}_WARP_PASSES {{" + ", ".join(str(len(v)) for v in plan["warp_t"]) + "}")
        lines.append(f"#define WFT_MEL{n_mels}_WARP_T0 {{" + ", ".join(str(v[0]) for v in plan["warp_t"]) + "}")
        lines.append(f"#define WFT_MEL{n_mels}_WARP_T1 {{" + ", ".join(str(v[1] if len(v) > 1 else 0) for v in plan["warp_t"]) + "}")
        lines.append(f"#define WFT_MEL{n_mels}_WARP_WBASE {{" + ", ".join(str(v) for v in plan["warp_wbase"]) + "}")
        lines.append(f"#define WFT_MEL{n_mels}_W_LEN {len(w)}")
        lines.append(f"#define WFT_MEL{n_mels}_W_INIT {{ \\")
        for r in range(0, len(w), 8):
            lines.append("  " + ", ".join(flit(v) for v in w[r : r + 8]) + ", \\")
        lines.append("}")
        lines.append(f"// per thread: row0 | start_bin0 << 7 | row1 << 15 | start_bin1 << 22   (pass 0 / pass 1)")
        lines.append(f"#define WFT_MEL{n_mels}_THREAD_INIT {{ \\")
        words = []
        for passes in plan["thread"]:
            (r0, s0) = passes[0]
            (r1, s1) = passes[1] if len(passes) > 1 else (0, 1)
            words.append("0x%08xu" % (r0 | (s0 << 7) | (r1 << 15) | (s1 << 22)))
        for r in range(0, len(words), 8):
            lines.append("  " + ", ".join(words[r : r + 8]) + ", \\")
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
