#!/usr/bin/env python
"""Section 3 of profiles/rNN_ncu_summary.md: the augmentation epilogue (wft::augment_staged_kernel) from one `ncu --set full`
capture of `tools/aug_harness 64 0 6` (run in the build container on the .ncu-rep a gpurun call left in gpurun_out/).

    python profiles/make_epilogue_summary.py gpurun_out/aug_prof.ncu-rep >> profiles/rNN_ncu_summary.md
"""
import csv
import subprocess
import sys
from collections import Counter


def ncu_csv(rep, page):
    txt = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(txt.splitlines()))


def main():
    rep = sys.argv[1]
    cells = float(sys.argv[2]) if len(sys.argv) > 2 else 64 * 128 * 3000
    raw = ncu_csv(rep, "raw")
    d, u = dict(zip(raw[0], raw[2])), dict(zip(raw[0], raw[1]))
    f = lambda k: float(d[k])
    out = ["", "## 3. The augmentation epilogue: `wft::augment_staged_kernel<false, 0>`, one B=64 launch (`tools/aug_harness 64 0 6`, "
           "`ncu --set full --import-source on`)", "",
           "| metric | value |", "|---|---|"]
    rows = [("duration (under ncu)", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
            ("registers / thread", "launch__registers_per_thread"), ("CTAs/SM (shared-memory limit)", "launch__occupancy_limit_shared_mem"),
            ("warp instructions executed", "smsp__inst_executed.sum"), ("IPC (of 4)", "sm__inst_executed.avg.per_cycle_elapsed"),
            ("dram__bytes_read.sum", "dram__bytes_read.sum"), ("dram__bytes_write.sum", "dram__bytes_write.sum"),
            ("DRAM throughput, % of peak over the launch", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            ("L2 hit rate", "lts__t_sector_hit_rate.pct"), ("local loads (spills)", "smsp__inst_executed_op_local_ld.sum")]
    for name, k in rows:
        if k in d:
            out.append(f"| {name} | {d[k]} {u.get(k, '')} |")
    inst = f("smsp__inst_executed.sum")
    out += ["", f"{inst / 1e6:.1f} M warp-instructions for {cells / 1e6:.1f} M cells = {32 * inst / cells:.1f} thread-instructions per cell "
            "(the generic kernel it replaces on aligned shapes: 28.0 M = 36 per cell, of which three quarters were 64-bit address "
            "arithmetic for 4-byte loads; the first staged cut, 512 frames x 16 rows with one thread issuing all bulk copies and an "
            "mbarrier test per row: 37.0 M).", ""]
    stalls = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            try:
                stalls.append((float(v), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    out += ["### Warp stall reasons (cycles per issued instruction)", "", "| reason | cycles/issue |", "|---|---|"]
    for v, k in sorted(stalls, reverse=True)[:8]:
        out.append(f"| {k} | {v:.2f} |")
    src = ncu_csv(rep, "source")
    hdr = src[1]
    ix = {k: i for i, k in enumerate(hdr)}
    ops, tot = Counter(), 0
    for r in src[2:]:
        try:
            n = int(r[ix["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        t = r[ix["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += n
        tot += n
    out += ["", "### Executed warp-instructions by opcode", "", "| opcode | share |", "|---|---|"]
    for k, v in ops.most_common(12):
        out.append(f"| {k} | {100 * v / tot:.1f} % |")
    print("\n".join(out))


if __name__ == "__main__":
    main()
