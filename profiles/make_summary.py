#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into profiles/rNN_ncu_summary.md (run in the build container)
and record the DRAM traffic of the captured launch in profiles/ncu_traffic.json (what bench.py reports as roofline.traffic).

    python profiles/make_summary.py gpurun_out/launches.csv gpurun_out/prof.ncu-rep [n_tiles_per_launch] [round] > profiles/rNN_ncu_summary.md
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict


def launch_list(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ix = {k: i for i, k in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) != len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0][-64:]
        agg[name][0] += 1
        agg[name][1] += float(r[ix["Metric Value"]])
    return agg


def ncu_csv(rep, page):
    txt = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(txt.splitlines()))


def main():
    launches, rep = sys.argv[1], sys.argv[2]
    n_tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 6016.0
    rnd = sys.argv[4] if len(sys.argv) > 4 else "1"
    out = [f"# Round {rnd} ncu evidence — `wft::frontend_kernel<128,float>` on B200 (sm_100a)", "",
           "All captures ran under `gpurun` on one B200 with `--clock-control none`. Times under ncu are cold-cache and "
           "serialised: they are evidence of SHARES and counters, never bench values.", ""]
    agg = launch_list(launches)
    tot = sum(v[1] for v in agg.values())
    out += ["## 1. Launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`",
            "`ncu --metrics gpu__time_duration.sum --clock-control none --csv` (all launches of the process: warm-up, timed "
            "steps, kernel-only loop, host pipeline).", "", "| kernel | launches | total µs | share of GPU time |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f} % |")
    out += ["", "A timed step is exactly two launches of this library: `frontend_kernel` and its `fixup_kernel` (the SpecAugment "
            "intervals are drawn inside the front-end kernel); synthetic PCM is generated on the host before timing, so no torch "
            "kernel runs inside the step.  `augment_kernel` / the draw kernels belong to the `value_with_time_warp` and epilogue loops.", ""]

    raw = ncu_csv(rep, "raw")
    d = dict(zip(raw[0], raw[2]))
    units = dict(zip(raw[0], raw[1]))

    def g(k):
        return d.get(k, "n/a")

    out += ["## 2. Full-set capture of one B=64 launch (`ncu --set full --import-source on`, torch-free harness)", "",
            "| metric | value |", "|---|---|"]
    for k, label in [("gpu__time_duration.sum", "duration (under ncu)"), ("launch__grid_size", "grid (persistent CTAs)"),
                     ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
                     ("launch__occupancy_limit_shared_mem", "CTAs/SM (shared-memory limit)"),
                     ("smsp__inst_executed.sum", "warp instructions executed"),
                     ("sm__inst_executed.avg.per_cycle_active", "IPC (of 4)"),
                     ("smsp__warps_active.avg.per_cycle_active", "active warps / scheduler"),
                     ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler"),
                     ("dram__bytes_read.sum", "dram__bytes_read.sum"), ("dram__bytes_write.sum", "dram__bytes_write.sum"),
                     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
                     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
                     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
                     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
                     ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data pipe (LSU wavefronts), % of peak over the launch"),
                     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
                     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "shared-load bank conflicts"),
                     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "shared-store bank conflicts"),
                     ("smsp__sass_inst_executed_op_local_ld.sum", "local loads (spills)")]:
        if k in d:
            out.append(f"| {label} | {g(k)} {units.get(k, '')} |")
    try:
        rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units["dram__bytes_read.sum"]]
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"), "w") as fh:
            json.dump({"dram_bytes_per_launch": (rd + wr) * scale, "read": rd * scale, "write": wr * scale,
                       "kernel": "wft::frontend_kernel<128,float>, B=64", "source": f"round {rnd}: ncu --set full of tools/harness 64 128, "
                       + os.path.basename(rep)}, fh)
            fh.write("\n")
        out += ["", f"DRAM traffic of the launch = {rd + wr:.1f} {units['dram__bytes_read.sum']} (read {rd:.1f} + write {wr:.1f}) "
                "against 221.2 MB of algorithmic bytes (122.9 MB PCM in + 98.3 MB features out): the input is read once, and "
                "the output is written at most once (the rest of the dirty lines leave L2 after the kernel) — the in-place "
                "fix-up is served from L2.", ""]
    except Exception:
        pass
    stalls = []
    for k in raw[0]:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            try:
                stalls.append((float(d[k]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    out += ["### Warp stall reasons (cycles per issued instruction)", "", "| reason | cycles/issue |", "|---|---|"]
    for v, k in sorted(stalls, reverse=True)[:9]:
        out.append(f"| {k} | {v:.2f} |")

    src = ncu_csv(rep, "source")
    hdr = src[1]
    ix = {k: i for i, k in enumerate(hdr)}
    data = src[2:]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0

    regions, cur = [], {"exec": 0.0, "samp": 0.0, "n": 0, "ops": defaultdict(float)}
    for r in data:
        s = r[ix["Source"]].strip()
        cur["exec"] += f(r, "Instructions Executed")
        cur["samp"] += f(r, "# Samples")
        cur["n"] += 1
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        cur["ops"][op] += f(r, "Instructions Executed")
        if s.startswith("BAR.SYNC") or "RET" in s or s.startswith("EXIT"):
            regions.append(cur)
            cur = {"exec": 0.0, "samp": 0.0, "n": 0, "ops": defaultdict(float)}
    regions.append(cur)
    tot_e = sum(x["exec"] for x in regions)
    tot_s = sum(x["samp"] for x in regions) or 1.0
    # fp32 flops the launch really executed (thread-level, predicated-on): FFMA = 2, packed FFMA2 = 4, FADD2 / FMUL2 = 2,
    # scalar add / mul / min-max / MUFU = 1 -- the numerator of the FMA-peak bound north_star asks to be compared with HBM
    flop_w = {"FFMA": 2, "FFMA2": 4, "FADD2": 2, "FMUL2": 2, "FADD": 1, "FMUL": 1, "FMNMX": 1, "FMNMX3": 2, "MUFU": 1}
    flops = 0.0
    for r in data:
        s = r[ix["Source"]].strip()
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        flops += flop_w.get(op, 0) * f(r, "Predicated-On Thread Instructions Executed")
    try:
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
        t = json.load(open(tpath))
        t["fp32_flops_per_launch"] = flops
        t["warp_instructions_per_launch"] = tot_e
        t["lsu_wavefronts_per_sm"] = float(d.get("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "nan"))
        t["data_pipe_pct_of_peak_under_ncu"] = float(d.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "nan"))
        with open(tpath, "w") as fh:
            json.dump(t, fh)
            fh.write("\n")
        out += ["", f"fp32 operations executed by the launch (FFMA = 2, FFMA2 = 4, other packed = 2, scalar = 1): "
                f"{flops / 1e9:.2f} GFLOP = {flops / 64 / 1e6:.1f} MFLOP per clip; L1 data-pipe wavefronts per SM: "
                f"{t['lsu_wavefronts_per_sm']:.0f} (one per cycle at best)."]
    except Exception as e:  # noqa: BLE001
        out += ["", f"(flop / wavefront record not written: {e})"]
    out += ["", f"### Instructions per barrier-delimited region of the SASS ({len(data)} instructions = {len(data) * 16 / 1024:.0f} KB)",
            "", "Regions in program order: 0-1 prologue / loop top / edge staging, 2 audio gather + window (+ first butterflies), 3 stage A "
            "DFT + twiddles + exchange stores (+ describe_tile), 4 stage B row loads + DFT + mirror shuffles, 5 power tile, 6 mel "
            "phase, 8-9 tile close + publication (red.max, tile minimum), 10+ helper functions (bulk-copy issue, edge mel, draw).", "", "| # | SASS instr | warp-instr / tile | share | stall-sample share | top opcodes (per tile) |", "|---|---|---|---|---|---|"]
    for i, x in enumerate(regions):
        if x["exec"] < 0.002 * tot_e:
            continue
        top = " ".join(f"{k}:{v / n_tiles:.0f}" for k, v in sorted(x["ops"].items(), key=lambda kv: -kv[1])[:6])
        out.append(f"| {i} | {x['n']} | {x['exec'] / n_tiles:.0f} | {100 * x['exec'] / tot_e:.1f} % | {100 * x['samp'] / tot_s:.1f} % | {top} |")
    out += ["", f"Total {tot_e / n_tiles:.0f} warp-instructions per 16-frame tile ({n_tiles:.0f} tiles in the launch)."]
    hot = sorted(data, key=lambda r: -f(r, "# Samples"))[:8]
    out += ["", "### Hottest instructions (warp-state samples)", "", "| share | SASS | dominant stall |", "|---|---|---|"]
    for r in hot:
        st = {k: f(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        k, v = max(st.items(), key=lambda kv: kv[1])
        out.append(f"| {100 * f(r, '# Samples') / tot_s:.1f} % | `{r[ix['Source']].strip()[:70]}` | {k[6:]} |")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    sys.stdout.write(main())
