#!/usr/bin/env python
"""BASELINE.json config 5: batch-size sweep 1 .. 2048 clips per GPU at 128 mel, int16 and float32 PCM, on 1..8 B200, next to
the reference's CPU path on the host cores.

    python tools/sweep.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py     # 8 GPUs

Every rank runs the same sweep on its own clips (batch shards, no data-path collective); a point's time is the MAX over
ranks of a CUDA-event-timed block and the reported clips/s is the whole job.  At every B the launches rotate over enough
input / output buffer sets that one pass over them touches more than twice the 126 MB L2 (small batches would otherwise be
L2 numbers, not HBM numbers -- VERDICT r1 #5).  Rank 0 prints one JSON object per point and writes the list to
``gpurun_out/sweep_r2_n<world>.json``; with ``--cpu`` it also times the oracle port in the reference's deployment shapes
(bench.py) so that the sweep carries its CPU baseline."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import whisper_finetune_b200 as w  # noqa: E402

L2_BYTES = 126e6
N_MELS = 128


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--max-batch", type=int, default=2048)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak = 6551.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    res = []
    # consecutive launches work on different buffer sets wherever more than one set is rotated: declared independent batches
    # (WFT_LAUNCH_OVERLAP); with a single set (B >= 128) the library sees the shared output buffer and launches the ordinary way
    w.set_overlap(True)
    g = torch.Generator().manual_seed(1000 + rank)
    base = (0.1 * torch.randn(64, 480000, generator=g)).clamp_(-1, 1).to(dev)
    for dtype in (torch.float32, torch.int16):
        B = 1
        while B <= args.max_batch:
            per_set = B * (480000 * (4 if dtype == torch.float32 else 2) + N_MELS * 3000 * 4)
            n_sets = max(1, min(64, int(2 * L2_BYTES // per_set) + 1))
            pcm_sets = []
            for s in range(n_sets):
                idx = (torch.arange(B, device=dev) + 7 * s) % 64
                x = base[idx].roll(997 * s, dims=1)
                pcm_sets.append((x * 32767).round().to(torch.int16) if dtype == torch.int16 else x.contiguous())
            out_sets = [torch.empty(B, N_MELS, 3000, device=dev) for _ in range(n_sets)]
            masks = w.draw_mask_params(42, rank * B, B, N_MELS, 3000, 100, 43, 1.0, dev)
            for s in range(min(n_sets, 3)):
                w.frontend_forward(pcm_sets[s], N_MELS, mask_params=masks, out=out_sets[s])
            torch.cuda.synchronize()
            n = max(n_sets, min(200, max(8, 8192 // B)))
            times = []
            for _ in range(3):
                if world > 1:
                    dist.barrier(device_ids=[local])
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(n):
                    w.frontend_forward(pcm_sets[i % n_sets], N_MELS, mask_params=masks, out=out_sets[i % n_sets])
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times.append(float(t.item()))
            ms = sorted(times)[1]
            r = dict(n_gpus=world, dtype=str(dtype).split(".")[-1], n_mels=N_MELS, clips_per_gpu=B, buffer_sets=n_sets,
                     bytes_rotated=n_sets * per_set, us_per_launch=ms * 1e3, clips_per_s=world * B / ms * 1e3,
                     overlapped_batches=n_sets > 1, us_per_clip_per_gpu=ms * 1e3 / B, algorithmic_GBps_per_gpu=per_set / ms / 1e6,
                     frac_of_measured_hbm_peak=per_set / ms / 1e6 / peak)
            res.append(r)
            if rank == 0:
                print(json.dumps(r), flush=True)
            del pcm_sets, out_sets
            torch.cuda.empty_cache()
            B *= 2
    if rank == 0:
        out = {"points": res}
        if args.cpu:
            import bench

            best, shapes = bench.cpu_reference_shapes(10, 1, budget_s=8.0)
            out["cpu_baseline"] = {"best_shape": best, "shapes": shapes, "note": "clips/s of the CPU path does not depend on the batch size (clip by clip)"}
            print(json.dumps(out["cpu_baseline"]), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_r2_n{world}.json"), "w"), indent=1)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
