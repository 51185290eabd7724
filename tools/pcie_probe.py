#!/usr/bin/env python
"""Host <-> device copy ceiling of the box, per rank and in aggregate, with NOTHING else running: N ranks (torchrun) each move
the byte counts of one bench step between pinned host memory and their GPU -- H2D only, D2H only, both at once.

    python tools/pcie_probe.py                                                    # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py

This is the measurement behind the flat end-to-end curve of SCALE_r01.json: if 8 ranks copying concurrently
reach no more aggregate GB/s than 2 do, the e2e number is bounded by the host (memory / PCIe root complexes), not by the
front end.  ``--numa`` additionally pins each rank's host buffers and thread to the NUMA node of its GPU (when /sys exposes
it) to see whether placement moves the ceiling."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, N_SAMPLES, N_MELS, N_FRAMES = 64, 480000, 128, 3000


def numa_node_of_gpu(index):
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
        return int(open(path).read())
    except Exception:  # noqa: BLE001
        return -1


def pin_to_node(node):
    try:
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return len(ids)
    except Exception:  # noqa: BLE001
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numa", action="store_true")
    ap.add_argument("--seconds", type=float, default=1.0)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    node = numa_node_of_gpu(local)
    pinned_cpus = pin_to_node(node) if (args.numa and node >= 0) else 0   # first-touch then places the pinned pages on that node
    shapes = {"h2d_f32_pcm": (B * N_SAMPLES * 4, "h2d"), "h2d_i16_pcm": (B * N_SAMPLES * 2, "h2d"),
              "d2h_features": (B * N_MELS * N_FRAMES * 4, "d2h"), "both_f32_pcm_and_features": (0, "both")}
    host_in = torch.empty(B * N_SAMPLES * 4, dtype=torch.uint8).pin_memory()
    host_out = torch.empty(B * N_MELS * N_FRAMES * 4, dtype=torch.uint8).pin_memory()
    host_in.fill_(1)
    host_out.fill_(1)
    dev_in = torch.empty_like(host_in, device=dev)
    dev_out = torch.empty_like(host_out, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    result = {}
    for name, (nbytes, kind) in shapes.items():
        def once():
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    n = nbytes or host_in.numel()
                    dev_in[:n].copy_(host_in[:n], non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    host_out.copy_(dev_out, non_blocking=True)
        for _ in range(3):
            once()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < args.seconds:
            for _ in range(4):
                once()
            torch.cuda.synchronize()
            reps += 4
        el = time.perf_counter() - t0
        moved = reps * ((nbytes or host_in.numel()) if kind != "d2h" else 0) + reps * (host_out.numel() if kind in ("d2h", "both") else 0)
        t = torch.tensor([moved / el / 1e9], dtype=torch.float64, device=dev)
        lo = t.clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        result[name] = {"aggregate_GBps": float(t.item()), "slowest_rank_GBps": float(lo.item()), "bytes_per_copy": nbytes or None}
    nodes = [None] * world
    if world > 1:
        dist.all_gather_object(nodes, node)
    else:
        nodes = [node]
    if rank == 0:
        line = {"n_gpus": world, "numa_pinning": bool(args.numa), "gpu_numa_nodes": nodes, "cpus_of_rank0_node": pinned_cpus,
                "host_cpus": os.cpu_count(), "copies": result}
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"pcie_probe_n{world}{'_numa' if args.numa else ''}.json"), "w") as f:
            json.dump(line, f, indent=1)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
