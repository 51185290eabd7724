#!/usr/bin/env python
"""bench.py -- throughput of the Whisper audio front end (30-s clips/s, log-mel + SpecAugment masks).

    python bench.py --gpus 1 --steps 50 --warmup 5            # this repo's CUDA path, one JSON line
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # one rank per GPU, each on its own shard

Workload (BASELINE.json configs[1] with the metric's SpecAugment masks on): n_mels=128, 64 synthetic 30-s clips of
float32 PCM per GPU per step.  A "step" is one pass of the hot path over one batch:

    mask draw (Philox, device) -> fused kernel: zero pad, STFT, |X|^2, mel, log10, per-clip max-8 floor, (x+4)/4,
    time + frequency masks -> x[64, 128, 3000]

`value`  : clips/s with the PCM already resident in HBM (CUDA events, max over ranks, whole job).
`e2e`    : clips/s through the public API with HOST buffers: pinned PCM -> H2D -> kernels -> D2H pinned features.
`roofline`: algorithmic bytes of one fused-kernel launch / its mean launch duration, against the measured HBM peak.
`cpu_baseline`: the oracle (torch.stft recipe + masks, clip by clip like data_loader.py:273-292) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_MELS = 128
BATCH = 64
N_SAMPLES = 480000
N_FRAMES = 3000
TIME_MASK = 100
FREQ_MASK = 43
SEED = 42
N_STREAMS = 2   # batches in flight in the throughput loop (the roofline loop stays strictly one launch after another)
BYTES_PER_CLIP = 4 * N_SAMPLES + 4 * N_MELS * N_FRAMES  # 3 456 000 (algorithmic bytes, f32 in, 128 mel)
METRIC = "30-s clips/sec log-mel+SpecAugment"
UNIT = "clips/s"
# dram__bytes_read.sum + dram__bytes_write.sum of one B=64 launch, from profiles/r01_ncu_summary.md
NCU_DRAM_BYTES_PER_LAUNCH = 182.6e6


def workload_config(n_gpus):
    return {
        "workload": "large-v3 front end: n_mels=128, 64 x 30 s float32 PCM per GPU per step, log-mel + SpecAugment "
                    "time/freq masks (T=100, F=43, p=1.0), device-resident input",
        "n_mels": N_MELS, "clips_per_gpu_per_step": BATCH, "global_batch": BATCH * n_gpus, "pcm_dtype": "float32",
        "spec_augment": True, "sharding": "DistributedSampler-style batch shards, no data-path collective",
        "cache": "4 rotating input/output buffer sets (885 MB per GPU) > 126 MB L2",
        "streams": N_STREAMS,
        "in_flight": "value: step i on stream i % 2 (two batches in flight); roofline: one launch after another on one stream",
    }


def synth_pcm(n_clips, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    return (0.1 * torch.randn(n_clips, N_SAMPLES, generator=g)).clamp_(-1, 1)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every ~2 ms from a thread (the timed region is
    only milliseconds long, so spawning nvidia-smi per sample would miss it); falls back to nvidia-smi if NVML is absent."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.sm, self.bits = [], 0
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.sm.append(float(self._nvml.nvmlDeviceGetClockInfo(self._h, self._nvml.NVML_CLOCK_SM)))
                    self.bits |= int(self._nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    self._stop.wait(0.002)
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                          "clocks_event_reasons.active", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.sm.append(float(out[0]))
                    self.sm_max = float(out[1])
                    self.bits |= int(out[2].strip(), 16)
                    self._stop.wait(0.05)
            except Exception:
                self._stop.wait(0.05)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.sm_max,
                "reasons": [name for bit, name in self.REASONS.items() if self.bits & bit], "samples": len(self.sm),
                "sampler": "nvml" if self._nvml is not None else "nvidia-smi"}


def cpu_reference_clips_per_s(min_seconds, max_clips=None, threads=None):
    """The reference's CPU path on this box's host cores: oracle log-mel + masks, clip by clip."""
    import torch

    from oracle import pipeline as OP
    from oracle import specaug as OS

    if threads:
        torch.set_num_threads(threads)
    n = BATCH  # the same 64-clip batch the CUDA arm runs per step
    pcm = synth_pcm(n, SEED)
    masks = OS.draw_mask_params(SEED, 0, n, N_MELS, N_FRAMES, TIME_MASK, FREQ_MASK, 1.0)
    OP.front_end_batch(pcm[:2], N_MELS, masks=masks[:2])  # warm-up (FFT plans, filter cache)
    done, t0 = 0, time.perf_counter()
    while True:
        OP.front_end_batch(pcm, N_MELS, masks=masks)
        done += n
        el = time.perf_counter() - t0
        if el >= min_seconds or (max_clips and done >= max_clips):
            break
    return done / el, done, el, torch.get_num_threads()


_RESULT_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (here: its restatement in oracle/, since
    whisper.audio is a third-party dependency that is not installable offline) on all host threads.  A step is a
    bounded sample of the workload: one 64-clip batch (the same batch size the CUDA arm runs per step)."""
    import torch

    from oracle import pipeline as OP
    from oracle import specaug as OS

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = BATCH
    pcm = synth_pcm(n, SEED)
    masks = OS.draw_mask_params(SEED, 0, n, N_MELS, N_FRAMES, TIME_MASK, FREQ_MASK, 1.0)
    OP.front_end_batch(pcm[:2], N_MELS, masks=masks[:2])
    for _ in range(args.warmup):
        OP.front_end_batch(pcm, N_MELS, masks=masks)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        OP.front_end_batch(pcm, N_MELS, masks=masks)
    el = time.perf_counter() - t0
    val = n * args.steps / el
    thr = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": thr, "kind": "port",
                         "sample": f"{n} clips per step ({n * args.steps} clips in {el:.1f} s), torch intra-op threads={thr} "
                                   f"of {cores} host cores, oracle/pipeline.py (torch.stft recipe + masks, clip by clip)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "audio_hours_per_s": val * 30.0 / 3600.0,
    }
    emit(line)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    import whisper_finetune_b200 as wft

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = wft._lib.load()

    fe = wft.FrontEnd(n_mels=N_MELS, device=dev, spec_augment=True,
                      spec_augment_params={"time_mask_param": TIME_MASK, "freq_mask_param": FREQ_MASK, "p": 1.0},
                      seed=SEED)
    n_sets = 4
    # every rank owns its shard of the synthetic "dataset": clip ids rank*B + step*world*B ... (weak scaling)
    pcm_sets = [synth_pcm(BATCH, SEED + 1000 * rank + s).to(dev) for s in range(n_sets)]
    out_sets = [torch.empty(BATCH, N_MELS, N_FRAMES, device=dev) for _ in range(n_sets)]
    # two batches in flight (a loader with one batch of prefetch): step i runs on stream i % 2, so the ramp-down of one
    # launch (CTAs running out of tiles) is filled by the next batch's CTAs.  Buffer set i % 4 is only ever used on
    # stream i % 2, so launches that share buffers stay ordered.
    streams = [torch.cuda.Stream(device=dev) for _ in range(N_STREAMS)]

    def step(i):
        s = i % n_sets
        with torch.cuda.stream(streams[i % N_STREAMS]):
            return fe(pcm_sets[s], clip_offset=(i * world + rank) * BATCH, out=out_sets[s])

    def fork():   # side streams start behind everything already queued on the current stream
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)

    def join():
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    fork()
    for i in range(max(args.warmup, 3)):
        step(i)
    join()
    barrier()
    lib.wft_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        fork()
        for i in range(args.steps):
            step(i)
        join()
        ev1.record()
        barrier()
        launches = int(lib.wft_launch_count(0))
        # the timed region lasts only milliseconds: keep the very same steps running for another ~0.4 s so that the
        # clock / throttle-reason samples describe the GPU under this load (not part of the timing)
        t_end = time.perf_counter() + 0.4
        i = args.steps
        while time.perf_counter() < t_end:
            fork()
            for _ in range(20):
                step(i)
                i += 1
            join()
            torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * BATCH * args.steps / (ms_max * 1e-3)

    # fused kernel alone (no mask draw), for the roofline: K back-to-back launches on the current stream
    masks = wft.draw_mask_params(SEED, 0, BATCH, N_MELS, N_FRAMES, TIME_MASK, FREQ_MASK, 1.0, dev)
    for i in range(3):
        wft.frontend_forward(pcm_sets[i % n_sets], N_MELS, mask_params=masks, out=out_sets[i % n_sets])
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(args.steps):
        wft.frontend_forward(pcm_sets[i % n_sets], N_MELS, mask_params=masks, out=out_sets[i % n_sets])
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps

    # end to end through the public API with host buffers (pinned): H2D + kernels + D2H every step
    def run_e2e(pcm_dtype, readback):
        host_pcm = [synth_pcm(BATCH, SEED + 1000 * rank + s) for s in range(2)]
        if pcm_dtype == torch.int16:
            host_pcm = [(x * 32767).round().to(torch.int16) for x in host_pcm]
        host_pcm = [x.pin_memory() for x in host_pcm]
        shape = (BATCH, N_MELS, N_FRAMES) if readback == "features" else (BATCH,)
        host_out = [torch.empty(shape).pin_memory() for _ in range(2)]
        pipe = wft.HostPipeline(fe, BATCH, pcm_dtype=pcm_dtype, n_chunks=4, n_streams=2, readback=readback)
        steps = max(3, min(args.steps, 20))
        for i in range(2):
            pipe(host_pcm[i % 2], host_out[i % 2], clip_offset=i * BATCH)
        pipe.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            pipe(host_pcm[i % 2], host_out[i % 2], clip_offset=(i * world + rank) * BATCH)
        pipe.join()   # the current stream (and so e1) waits for the last D2H copy
        e1.record()
        pipe.synchronize()
        barrier()
        probe = float(host_out[(steps - 1) % 2].reshape(BATCH, -1)[0, :8].sum())  # the result really is on the host
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * BATCH * steps / (float(te.item()) * 1e-3), pipe.h2d_bytes, pipe.d2h_bytes, steps, probe

    e2e_value, e2e_h2d, e2e_d2h, e2e_steps, checksum = run_e2e(torch.float32, "features")
    # informational variants (not the headline): int16 PCM halves the H2D bytes; a training step consumes the features
    # on the device, so only a per-clip probe has to return
    e2e_variants = {}
    for name, dt, rb in (("int16_pcm_features_to_host", torch.int16, "features"),
                         ("f32_pcm_features_stay_on_device", torch.float32, "probe"),
                         ("int16_pcm_features_stay_on_device", torch.int16, "probe")):
        v, h2d, d2h, _, _ = run_e2e(dt, rb)
        e2e_variants[name] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        achieved = BATCH * BYTES_PER_CLIP / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "audio_hours_per_s": value * 30.0 / 3600.0,
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                    "steps": e2e_steps, "pipeline": "4 chunks on 2 streams, consecutive batches overlap, pinned host buffers, float32 PCM in, full "
                    "float32 features back to the host", "checksum": checksum},
            "e2e_variants": e2e_variants,
            "gpu_launches": launches,
            "gpu_launches_per_step": launches / max(args.steps, 1),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "peak_source": peak_src,
                         "kernel": "wft::frontend_kernel<128,float>", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": BATCH * BYTES_PER_CLIP},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, done, el, thr = cpu_reference_clips_per_s(12.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": thr, "kind": "port",
                                    "sample": f"{done} clips in {el:.1f} s, oracle/pipeline.py clip by clip, torch "
                                              f"intra-op threads={thr} of {os.cpu_count()} host cores"}
        emit(line)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps < 1:
        ap.error("--steps must be >= 1")
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl != "reference":
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    # stdout carries exactly ONE JSON line: anything a library prints on file descriptor 1 while the run is going on (NCCL's
    # version banner, for one) is sent to stderr instead, and the line is written to the real stdout at the end
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
