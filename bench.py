#!/usr/bin/env python
"""bench.py -- throughput of the Whisper audio front end (30-s clips/s, log-mel + SpecAugment masks).

    python bench.py --gpus 1 --steps 50 --warmup 5            # this repo's CUDA path, one JSON line
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # one rank per GPU, each on its own shard

Workload (BASELINE.json configs[1] with the metric's SpecAugment masks on): n_mels=128, 64 synthetic 30-s clips of
float32 PCM per GPU per step.  A "step" is one pass of the hot path over one batch:

    front-end grid: zero pad, STFT, |X|^2, mel, log10, (x+4)/4, time + frequency masks (intervals drawn in the kernel: Philox)
    -> fix-up grid: per-clip max-8 floor where it binds, min-value pad -> x[64, 128, 3000]

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
TIME_WARP_W = 80
SEED = 42
N_STREAMS = 1   # one stream; consecutive batches are declared independent (WFT_LAUNCH_OVERLAP), so batch i + 1 fills the SM slots
                # batch i frees as it runs out of tiles
BYTES_PER_CLIP = 4 * N_SAMPLES + 4 * N_MELS * N_FRAMES  # 3 456 000 (algorithmic bytes, f32 in, 128 mel)
METRIC = "30-s clips/sec log-mel+SpecAugment"
UNIT = "clips/s"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one B=64 launch of the fused kernel, from the committed summary of
    the latest `ncu --set full` capture (profiles/ncu_traffic.json, written by profiles/make_summary.py); (None, why) if
    there is none.  ncu cannot run inside the timed bench, so this is a recorded measurement, not a live one."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_launch"]), f"profiles/ncu_traffic.json ({d.get('source', '?')})"
    except Exception as e:  # noqa: BLE001
        return None, f"no committed ncu capture ({type(e).__name__})"


def other_bounds(kernel_ms, sm_count, sm_mhz):
    """The bounds next to HBM that north_star asks about ("the slower of bytes at the HBM peak and DFT + mel FLOPs at the FMA
    peak"), plus the SM resource ncu shows closest to saturation (L1 / shared-memory data pipe, one wavefront per cycle and
    SM).  Numerators are COUNTED per B=64 launch by the committed ncu capture (profiles/ncu_traffic.json), the time is this
    run's kernel_ms; None if there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        hz = sm_mhz * 1e6
        fma_peak = sm_count * 128 * 2 * hz                      # fp32: 128 FFMA lanes per SM and cycle
        flops = float(d["fp32_flops_per_launch"])
        wf = float(d["lsu_wavefronts_per_sm"])
        t = kernel_ms * 1e-3
        return {"fma_fp32": {"flops_per_launch": flops, "achieved_tflops": flops / t / 1e12, "peak_tflops": fma_peak / 1e12,
                             "frac": flops / t / fma_peak, "ms_at_peak": flops / fma_peak * 1e3,
                             "what": "fp32 operations the kernel executes (ncu thread-instruction counts: FFMA = 2, FFMA2 = 4), "
                                     "peak = SMs x 128 FFMA x 2 x SM clock; smaller time than the HBM bound, so HBM stays the roofline"},
                "l1_data_pipe": {"wavefronts_per_sm_per_launch": wf, "ms_at_peak": wf / hz * 1e3, "frac": wf / hz / t,
                                 "what": "L1 / shared-memory data-pipe wavefronts per SM (ncu l1tex__data_pipe_lsu_wavefronts), one "
                                         "per cycle at best: the SM resource that binds this kernel (FFT exchange, power tile, "
                                         "mel taps through shared memory)"},
                "source": "profiles/ncu_traffic.json (counted by ncu, not live)"}
    except Exception:  # noqa: BLE001
        return None


def workload_config(n_gpus):
    return {
        "workload": "large-v3 front end: n_mels=128, 64 x 30 s float32 PCM per GPU per step, log-mel + SpecAugment "
                    "time/freq masks (T=100, F=43, p=1.0), device-resident input",
        "n_mels": N_MELS, "clips_per_gpu_per_step": BATCH, "global_batch": BATCH * n_gpus, "pcm_dtype": "float32",
        "spec_augment": True, "sharding": "DistributedSampler-style batch shards, no data-path collective",
        "cache": "4 rotating PCM sets (492 MB) and 16 rotating output sets (1.57 GB) per GPU > 126 MB L2; 16 output sets because an "
                 "independent launch must stay clear of every batch that may still be in flight (the workspace ring is 16 deep)",
        "streams": N_STREAMS,
        "in_flight": "one stream, consecutive batches declared independent (wft.set_overlap(True) -> WFT_LAUNCH_OVERLAP: programmatic "
                     "dependent launches that do not wait for the previous batch's grids, workspace ring of 16 counter sets); "
                     "roofline.serialized is the same loop with every batch waiting for the one before it",
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


# ---- the reference's CPU path (oracle port) in the shapes the reference can run it in -------------------------------------
# The reference computes features inside DataLoader workers: min(cpu_count, 8) worker PROCESSES, each single-clip,
# (src/whisper_finetune/scripts/finetune.py:631, 641-664), so process-level parallelism is its real deployment shape;
# torch intra-op threading of one process is the other way a user could run it.  Every shape runs the SAME work: the
# 64-clip batch of the CUDA arm, log-mel + masks clip by clip like data_loader.py:273-292 (oracle/pipeline.py).

def _ref_worker(w, n_workers, steps, warmup, barrier, queue):
    """One DataLoader-style worker process: 1 torch thread, clips w::n_workers of every 64-clip batch."""
    import torch

    from oracle import pipeline as OP
    from oracle import specaug as OS

    torch.set_num_threads(1)
    pcm = synth_pcm(BATCH, SEED)[w::n_workers].contiguous()
    masks = OS.draw_mask_params(SEED, 0, BATCH, N_MELS, N_FRAMES, TIME_MASK, FREQ_MASK, 1.0)[w::n_workers]
    if pcm.shape[0]:
        OP.front_end_batch(pcm[:1], N_MELS, masks=masks[:1])
        for _ in range(warmup):
            OP.front_end_batch(pcm, N_MELS, masks=masks)
    barrier.wait()
    t0 = time.perf_counter()      # CLOCK_MONOTONIC: comparable across processes of one host
    for _ in range(steps):
        if pcm.shape[0]:
            OP.front_end_batch(pcm, N_MELS, masks=masks)
    queue.put((w, t0, time.perf_counter()))


def _time_worker_processes(n_workers, steps, warmup):
    """-> seconds for `steps` 64-clip batches split over `n_workers` single-threaded processes."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    barrier, queue = ctx.Barrier(n_workers), ctx.Queue()
    procs = [ctx.Process(target=_ref_worker, args=(w, n_workers, steps, warmup, barrier, queue), daemon=True)
             for w in range(n_workers)]
    for p in procs:
        p.start()
    spans = [queue.get(timeout=1800) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    return max(s[2] for s in spans) - min(s[1] for s in spans)


def _time_one_process(threads, steps, warmup):
    import torch

    from oracle import pipeline as OP
    from oracle import specaug as OS

    torch.set_num_threads(threads)
    pcm = synth_pcm(BATCH, SEED)
    masks = OS.draw_mask_params(SEED, 0, BATCH, N_MELS, N_FRAMES, TIME_MASK, FREQ_MASK, 1.0)
    OP.front_end_batch(pcm[:2], N_MELS, masks=masks[:2])  # FFT plans, filter cache
    for _ in range(warmup):
        OP.front_end_batch(pcm, N_MELS, masks=masks)
    t0 = time.perf_counter()
    for _ in range(steps):
        OP.front_end_batch(pcm, N_MELS, masks=masks)
    return time.perf_counter() - t0


def cpu_reference_shapes(steps, warmup, budget_s=None):
    """Time the CPU path in every shape -> (best shape name, {shape: {...}}).  `budget_s` bounds each shape's sample."""
    cores = os.cpu_count() or 1
    shapes = {}

    def record(name, threads, procs, k, el):
        shapes[name] = {"value": BATCH * k / el, "unit": UNIT, "processes": procs, "threads_per_process": threads,
                        "cores": procs * threads, "clips": BATCH * k, "seconds": el}

    def bounded(k, per_step_guess):
        if budget_s is None:
            return k
        return max(1, min(k, int(budget_s / max(per_step_guess, 1e-3))))

    k1 = bounded(steps, 1.0)                       # ~110 clips/s on one thread
    record("1_process_x_1_thread", 1, 1, k1, _time_one_process(1, k1, min(warmup, 1)))
    per_step_1t = shapes["1_process_x_1_thread"]["seconds"] / k1
    k = bounded(steps, per_step_1t / min(cores, 4))
    record(f"1_process_x_{cores}_threads_intra_op", cores, 1, k, _time_one_process(cores, k, warmup))
    w8 = min(cores, 8)                             # finetune.py:631
    k = bounded(steps, per_step_1t / w8)
    record(f"{w8}_worker_processes_x_1_thread", 1, w8, k, _time_worker_processes(w8, k, warmup))
    if cores > w8:
        k = bounded(steps, per_step_1t / cores)
        record(f"{cores}_worker_processes_x_1_thread", 1, cores, k, _time_worker_processes(cores, k, warmup))
    best = max(shapes, key=lambda n: shapes[n]["value"])
    return best, shapes


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
    whisper.audio is a third-party dependency that is not installable offline), timed in every shape the reference can
    run it in; `value` is the BEST of them (the process-parallel DataLoader shape on any multi-core host).  A step is a
    bounded sample of the workload: one 64-clip batch (the same batch the CUDA arm runs per GPU per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    best, shapes = cpu_reference_shapes(args.steps, args.warmup)
    b = shapes[best]
    val = b["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * BATCH / val, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": b["cores"], "kind": "port", "shape": best,
                         "sample": f"{b['clips']} clips in {b['seconds']:.1f} s ({BATCH} clips per step), {b['processes']} "
                                   f"process(es) x {b['threads_per_process']} torch thread(s) on {cores} host cores, "
                                   "oracle/pipeline.py (torch.stft recipe + masks, clip by clip); best of `shapes`",
                         "shapes": shapes},
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
    n_sets = 4     # PCM (read only)
    n_out = 16     # outputs: an independent launch (set_overlap) stays clear of every call since the last one that waited
    # every rank owns its shard of the synthetic "dataset": clip ids rank*B + step*world*B ... (weak scaling)
    pcm_sets = [synth_pcm(BATCH, SEED + 1000 * rank + s).to(dev) for s in range(n_sets)]
    out_sets = [torch.empty(BATCH, N_MELS, N_FRAMES, device=dev) for _ in range(n_out)]
    # two batches in flight (a loader with one batch of prefetch): step i runs on stream i % 2, so the ramp-down of one
    # launch (CTAs running out of tiles) is filled by the next batch's CTAs.  Buffer set i % 4 is only ever used on
    # stream i % 2, so launches that share buffers stay ordered.
    streams = [torch.cuda.Stream(device=dev) for _ in range(N_STREAMS)]

    def step(i):
        with torch.cuda.stream(streams[i % N_STREAMS]):
            return fe(pcm_sets[i % n_sets], clip_offset=(i * world + rank) * BATCH, out=out_sets[i % n_out])

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

    def timed_block(first_step, body):
        """EXACTLY args.steps steps between two events on the current stream, barrier + synchronize on both sides;
        -> milliseconds, max over ranks."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        body(first_step)
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def repeat_blocks(body, min_gpu_seconds=0.5, max_blocks=400):
        """One K-step block lasts only milliseconds: repeat it until >= 0.5 s of GPU time has been timed and report the
        MEDIAN block (every rank runs the same number of blocks: the count comes from the max-over-ranks first block)."""
        first = timed_block(0, body)
        n_blocks = int(min(max_blocks, max(3, min_gpu_seconds * 1e3 / max(first, 1e-3))))
        times = [first] + [timed_block((k + 1) * args.steps, body) for k in range(n_blocks)]
        return statistics.median(times), times

    def value_body(first_step):
        fork()
        for i in range(first_step, first_step + args.steps):
            step(i)
        join()

    # consecutive steps use different buffer sets (4 rotating sets), so they are independent batches
    wft.set_overlap(True)
    with ClockSampler(local_rank) as clocks:
        block_ms, block_times = repeat_blocks(value_body)
        launches_total = int(lib.wft_launch_count(0))
    launches = launches_total // len(block_times)      # launches of ONE timed K-step block
    ms_max = block_ms
    value = world * BATCH * args.steps / (ms_max * 1e-3)

    # fused kernel alone (no mask draw), for the roofline: K back-to-back launches on the current stream per block
    masks = wft.draw_mask_params(SEED, 0, BATCH, N_MELS, N_FRAMES, TIME_MASK, FREQ_MASK, 1.0, dev)
    for i in range(3):
        wft.frontend_forward(pcm_sets[i % n_sets], N_MELS, mask_params=masks, out=out_sets[i % n_out])
    torch.cuda.synchronize()

    def kernel_body(first_step):
        for i in range(first_step, first_step + args.steps):
            wft.frontend_forward(pcm_sets[i % n_sets], N_MELS, mask_params=masks, out=out_sets[i % n_out])

    kernel_block_ms, kernel_times = repeat_blocks(kernel_body, min_gpu_seconds=0.3)
    kernel_ms = kernel_block_ms / args.steps
    wft.set_overlap(False)       # the same loop, every batch waiting for the one in front of it (plain programmatic launches)
    serial_block_ms, _ = repeat_blocks(kernel_body, min_gpu_seconds=0.2)
    serial_ms = serial_block_ms / args.steps
    wft.set_overlap(True)

    # the production configs also time-warp (configs/config_turbo_best.yaml:97, config_large_v3_best_muon_ddp4.yaml:120:
    # time_warp_w = 80; applied at data_loader.py:285): front-end kernel -> ONE fused epilogue pass (warp + masks)
    fe_warp = wft.FrontEnd(n_mels=N_MELS, device=dev, spec_augment=True, seed=SEED,
                           spec_augment_params={"time_mask_param": TIME_MASK, "freq_mask_param": FREQ_MASK,
                                                "time_warp_w": TIME_WARP_W, "p": 1.0})
    warp_out = [torch.empty(BATCH, N_MELS, N_FRAMES, device=dev) for _ in range(n_out)]

    def warp_body(first_step):
        fork()
        for i in range(first_step, first_step + args.steps):
            with torch.cuda.stream(streams[i % N_STREAMS]):
                fe_warp(pcm_sets[i % n_sets], clip_offset=(i * world + rank) * BATCH, out=warp_out[i % n_out])
        join()

    warp_body(0)
    torch.cuda.synchronize()
    lib.wft_launch_count(1)
    warp_block_ms, warp_times = repeat_blocks(warp_body, min_gpu_seconds=0.3)
    warp_launches = int(lib.wft_launch_count(0)) // len(warp_times)
    wft.set_overlap(False)
    value_warp = world * BATCH * args.steps / (warp_block_ms * 1e-3)
    # the epilogue kernel alone, for its own roofline: one read + one write of the features
    warps = wft.draw_warp_params(SEED, 0, BATCH, N_FRAMES, TIME_WARP_W, 1.0, dev)

    def epilogue_body(first_step):
        for i in range(first_step, first_step + args.steps):
            wft.augment_epilogue(out_sets[i % n_out], warps, masks, None, 0.0, out=warp_out[i % n_out])

    epilogue_body(0)
    torch.cuda.synchronize()
    epi_block_ms, _ = repeat_blocks(epilogue_body, min_gpu_seconds=0.2)
    epi_ms = epi_block_ms / args.steps

    # end to end through the public API with host buffers (pinned): H2D + kernels + D2H every step
    def run_e2e(pcm_dtype, readback, front_end=None):
        host_pcm = [synth_pcm(BATCH, SEED + 1000 * rank + s) for s in range(2)]
        if pcm_dtype == torch.int16:
            host_pcm = [(x * 32767).round().to(torch.int16) for x in host_pcm]
        host_pcm = [x.pin_memory() for x in host_pcm]
        shape = (BATCH, N_MELS, N_FRAMES) if readback == "features" else (BATCH,)
        host_out = [torch.empty(shape).pin_memory() for _ in range(2)]
        pipe = wft.HostPipeline(front_end or fe, BATCH, pcm_dtype=pcm_dtype, n_chunks=4, n_streams=2, readback=readback)
        steps = max(3, min(args.steps, 20))
        for i in range(2):
            pipe(host_pcm[i % 2], host_out[i % 2], clip_offset=i * BATCH)
        pipe.synchronize()

        def body(first_step):
            for i in range(first_step, first_step + steps):
                pipe(host_pcm[i % 2], host_out[i % 2], clip_offset=(i * world + rank) * BATCH)
            pipe.join()   # the current stream (and so the closing event) waits for the last D2H copy

        # a block is `steps` batches; repeated until >= 0.3 s are timed, median block (same helper as `value`)
        saved, args.steps = args.steps, steps
        try:
            ms, times = repeat_blocks(body, min_gpu_seconds=0.3, max_blocks=20)
        finally:
            args.steps = saved
        pipe.synchronize()
        probe = float(host_out[(steps - 1) % 2].reshape(BATCH, -1)[0, :8].sum())  # the result really is on the host
        return world * BATCH * steps / (ms * 1e-3), pipe.h2d_bytes, pipe.d2h_bytes, steps, probe

    # Declared end-to-end shape (the only one a trainer uses): DataLoader workers hand over int16 PCM in
    # pinned host memory, the features stay in HBM for the model (train_step moves x to the device anyway,
    # model/model_utils.py:59-62) and one float32 per clip comes back as the step's read-back.  The other three
    # combinations (float32 PCM, full features back to the host = what the reference's CPU path yields) are variants.
    e2e_value, e2e_h2d, e2e_d2h, e2e_steps, checksum = run_e2e(torch.int16, "probe")
    e2e_variants = {}
    for name, dt, rb in (("f32_pcm_features_to_host", torch.float32, "features"),
                         ("int16_pcm_features_to_host", torch.int16, "features"),
                         ("f32_pcm_features_stay_on_device", torch.float32, "probe")):
        v, h2d, d2h, _, _ = run_e2e(dt, rb)
        e2e_variants[name] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    # the declared shape with the production configs' time-warp (front-end grid -> epilogue grid per chunk): the bus is still
    # what bounds it
    v, h2d, d2h, _, _ = run_e2e(torch.int16, "probe", front_end=fe_warp)
    e2e_variants["int16_pcm_time_warp_features_stay_on_device"] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    # multi-GPU correctness on the hardware (outside every timed region): each rank computes ITS DistributedSampler shard of
    # a fixed set of 8 * world clips (masks keyed by the global clip index), the shards are all-gathered over NCCL, and
    # rank 0 recomputes the whole set alone: the union of the shards must equal the single-GPU result bit for bit.
    multi = None
    if world > 1:
        per = 8
        total = per * world
        idx = wft.shard_indices(total, world, rank, epoch=0, seed=SEED, shuffle=True)
        all_pcm = synth_pcm(total, SEED + 77)
        gmasks = wft.draw_mask_params(SEED, 0, total, N_MELS, N_FRAMES, TIME_MASK, FREQ_MASK, 1.0, dev)
        mine = wft.frontend_forward(all_pcm[idx].to(dev), N_MELS, mask_params=gmasks[torch.as_tensor(idx, device=dev)].contiguous())
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gathered = wft.all_gather_features(mine)          # warm-up (NCCL channel setup)
        barrier()
        g0.record()
        gathered = wft.all_gather_features(mine)
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        order = torch.tensor([i for r in range(world) for i in wft.shard_indices(total, world, r, epoch=0, seed=SEED, shuffle=True)])
        equal = torch.tensor([1], device=dev)
        if rank == 0:
            alone = wft.frontend_forward(all_pcm.to(dev), N_MELS, mask_params=gmasks)
            equal = torch.tensor([int(torch.equal(gathered, alone[order.to(dev)]))], device=dev)
        dist.broadcast(equal, 0)
        nbytes = gathered.numel() * 4
        multi = {"shard_union_equal": bool(equal.item()), "clips": total,
                 "all_gather": {"backend": dist.get_backend(), "bytes_out_per_rank": nbytes, "ms": float(tg.item()),
                                "algbw_gbs": nbytes / (float(tg.item()) * 1e-3) / 1e9,
                                "busbw_gbs": nbytes * (world - 1) / world / (float(tg.item()) * 1e-3) / 1e9}}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        achieved = BATCH * BYTES_PER_CLIP / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "audio_hours_per_s": value * 30.0 / 3600.0,
            "timing": {"what": "median of repeated K-step blocks, each bracketed by barrier + synchronize, CUDA events, max over ranks",
                       "blocks": len(block_times), "block_ms_median": block_ms, "block_ms_min": min(block_times),
                       "block_ms_max": max(block_times), "gpu_seconds_timed": sum(block_times) * 1e-3,
                       "kernel_blocks": len(kernel_times), "kernel_gpu_seconds_timed": sum(kernel_times) * 1e-3},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                    "steps": e2e_steps, "pipeline": "4 chunks on 2 streams, consecutive batches overlap, pinned host buffers, int16 PCM in "
                    "(what a DataLoader worker hands over), features stay in HBM for the model, one float32 per clip read back",
                    "checksum": checksum},
            "e2e_variants": e2e_variants,
            "gpu_launches": launches,
            "gpu_launches_per_step": launches / max(args.steps, 1),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "wft::frontend_kernel<128,float> (+ its fix-up grid)", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": BATCH * BYTES_PER_CLIP,
                         "what": "K back-to-back batches on one stream, pre-drawn masks, independent batches overlapped; "
                                 "kernel_ms = timed region / K",
                         "serialized": {"kernel_ms": serial_ms, "achieved": BATCH * BYTES_PER_CLIP / (serial_ms * 1e-3) / 1e9,
                                        "frac": BATCH * BYTES_PER_CLIP / (serial_ms * 1e-3) / 1e9 / peak,
                                        "what": "same loop, every batch waits for the previous one (no overlap)"}},
        }
        epi_bytes = 2 * 4 * N_MELS * N_FRAMES * BATCH
        line["value_with_time_warp"] = {
            "value": value_warp, "unit": UNIT, "ms_per_step": warp_block_ms / args.steps, "gpu_launches_per_step": warp_launches / args.steps,
            "what": "ONE call, two grids: front-end grid -> fused epilogue grid (finishes the cells on load: max-8 floor / pad; time-warp "
                    "W=80 -> time mask -> frequency mask) that draws the clip's warp point and mask intervals itself; device-resident "
                    "PCM, same blocks / stream as `value`",
            "epilogue_roofline": {"bound": "hbm", "kernel": "augment_staged_kernel<false, 0>", "kernel_ms": epi_ms,
                                  "algorithmic_bytes_per_launch": epi_bytes, "achieved": epi_bytes / (epi_ms * 1e-3) / 1e9,
                                  "peak": peak, "unit": "GB/s", "frac": epi_bytes / (epi_ms * 1e-3) / 1e9 / peak}}
        props = torch.cuda.get_device_properties(local_rank)
        sm_mhz = (line["clocks"] or {}).get("sm_mhz") or (line["clocks"] or {}).get("sm_max_mhz") or 1965.0
        line["roofline"]["other_bounds"] = other_bounds(kernel_ms, props.multi_processor_count, float(sm_mhz))
        if multi is not None:
            line["multi_gpu"] = multi
            line["shard_union_equal"] = multi["shard_union_equal"]
        if world == 1 and not args.no_cpu_baseline:
            best, shapes = cpu_reference_shapes(20, 1, budget_s=6.0)   # bounded: ~6 s per shape
            b = shapes[best]
            line["cpu_baseline"] = {"value": b["value"], "unit": UNIT, "cores": b["cores"], "kind": "port", "shape": best,
                                    "sample": f"{b['clips']} clips in {b['seconds']:.1f} s, oracle/pipeline.py clip by clip, "
                                              f"{b['processes']} process(es) x {b['threads_per_process']} thread(s) of "
                                              f"{os.cpu_count()} host cores; best of `shapes`",
                                    "shapes": shapes}
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
