"""CPU tests of the host side: C-ABI surface, package helpers, error behaviour without a GPU, sharding (gloo, 2 ranks)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "wft.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wft_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol(wft):
    lib_path = wft._lib.LIB_PATH
    if not os.path.exists(lib_path):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(lib_path)
    declared = _header_functions()
    assert declared, "no functions parsed from include/wft.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wft.h but not exported"
    assert sorted(wft._lib.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert wft._lib.load().wft_abi_version() == wft._lib.ABI_VERSION == 9


def test_host_validation_without_gpu(wft):
    """Argument errors are reported through the C ABI's error channel (no compute call is made)."""
    lib = wft._lib.load()
    need = ctypes.c_size_t(0)
    assert lib.wft_frontend_workspace_bytes(64, 480000, 3000, ctypes.byref(need)) == 0
    assert need.value >= 16 + 8 * 64 + 4 * 64 * 188 and need.value % 256 == 0
    assert lib.wft_frontend_workspace_bytes(0, 480000, 3000, ctypes.byref(need)) == wft._lib.WFT_ERR_INVALID
    assert b"batch" in lib.wft_last_error()
    assert lib.wft_frontend_workspace_bytes(1, 200, 0, ctypes.byref(need)) == wft._lib.WFT_ERR_INVALID
    args = wft._lib.FrontendArgs(n_mels=64, batch=1, n_samples=16000, clip_stride=16000)
    args.pcm = args.out = args.workspace = 1  # never dereferenced: validation fails first
    assert lib.wft_frontend_forward(ctypes.byref(args), None) == wft._lib.WFT_ERR_INVALID
    assert b"n_mels" in lib.wft_last_error()
    # workspace modes / launch flags (include/wft.h): an overlapping launch needs the workspace ring, unknown modes are rejected
    need16 = ctypes.c_size_t(0)
    assert lib.wft_frontend_workspace_bytes(1, 16000, 0, ctypes.byref(need16)) == 0
    ok = dict(n_mels=80, batch=1, n_samples=16000, clip_stride=16000, workspace_bytes=need16.value)
    for mode, flags, word in ((wft._lib.WFT_WS_MEMSET, wft._lib.WFT_LAUNCH_OVERLAP, b"WFT_WS_RING"),
                              (wft._lib.WFT_WS_PHASE_A, wft._lib.WFT_LAUNCH_OVERLAP, b"WFT_WS_RING"),
                              (3, 0, b"workspace_mode"), (wft._lib.WFT_WS_RING + wft._lib.WFT_WS_PHASES, 0, b"workspace_mode")):
        args = wft._lib.FrontendArgs(workspace_mode=mode, launch_flags=flags, **ok)
        args.pcm = args.out = args.workspace = 16
        assert lib.wft_frontend_forward(ctypes.byref(args), None) == wft._lib.WFT_ERR_INVALID
        assert word in lib.wft_last_error(), lib.wft_last_error()
    # the ring needs WFT_WS_PHASES copies of the counters and of the per-tile scratch
    assert need16.value >= wft._lib.WFT_WS_PHASES * (16 + 8 * 1 + 4 * 7)
    with pytest.raises(ValueError):
        wft._lib.check(wft._lib.WFT_ERR_INVALID)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(wft):
    with pytest.raises(RuntimeError, match="CUDA"):
        wft.log_mel_spectrogram(np.zeros(16000, dtype=np.float32))
    with pytest.raises(RuntimeError, match="CUDA"):
        wft.pad_or_trim(torch.zeros(80, 10), 3000)
    with pytest.raises(RuntimeError, match="CUDA"):
        wft.FrontEnd(n_mels=80)
    with pytest.raises(NotImplementedError):
        wft.log_mel_spectrogram("clip.wav")
    with pytest.raises(ValueError):
        wft.log_mel_spectrogram(np.zeros(16000, dtype=np.float32), n_mels=64)
    x = torch.zeros(4, 3000)
    assert wft.pad_or_trim(x, 3000) is x  # the no-op needs no device


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "whisper-finetune_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f"{f} touches oracle/"


@pytest.mark.parametrize("n_mels", [80, 128])
def test_package_mel_bank_equals_oracle_bank(wft, n_mels):
    from oracle.mel_filters import mel_filters

    assert np.array_equal(wft.slaney_mel_bank(n_mels), mel_filters(n_mels))
    with pytest.raises(ValueError):
        wft.slaney_mel_bank(64)


def test_generated_tables_are_current(wft):
    """csrc/wft_tables.inc on disk is what gen_tables.py generates (window = torch.hann_window, twiddles, mel plan)."""
    import importlib.util

    path = os.path.join(ROOT, "whisper-finetune_b200", "csrc", "gen_tables.py")
    spec = importlib.util.spec_from_file_location("_gen", path)
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert gen.generate() == open(os.path.join(os.path.dirname(path), "wft_tables.inc")).read()
    w = gen.window_table()
    assert np.array_equal(w.T.reshape(-1), torch.hann_window(400).numpy())  # [n2][n1] -> w[20 n1 + n2]
    # thread-level mel plan == dense bank product with every (row, frame) cell computed exactly once, for both banks
    for n_mels in (80, 128):
        bank = wft.slaney_mel_bank(n_mels)
        plan = gen.mel_plan(bank)
        gen.simulate(bank, plan)
        rows = sorted({row for passes in plan["thread"] for row, _ in passes})
        assert len(plan["thread"]) == 160 and rows == list(range(n_mels))
        for i, passes in enumerate(plan["thread"]):
            assert len(passes) == len(plan["warp_t"][i // 32]) <= 2
            assert all(1 <= start and start + taps - 1 <= 199 for (_, start), taps in zip(passes, plan["warp_t"][i // 32]))


@pytest.mark.parametrize("n,world,drop_last,shuffle", [(100, 4, False, True), (101, 4, True, True), (7, 8, False, True),
                                                         (64, 2, True, False), (1000, 8, False, True)])
def test_shard_indices_match_distributed_sampler(wft, n, world, drop_last, shuffle):
    from torch.utils.data import DistributedSampler

    for epoch in (0, 5):
        seen = []
        for rank in range(world):
            ds = DistributedSampler(range(n), num_replicas=world, rank=rank, shuffle=shuffle, seed=7, drop_last=drop_last)
            ds.set_epoch(epoch)
            mine = wft.shard_indices(n, world, rank, epoch=epoch, seed=7, shuffle=shuffle, drop_last=drop_last)
            assert mine == list(iter(ds))
            seen += mine
        if not drop_last:
            assert set(seen) == set(range(n))
    with pytest.raises(ValueError):
        wft.shard_indices(10, 2, 2)


def test_reference_data_loader_boundary_names(wft):
    """install() rebinds exactly the module attributes the reference resolves at call time
    (data_loader.py:13,16-20; tests/test_data_loader.py:203-207 patch the same names)."""
    import types

    dl = types.ModuleType("whisper_finetune.data.data_loader")
    dl.log_mel_spectrogram = dl.pad_or_trim = None
    wa = types.ModuleType("whisper.audio")
    wa.log_mel_spectrogram = None
    saved = {k: sys.modules.get(k) for k in ("whisper.audio", "whisper_finetune.data.data_loader")}
    sys.modules["whisper.audio"], sys.modules["whisper_finetune.data.data_loader"] = wa, dl
    try:
        wft.install()
        assert dl.log_mel_spectrogram is wft.log_mel_spectrogram and dl.pad_or_trim is wft.pad_or_trim
        assert wa.log_mel_spectrogram is wft.log_mel_spectrogram
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import whisper_finetune_b200 as wft
from oracle import specaug as OS
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n, per = 16, 8
mine = wft.shard_indices(n, world, rank, epoch=1, seed=42, shuffle=True, drop_last=True)
assert len(mine) == per
# every rank "computes" a stand-in feature block for its shard whose content depends only on the GLOBAL clip index
# (mask parameters from the counter-based draw), then the optional all-gather reassembles DistributedSampler order
local = torch.stack([torch.from_numpy(OS.draw_mask_params(42, i, 1, 128, 3000, 100, 27, 1.0)[0]).float() for i in mine])
full = wft.all_gather_features(local.view(per, 1, 4))
assert full.shape == (world * per, 1, 4)
order = full.view(world, per, 4).transpose(0, 1).reshape(-1, 4)
g = torch.Generator(); g.manual_seed(42 + 1)
perm = torch.randperm(n, generator=g).tolist()
want = torch.stack([torch.from_numpy(OS.draw_mask_params(42, i, 1, 128, 3000, 100, 27, 1.0)[0]).float() for i in perm])
assert torch.equal(order, want), "gathered shards must equal the single-process result"
dist.barrier(); dist.destroy_process_group()
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"rank{rank}.ok"), "w").write("ok")
"""


def test_two_rank_gloo_sharding_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()  # (stdout of the two ranks interleaves)


def test_bench_reference_arm_prints_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
                "cpu_baseline", "e2e"):
        assert key in line
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_pcm_record_round_trip(wft):
    """encode_pcm_record / decode_pcm_records: samples, length, cut, gate and extremes survive, for both PCM types."""
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(12345) * 0.1).astype(np.float32)
    a[-20:] = 0                                           # trailing zeros are padding
    for dt in (torch.float32, torch.int16):
        recs = [wft.encode_pcm_record(a, 1234, True, dt, (3, 4)), wft.encode_pcm_record(np.zeros(7, np.float32), None, False, dt)]
        assert all(r.dtype == dt and r.shape == (480008,) for r in recs)
        b = wft.decode_pcm_records(recs)
        assert b.pcm.dtype == dt and tuple(b.pcm.shape) == (2, 480000)
        assert b.lengths.tolist() == [12325, 0] and b.n_valid_frames.tolist() == [1234, -1]
        assert b.augment.tolist() == [1, 0] and b.extremes.tolist() == [[3, 4], [0, 0]]
        if dt == torch.float32:
            assert np.array_equal(b.pcm[0, :12345].numpy(), a)
        else:
            assert np.array_equal(b.pcm[0, :12345].numpy(), np.rint(a.astype(np.float64) * 32768).astype(np.int16))
        assert not b.pcm[0, 12345:].any()
    assert wft.encode_pcm_record(a, 5000)[480001] == 3001    # a cut beyond the clip keeps every frame


_LOADER_DRIVER = r"""
import sys, types
root, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [root, ref + "/src"]
import numpy as np, torch
import whisper_finetune_b200 as wft

w = types.ModuleType("whisper"); wa = types.ModuleType("whisper.audio"); wt = types.ModuleType("whisper.tokenizer")
wa.CHUNK_LENGTH, wa.HOP_LENGTH, wa.N_FFT, wa.N_FRAMES, wa.N_SAMPLES = 30, 160, 400, 3000, 480000
wa.log_mel_spectrogram = lambda *a, **k: (_ for _ in ()).throw(AssertionError("features must not be computed on the host"))
wt.LANGUAGES, wt.TO_LANGUAGE_CODE, wt.Tokenizer = {"de": "german"}, {"german": "de"}, object
w.audio, w.tokenizer = wa, wt
sys.modules.update({"whisper": w, "whisper.audio": wa, "whisper.tokenizer": wt})
class _Any(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"): raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, x, **k: x})
sys.modules["audiomentations"] = _Any("audiomentations")
from whisper_finetune.data import data_loader as dl

ref_collate = dl.collate_fn
ref_gate = dl.AudioDataset._should_apply_spec_augment
wft.install_loader(dl, pcm_dtype=torch.int16)
assert dl.collate_fn is wft.pcm_collate_fn and dl.AudioDataset._calculate_mel is wft.deferred_calculate_mel

ds = dl.AudioDataset.__new__(dl.AudioDataset)          # the way the reference's own tests build one
ds.aud_augment, ds.n_mels, ds.device = None, 128, None
ds.num_frames_per_second = dl.N_FRAMES / dl.CHUNK_LENGTH
ds.spec_augment, ds.spec_augment_p, ds.extreme_freq_masking = True, 0.5, dl.ExtremesFrequencyMasking(10, 10)
audio = np.zeros(480000, np.float32); audio[:32000] = 0.25
items = []
for seed in range(8):
    torch.manual_seed(seed)
    want_gate = ref_gate(ds)                            # the reference's draw under this seed ...
    r = torch.rand(1).item()                            # ... followed by the extremes ratio (data/utils.py:173)
    torch.manual_seed(seed)
    rec = ds._calculate_mel(audio, 12.34 if seed % 2 else None, True)
    b = wft.decode_pcm_records([rec])
    assert b.pcm.dtype == torch.int16 and int(b.pcm[0, 0]) == 8192 and b.lengths.tolist() == [32000]
    assert b.augment.tolist() == [int(want_gate)], seed
    assert b.n_valid_frames.tolist() == [1234 if seed % 2 else -1]
    assert b.extremes.tolist() == [[int(round(r * 10)), int(round(r * 10))]]
    items.append((rec, torch.arange(4 + seed), torch.full((3 + seed,), 7)))
batch, y_in, y_out = dl.collate_fn(items)
_, ry_in, ry_out = ref_collate([(torch.zeros(1, 3000), i, o) for _, i, o in items])
assert torch.equal(y_in, ry_in) and torch.equal(y_out, ry_out) and len(batch) == 8
try:
    ds._calculate_mel(audio, 0.001, True)               # cut at frame 0: the reference dies in pad_or_trim (torch.min of nothing)
    raise SystemExit("expected RuntimeError")
except RuntimeError:
    pass
print("LOADER-OK")
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present (GPU box)")
def test_install_loader_rewires_reference_dataset(tmp_path):
    """install_loader on the REAL reference module: gate / cut / extremes decisions equal the reference's, no host features."""
    import subprocess

    script = tmp_path / "drive.py"
    script.write_text(_LOADER_DRIVER)
    res = subprocess.run([sys.executable, str(script), ROOT, "/root/reference"], capture_output=True, text=True, timeout=600,
                         cwd=str(tmp_path))
    assert res.returncode == 0 and "LOADER-OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_torch_custom_ops_are_registered_with_fake_kernels(wft):
    """north_star: "thin torch custom ops over a C-ABI shim" -- every op exists in torch.ops.wft, carries a schema, and its
    fake kernel yields the right shape / dtype / device without a GPU (what torch.compile and make_fx trace through)."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    names = {"frontend_forward", "frontend_forward_out", "pad_or_trim", "specaug_apply", "specaug_apply_", "augment",
             "augment_out", "augment_", "specaug_draw", "time_warp_draw", "mask_bsd"}
    assert names <= {n for n in dir(torch.ops.wft) if not n.startswith("_")}
    assert "!) out" in str(torch.ops.wft.frontend_forward_out.default._schema), "out is declared as mutated"
    with FakeTensorMode():
        pcm = torch.empty(4, 480000, device="cuda")
        y = torch.ops.wft.frontend_forward(pcm, 128, 0, None, 3000, None, None, 0.0)
        assert tuple(y.shape) == (4, 128, 3000) and y.dtype == torch.float32 and y.device.type == "cuda"
        y = torch.ops.wft.frontend_forward(torch.empty(2, 16000, device="cuda", dtype=torch.int16), 80, 160, None, 0, None, None, 0.0)
        assert tuple(y.shape) == (2, 80, 101)
        assert tuple(torch.ops.wft.pad_or_trim(torch.empty(3, 70, 5, device="cuda"), 3000).shape) == (3, 3000, 5)
        mel = torch.empty(4, 128, 3000, device="cuda")
        assert tuple(torch.ops.wft.augment(mel, None, None, None, 0.0, False).shape) == (4, 128, 3000)
        assert tuple(torch.ops.wft.specaug_draw(mel, 1, 2, 7, 128, 3000, 100, 27, 1.0).shape) == (7, 4)
        assert tuple(torch.ops.wft.time_warp_draw(mel, 1, 2, 7, 3000, 80, 1.0).shape) == (7, 2)
        act = torch.empty(2, 1500, 1280, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        z = torch.ops.wft.mask_bsd(act, 1, 2, 3, 4)
        assert z.dtype == torch.bfloat16 and z.requires_grad, "the autograd formula is registered"


def test_independent_launch_bookkeeping(wft):
    """ops._record_call: a front-end call is launched as an independent batch (WFT_LAUNCH_OVERLAP: it waits for nothing) only
    if its buffers stay clear of EVERY call enqueued since the last waiting launch -- with small batches whole calls run side
    by side, so the call directly in front is not the only one that may still be in flight -- and a call whose epilogue grid
    outgrows the device bounds that set to itself."""
    from whisper_finetune_b200.ops import _record_call

    def call(pcm, out, scratch=None):
        return [(pcm, pcm + 100)], [(out, out + 100)] + ([(scratch, scratch + 100)] if scratch is not None else [])

    # first call on a stream: nothing to be independent of
    ind, hist = _record_call(None, *call(0, 1000), may_overlap=True, bounds_in_flight=False)
    assert not ind and len(hist) == 1
    # distinct buffers: independent, the history grows
    ind, hist = _record_call(hist, *call(200, 1200), may_overlap=True, bounds_in_flight=False)
    assert ind and len(hist) == 2
    ind, hist = _record_call(hist, *call(400, 1400), may_overlap=True, bounds_in_flight=False)
    assert ind and len(hist) == 3
    # re-using the output of the call THREE back: must wait (it may still be running), the history restarts
    ind, hist = _record_call(hist, *call(600, 1000), may_overlap=True, bounds_in_flight=False)
    assert not ind and len(hist) == 1
    # reading what an earlier call writes, or writing what it reads: wait
    ind, hist2 = _record_call(hist, *call(1000, 1600), may_overlap=True, bounds_in_flight=False)
    assert not ind
    ind, hist2 = _record_call(hist, *call(700, 600), may_overlap=True, bounds_in_flight=False)
    assert not ind
    # overlap switched off / ring wrap: never independent
    ind, hist2 = _record_call(hist, *call(800, 1800), may_overlap=False, bounds_in_flight=False)
    assert not ind and len(hist2) == 1
    # production calls with two alternating scratch buffers: fine once the epilogue grid bounds what is in flight ...
    hist = None
    for k in range(6):
        ind, hist = _record_call(hist, *call(5000, 6000 + 200 * k, scratch=9000 + 200 * (k % 2)), may_overlap=True, bounds_in_flight=True)
        assert ind == (k > 0) and len(hist) == 1
    # ... and a conflict with the call two back when it does not (small batches)
    hist, seen = None, []
    for k in range(6):
        ind, hist = _record_call(hist, *call(5000, 6000 + 200 * k, scratch=9000 + 200 * (k % 2)), may_overlap=True, bounds_in_flight=False)
        seen.append(ind)
    assert seen == [False, True, False, True, False, True]


def test_bench_other_bounds_come_from_the_committed_ncu_counts():
    """roofline.other_bounds (fp32 FMA-peak bound, L1 data-pipe bound) is arithmetic on profiles/ncu_traffic.json: the HBM
    bound must stay the slower of HBM and FMA (north_star's roofline rule), and the data pipe the tighter SM resource."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("_bench_for_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    for k in ("dram_bytes_per_launch", "fp32_flops_per_launch", "lsu_wavefronts_per_sm"):
        assert rec[k] > 0, k
    kernel_ms = 0.083
    ob = bench.other_bounds(kernel_ms, 148, 1965.0)
    assert ob is not None
    fma, pipe = ob["fma_fp32"], ob["l1_data_pipe"]
    assert abs(fma["peak_tflops"] - 148 * 128 * 2 * 1.965e9 / 1e12) < 1e-6
    assert abs(fma["frac"] - rec["fp32_flops_per_launch"] / (kernel_ms * 1e-3) / (fma["peak_tflops"] * 1e12)) < 1e-9
    assert abs(pipe["ms_at_peak"] - rec["lsu_wavefronts_per_sm"] / 1.965e9 * 1e3) < 1e-9
    hbm_ms = bench.BATCH * bench.BYTES_PER_CLIP / 6547.2e9 * 1e3
    assert fma["ms_at_peak"] < hbm_ms < pipe["ms_at_peak"] < kernel_ms
    assert 0.0 < fma["frac"] < pipe["frac"] < 1.0


def test_run_eager_takes_the_dispatcher_whenever_a_mode_is_active(wft):
    """ops.run_eager: the package's eager hot paths call an op's Python function directly only when nothing needs the
    dispatcher; under a dispatch mode (fake tensors, tracing) or a torch-function mode the registered op is what runs."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from torch.overrides import TorchFunctionMode
    from whisper_finetune_b200 import ops

    assert not ops._dispatch_modes_active()
    with FakeTensorMode():
        assert ops._dispatch_modes_active()
    with TorchFunctionMode():
        assert ops._dispatch_modes_active()
    assert not ops._dispatch_modes_active()

    class _Op:   # stands in for a CustomOpDef: __call__ = dispatcher, _init_fn = the registered function
        def __call__(self, *a):
            return ("dispatcher", a)

        @staticmethod
        def _init_fn(*a):
            return ("direct", a)

    assert ops.run_eager(_Op(), 1, 2) == ("direct", (1, 2))
    with FakeTensorMode():
        assert ops.run_eager(_Op(), 1, 2) == ("dispatcher", (1, 2))
    for op in (ops.frontend_forward, ops.frontend_forward_out, ops.frontend_forward_drawn_out, ops.frontend_augment_drawn_out):
        assert callable(op._init_fn)
