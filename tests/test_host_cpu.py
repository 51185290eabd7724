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
    assert wft._lib.load().wft_abi_version() == wft._lib.ABI_VERSION == 2


def test_host_validation_without_gpu(wft):
    """Argument errors are reported through the C ABI's error channel (no compute call is made)."""
    lib = wft._lib.load()
    need = ctypes.c_size_t(0)
    assert lib.wft_frontend_workspace_bytes(64, 480000, 3000, ctypes.byref(need)) == 0
    assert need.value >= 16 + 16 * 64 + 4 * 64 * 94 and need.value % 256 == 0
    assert lib.wft_frontend_workspace_bytes(0, 480000, 3000, ctypes.byref(need)) == wft._lib.WFT_ERR_INVALID
    assert b"batch" in lib.wft_last_error()
    assert lib.wft_frontend_workspace_bytes(1, 200, 0, ctypes.byref(need)) == wft._lib.WFT_ERR_INVALID
    args = wft._lib.FrontendArgs(n_mels=64, batch=1, n_samples=16000, clip_stride=16000)
    args.pcm = args.out = args.workspace = 1  # never dereferenced: validation fails first
    assert lib.wft_frontend_forward(ctypes.byref(args), None) == wft._lib.WFT_ERR_INVALID
    assert b"n_mels" in lib.wft_last_error()
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
        assert len(plan["thread"]) == 160 and sorted({t[0] for t in plan["thread"]}) == list(range(n_mels))
        assert all(1 <= t[1] and t[1] + plan["classes"][plan["warp_class"][i // 32]][0] - 1 <= 199
                   for i, t in enumerate(plan["thread"]))


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
