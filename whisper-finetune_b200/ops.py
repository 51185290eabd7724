"""``torch.ops.wft.*``: the C ABI (include/wft.h) as torch custom ops.

``north_star`` asks for "thin torch custom ops over a C-ABI shim": every kernel entry point of ``libwft_b200.so`` is
registered here with ``torch.library`` -- schema, CUDA implementation (a ctypes call on the tensor's data pointers and the
current stream), a fake / meta implementation (so ``torch.compile``, ``make_fx`` and FakeTensor tracing see shapes without
running anything) and, for the one op that sits on an autograd graph (``mask_bsd``, the deep-SpecAugment mask of
``model/model_utils.py:382-437``), its backward formula.  The reference-facing Python functions in ``audio.py`` /
``augment.py`` / ``deep.py`` validate their arguments and then call these ops; nothing else touches the library.

    wft::frontend_forward      PCM [B, N] (+ lengths, n_valid_frames, mask_params)        -> features [B, n_mels, T]
    wft::frontend_forward_out  same, written into a caller-owned buffer                   (mutates ``out``)
    wft::frontend_forward_drawn_out  the same with the SpecAugment intervals drawn inside the call (one host call per batch)
    wft::pad_or_trim           [outer, len_in, inner] float32 -> [outer, length, inner], min-value pad (data/utils.py:380-404)
    wft::specaug_apply         features, mask_params [B, 4]                               -> masked copy
    wft::specaug_apply_        in place
    wft::augment               time-warp -> masks -> extremes mask in one pass            -> new tensor
    wft::augment_out           same into ``out`` (``out`` may alias the input when there is no warp)
    wft::specaug_draw          (seed, clip_offset, ...)                                   -> int32 [B, 4]
    wft::time_warp_draw        (seed, clip_offset, ...)                                   -> int32 [B, 2]
    wft::mask_bsd              activations [B, S, D] fp32 / fp16 / bf16, spans            -> masked copy (differentiable)

CUDA only: there is no CPU kernel behind any of them (the ops are registered for ``device_types="cuda"``).
"""
import contextlib
import ctypes
from typing import Optional

import torch
from torch import Tensor

from . import _lib

HOP_LENGTH = 160
# (device, stream) -> [(reads, writes), ...]: byte ranges of every front-end call enqueued there since the last launch that WAITED
# for everything in front of it (an ordinary launch, the memset of a wrapping workspace ring, any other op of this package).
# An independent launch (WFT_LAUNCH_OVERLAP) never waits, so grids of ALL of these calls may still be running next to it --
# with small batches several whole calls are resident at once -- and it has to stay clear of every one of them, not just of
# the call directly in front.  (Grids of a stream complete in order, so a launch that waits for its predecessor has waited for
# all of them.)
_LAST_CALL = {}
_ELEM_BYTES = {torch.float32: 4, torch.float16: 2, torch.bfloat16: 2}


def _stream(device: torch.device) -> int:
    """Stream handle for any op OTHER than the front end: also forgets the stream's last front-end call, so that the next one
    is launched the ordinary way (an overlapping launch may only follow another front-end launch directly)."""
    st = torch.cuda.current_stream(device).cuda_stream
    _LAST_CALL.pop((device.index, st), None)
    return st


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


# Workspaces (include/wft.h, enum wft_workspace_mode): one per (device, stream, batch, size), used in WFT_WS_RING mode -- the
# library keeps WFT_WS_PHASES copies of its counters in it, every launch takes the next copy and the launch that wraps the ring
# zeroes all of them with one memset, so consecutive launches on a stream chain kernel to kernel.  A workspace is only ever
# used on the stream it was created for.
_WORKSPACES = {}
_MAX_WORKSPACES = 64


# Programmatic dependent launch of the fused kernel (include/wft.h, WFT_LAUNCH_PDL): on by default -- a training loop feeds
# one stream.  Code that keeps several streams of this GPU busy at once turns it off.
_PDL = {"enabled": True}
# Independent batches (WFT_LAUNCH_OVERLAP): off by default.  A caller that enqueues NOTHING but front-end calls on the stream
# between two batches (a loader that keeps several PCM batches resident, bench.py) switches it on: a launch then does not
# wait for the previous batch's grids to complete and fills the SM slots its tail frees.  The buffers of consecutive calls
# are still checked here -- a call that touches anything the previous call wrote (or writes anything it read) is launched
# the ordinary way.
_OVERLAP = {"enabled": False}


def set_programmatic_launch(enabled: bool) -> bool:
    """-> the previous setting."""
    old = _PDL["enabled"]
    _PDL["enabled"] = bool(enabled)
    return old


def set_overlap(enabled: bool) -> bool:
    """Declare consecutive front-end calls of a stream independent batches (see ``_OVERLAP``) -> the previous setting."""
    old = _OVERLAP["enabled"]
    _OVERLAP["enabled"] = bool(enabled)
    return old


_WS_BYTES = {}                       # (batch, n_total, n_frames_out) -> c_size_t: a pure function of the shape
_NO_GUARD = contextlib.nullcontext()   # the tensor's device is already current: no device switch around the call


def run_eager(op, *args):
    """Call a ``wft::`` op from this package's own eager hot paths: straight into the Python function the op was registered
    from when nothing needs the dispatcher (no compile / export tracing, no dispatch or function mode active) -- the
    registered op, its fake kernel and its schema stay what ``torch.compile`` and ``torch.ops.wft.*`` callers see.  A trip
    through the custom-op dispatcher costs ~25 us of host time per call, a third of a 64-clip batch's kernel time."""
    if torch.compiler.is_compiling() or _dispatch_modes_active():
        return op(*args)
    return op._init_fn(*args)


def _dispatch_modes_active() -> bool:
    try:
        return torch._C._len_torch_dispatch_stack() > 0 or torch._C._len_torch_function_stack() > 0
    except AttributeError:   # private counters moved: take the dispatcher
        return True


def _workspace(dev: torch.device, stream: int, batch: int, nbytes: int):
    # the layout inside a workspace depends on the batch size, so a workspace is only ever re-used for the same (batch, size)
    key = (dev.index, stream, batch, nbytes)
    ent = _WORKSPACES.get(key)
    if ent is None:
        if len(_WORKSPACES) >= _MAX_WORKSPACES:
            _WORKSPACES.pop(next(iter(_WORKSPACES)))   # freed stream-ordered by the caching allocator
        ent = _WORKSPACES[key] = [torch.empty(nbytes, dtype=torch.uint8, device=dev), -1]
    ent[1] += 1
    return key, ent, _lib.WFT_WS_RING + ent[1] % _lib.WFT_WS_PHASES


def _span(t: Optional[Tensor]):
    """Byte range [lo, hi) a (non-overlapping, positively strided) tensor touches."""
    if t is None or t.numel() == 0:
        return None
    lo = t.data_ptr()
    return (lo, lo + (1 + sum((n - 1) * st for n, st in zip(t.shape, t.stride()))) * t.element_size())


def _disjoint(a, b) -> bool:
    return a is None or b is None or a[1] <= b[0] or b[1] <= a[0]


def _launch_frontend(pcm: Tensor, n_mels: int, padding: int, lengths: Optional[Tensor], n_frames_out: int,
                     n_valid_frames: Optional[Tensor], mask_params: Optional[Tensor], mask_value: float, out: Tensor,
                     draw=None, aug=None) -> None:
    """``draw`` = (seed, clip_offset, time_mask_param, freq_mask_param, p): the intervals are drawn inside the call.
    ``aug`` = ``_lib.AugmentArgs`` + the tensors it points to: the call is ``wft_frontend_augment_forward`` (front-end grid ->
    augmentation epilogue that finishes the cells on load; ``out`` is then the scratch of the un-augmented features)."""
    lib = _lib.load()
    B, N = pcm.shape
    dev = pcm.device
    with (_NO_GUARD if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)):
        need = _WS_BYTES.get((B, N + padding, out.shape[2]))
        if need is None:
            need = ctypes.c_size_t(0)
            _lib.check(lib.wft_frontend_workspace_bytes(B, N + padding, out.shape[2], ctypes.byref(need)))
            if len(_WS_BYTES) < 1024:
                _WS_BYTES[(B, N + padding, out.shape[2])] = need
        stream = torch.cuda.current_stream(dev).cuda_stream
        flags = _lib.WFT_LAUNCH_PDL if _PDL["enabled"] else 0
        if torch.cuda.is_current_stream_capturing():
            # CUDA graph capture: what is recorded now is replayed verbatim, so neither the ring position (a host-side count)
            # nor the overlap decision may be baked in.  The call gets a workspace of its own from the graph's pool and the
            # mode whose zeroing is part of the call itself (a memset node in front of the kernels).
            key, mode = None, _lib.WFT_WS_MEMSET
            ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
            _LAST_CALL.pop((dev.index, stream), None)
        else:
            key, ent, mode = _workspace(dev, stream, B, need.value)
            ws = ent[0]
            reads = [_span(t) for t in (pcm, lengths, n_valid_frames, mask_params)]
            writes = [_span(out)]
            if aug is not None:
                reads += [_span(t) for t in aug[1]]
                writes += [_span(aug[2])]
            independent, hist = _record_call(
                _LAST_CALL.get((dev.index, stream)), reads, writes,
                may_overlap=_OVERLAP["enabled"] and mode != _lib.WFT_WS_RING,      # (ring position 0: the memset waits anyway)
                bounds_in_flight=aug is not None and _epilogue_outgrows_device(dev, B, n_mels, out.shape[2]))
            _LAST_CALL[(dev.index, stream)] = hist
            if independent:
                flags |= _lib.WFT_LAUNCH_OVERLAP
        args = _lib.FrontendArgs(
            pcm=pcm.data_ptr(),
            pcm_dtype=_lib.WFT_PCM_F32 if pcm.dtype == torch.float32 else _lib.WFT_PCM_I16,
            batch=B,
            clip_stride=pcm.stride(0) if B > 1 else max(pcm.stride(0), N),
            n_samples=N,
            padding=padding,
            lengths=_ptr(lengths),
            n_mels=n_mels,
            n_frames_out=out.shape[2],
            n_valid_frames=_ptr(n_valid_frames),
            mask_params=_ptr(mask_params),
            mask_value=mask_value,
            out=out.data_ptr(),
            workspace=ws.data_ptr(),
            workspace_bytes=need.value,
            workspace_mode=mode,
            launch_flags=flags,
        )
        if draw is not None:
            args.draw_masks = 1
            args.draw_seed = draw[0] & (2**64 - 1)
            args.draw_clip_offset = draw[1] & (2**64 - 1)
            args.draw_time_mask_param, args.draw_freq_mask_param, args.draw_p = int(draw[2]), int(draw[3]), float(draw[4])
        if aug is not None:
            rc = lib.wft_frontend_augment_forward(ctypes.byref(args), ctypes.byref(aug[0]), stream)
        else:
            rc = lib.wft_frontend_forward(ctypes.byref(args), stream)
        if rc != 0 and key is not None:
            _WORKSPACES.pop(key, None)   # a launch that did not happen leaves the ring bookkeeping undefined
        _lib.check(rc)


def _record_call(hist, reads, writes, may_overlap: bool, bounds_in_flight: bool):
    """The launch decision for a front-end call on a stream whose calls since the last waiting launch are ``hist`` (a list of
    ``(reads, writes)`` byte ranges, or None) -> ``(independent, new_hist)``.

    The call may be launched as an independent batch (it will not wait for anything in front of it) only if it writes nothing
    any call in ``hist`` reads or writes and reads nothing any of them writes; otherwise it waits, and everything in front of
    a waiting launch is complete by the time it runs, so the history starts again with this call.  ``bounds_in_flight``: this
    call's epilogue grid has more CTAs than the device can hold.  The grid behind it may only be scheduled once every
    epilogue CTA has started, so some of them will have finished by then, i.e. got past their wait for this call's front-end
    grid -- and grids of a stream complete in order: nothing older than this call can be in flight next to a later launch."""
    # (plain loops over the few byte ranges a call has: this runs once per batch on the host, next to an 80 us kernel)
    reads = [r for r in reads if r is not None]
    writes = [w for w in writes if w is not None]
    independent = bool(may_overlap and hist)
    if independent:
        for prev_reads, prev_writes in hist:
            for lo, hi in writes:
                for qlo, qhi in prev_writes:
                    if lo < qhi and qlo < hi:
                        independent = False
                for qlo, qhi in prev_reads:
                    if lo < qhi and qlo < hi:
                        independent = False
            for lo, hi in reads:
                for qlo, qhi in prev_writes:
                    if lo < qhi and qlo < hi:
                        independent = False
            if not independent:
                break
    hist = (hist + [(reads, writes)]) if independent else [(reads, writes)]
    if bounds_in_flight:
        hist = hist[-1:]
    return independent, hist


_SM_COUNT = {}


def _epilogue_outgrows_device(dev: torch.device, batch: int, n_rows: int, n_frames: int) -> bool:
    """More CTAs in the augmentation epilogue's grid than can be resident at once?  (Either instance of the kernel covers at
    most 1024 frames x 16 rows with a 256-thread CTA; 2048 threads per SM bound the residents whatever else limits them.)"""
    sms = _SM_COUNT.get(dev.index)
    if sms is None:
        sms = _SM_COUNT[dev.index] = torch.cuda.get_device_properties(dev).multi_processor_count
    return ((n_frames + 1023) // 1024) * ((n_rows + 15) // 16) * batch > sms * 8


def _frames(pcm: Tensor, padding: int, n_frames_out: int) -> int:
    return n_frames_out if n_frames_out > 0 else (pcm.shape[1] + padding) // HOP_LENGTH


def _check_frontend_inputs(pcm, n_mels, lengths, n_valid_frames, mask_params):
    if n_mels not in (80, 128):
        raise ValueError(f"Unsupported n_mels: {n_mels}")
    if pcm.dim() != 2 or pcm.stride(1) != 1 or pcm.dtype not in (torch.float32, torch.int16) or pcm.shape[0] < 1:
        raise ValueError("pcm must be a non-empty float32 / int16 tensor of shape [B, N] with unit stride along N")
    B = pcm.shape[0]
    for t, name, shape in ((lengths, "lengths", (B,)), (n_valid_frames, "n_valid_frames", (B,)), (mask_params, "mask_params", (B, 4))):
        if t is not None and (t.dtype != torch.int32 or tuple(t.shape) != shape or not t.is_contiguous() or t.device != pcm.device):
            raise ValueError(f"{name} must be a contiguous int32 tensor of shape {shape} on {pcm.device}")


@torch.library.custom_op("wft::frontend_forward", mutates_args=(), device_types="cuda")
def frontend_forward(pcm: Tensor, n_mels: int, padding: int, lengths: Optional[Tensor], n_frames_out: int,
                     n_valid_frames: Optional[Tensor], mask_params: Optional[Tensor], mask_value: float) -> Tensor:
    _check_frontend_inputs(pcm, n_mels, lengths, n_valid_frames, mask_params)
    out = torch.empty((pcm.shape[0], n_mels, _frames(pcm, padding, n_frames_out)), dtype=torch.float32, device=pcm.device)
    _launch_frontend(pcm, n_mels, padding, lengths, n_frames_out, n_valid_frames, mask_params, mask_value, out)
    return out


@frontend_forward.register_fake
def _(pcm, n_mels, padding, lengths, n_frames_out, n_valid_frames, mask_params, mask_value):
    return pcm.new_empty((pcm.shape[0], n_mels, _frames(pcm, padding, n_frames_out)), dtype=torch.float32)


@torch.library.custom_op("wft::frontend_forward_out", mutates_args=("out",), device_types="cuda")
def frontend_forward_out(pcm: Tensor, n_mels: int, padding: int, lengths: Optional[Tensor], n_frames_out: int,
                         n_valid_frames: Optional[Tensor], mask_params: Optional[Tensor], mask_value: float, out: Tensor) -> None:
    _check_frontend_inputs(pcm, n_mels, lengths, n_valid_frames, mask_params)
    want = (pcm.shape[0], n_mels, _frames(pcm, padding, n_frames_out))
    if out.dtype != torch.float32 or tuple(out.shape) != want or not out.is_contiguous() or out.device != pcm.device:
        raise ValueError(f"out must be a contiguous CUDA float32 tensor of shape {want}")
    _launch_frontend(pcm, n_mels, padding, lengths, n_frames_out, n_valid_frames, mask_params, mask_value, out)


@frontend_forward_out.register_fake
def _(pcm, n_mels, padding, lengths, n_frames_out, n_valid_frames, mask_params, mask_value, out):
    return None


@torch.library.custom_op("wft::frontend_forward_drawn_out", mutates_args=("out",), device_types="cuda")
def frontend_forward_drawn_out(pcm: Tensor, n_mels: int, padding: int, lengths: Optional[Tensor], n_frames_out: int,
                               n_valid_frames: Optional[Tensor], seed: int, clip_offset: int, time_mask_param: int,
                               freq_mask_param: int, p: float, mask_value: float, out: Tensor) -> None:
    """The augmented batch as ONE call: SpecAugment intervals drawn on the device (Philox keyed by ``(seed, clip_offset + b)``,
    same draw as ``wft::specaug_draw``) and the fused front end, two launches behind one trip through the dispatcher."""
    _check_frontend_inputs(pcm, n_mels, lengths, n_valid_frames, None)
    if not 0.0 <= p <= 1.0:
        raise ValueError(f"spec_augment p must be between 0 and 1, got {p}")
    want = (pcm.shape[0], n_mels, _frames(pcm, padding, n_frames_out))
    if out.dtype != torch.float32 or tuple(out.shape) != want or not out.is_contiguous() or out.device != pcm.device:
        raise ValueError(f"out must be a contiguous CUDA float32 tensor of shape {want}")
    _launch_frontend(pcm, n_mels, padding, lengths, n_frames_out, n_valid_frames, None, mask_value, out,
                     draw=(seed, clip_offset, time_mask_param, freq_mask_param, p))


@frontend_forward_drawn_out.register_fake
def _(pcm, n_mels, padding, lengths, n_frames_out, n_valid_frames, seed, clip_offset, time_mask_param, freq_mask_param, p,
      mask_value, out):
    return None


@torch.library.custom_op("wft::frontend_augment_drawn_out", mutates_args=("scratch", "out"), device_types="cuda")
def frontend_augment_drawn_out(pcm: Tensor, n_mels: int, padding: int, lengths: Optional[Tensor], n_frames_out: int,
                               n_valid_frames: Optional[Tensor], seed: int, clip_offset: int, time_mask_param: int,
                               freq_mask_param: int, time_warp_w: int, p: float, extremes: Optional[Tensor], mask_value: float,
                               spline_f32: bool, scratch: Tensor, out: Tensor) -> None:
    """A production batch (every reference config time-warps) as ONE call and TWO grids: the front-end grid writes the
    un-augmented features to ``scratch``, the augmentation epilogue behind it finishes every cell as it loads it (floor, pad,
    silent tiles: what the fix-up grid would have rewritten) and writes warp -> masks -> extremes mask to ``out``; warp point
    and intervals are drawn inside the kernel for ``(seed, clip_offset + b)``.  Bit-identical to ``frontend_forward_out`` +
    ``augment_drawn_out``."""
    _check_frontend_inputs(pcm, n_mels, lengths, n_valid_frames, None)
    if not 0.0 <= p <= 1.0:
        raise ValueError(f"spec_augment p must be between 0 and 1, got {p}")
    B = pcm.shape[0]
    want = (B, n_mels, _frames(pcm, padding, n_frames_out))
    for t, name in ((scratch, "scratch"), (out, "out")):
        if t.dtype != torch.float32 or tuple(t.shape) != want or not t.is_contiguous() or t.device != pcm.device:
            raise ValueError(f"{name} must be a contiguous CUDA float32 tensor of shape {want}")
    if extremes is not None and (extremes.dtype != torch.int32 or tuple(extremes.shape) != (B, 2) or not extremes.is_contiguous()
                                 or extremes.device != pcm.device):
        raise ValueError(f"extremes must be a contiguous int32 tensor of shape {(B, 2)} on {pcm.device}")
    g = _lib.AugmentArgs(out=out.data_ptr(), warp_params=None, mask_params=None, extremes=_ptr(extremes),
                         mask_value=float(mask_value), spline_f32=1 if spline_f32 else 0, draw=1,
                         draw_time_mask_param=int(time_mask_param), draw_freq_mask_param=int(freq_mask_param),
                         draw_time_warp_w=int(time_warp_w), draw_p=float(p), draw_seed=seed & (2**64 - 1),
                         draw_clip_offset=clip_offset & (2**64 - 1))
    _launch_frontend(pcm, n_mels, padding, lengths, n_frames_out, n_valid_frames, None, mask_value, scratch,
                     aug=(g, [extremes], out))


@frontend_augment_drawn_out.register_fake
def _(pcm, n_mels, padding, lengths, n_frames_out, n_valid_frames, seed, clip_offset, time_mask_param, freq_mask_param,
      time_warp_w, p, extremes, mask_value, spline_f32, scratch, out):
    return None


@torch.library.custom_op("wft::pad_or_trim", mutates_args=(), device_types="cuda")
def pad_or_trim(x: Tensor, length: int) -> Tensor:
    """``x`` is ``[outer, len_in, inner]`` float32 contiguous -> ``[outer, length, inner]`` (min-value pad or trim)."""
    lib = _lib.load()
    if x.dim() != 3 or x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("x must be a contiguous float32 tensor of shape [outer, len_in, inner]")
    outer, len_in, inner = x.shape
    if length > len_in and x.numel() == 0:
        raise RuntimeError("pad_or_trim: min(): cannot take the minimum of an empty tensor")
    out = torch.empty((outer, length, inner), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        scratch = torch.empty(16, dtype=torch.uint8, device=x.device)
        _lib.check(lib.wft_pad_or_trim_f32(x.data_ptr(), outer, len_in, inner, length, out.data_ptr(), scratch.data_ptr(),
                                           _stream(x.device)))
    return out


@pad_or_trim.register_fake
def _(x, length):
    return x.new_empty((x.shape[0], length, x.shape[2]))


def _launch_augment(mel: Tensor, warp_params: Optional[Tensor], mask_params: Optional[Tensor], extremes: Optional[Tensor],
                    mask_value: float, spline_f32: bool, out: Tensor) -> None:
    lib = _lib.load()
    if mel.dim() != 3 or mel.dtype != torch.float32 or not mel.is_contiguous():
        raise ValueError("mel must be a contiguous CUDA float32 tensor of shape [B, R, T]")
    B, R, T = mel.shape
    for t, name, cols in ((warp_params, "warp_params", 2), (mask_params, "mask_params", 4), (extremes, "extremes", 2)):
        if t is not None and (t.dtype != torch.int32 or tuple(t.shape) != (B, cols) or not t.is_contiguous() or t.device != mel.device):
            raise ValueError(f"{name} must be a contiguous int32 tensor of shape {(B, cols)} on {mel.device}")
    if out.dtype != torch.float32 or tuple(out.shape) != (B, R, T) or not out.is_contiguous() or out.device != mel.device:
        raise ValueError(f"out must be a contiguous CUDA float32 tensor of shape {(B, R, T)}")
    with torch.cuda.device(mel.device):
        _lib.check(lib.wft_augment_f32(mel.data_ptr(), out.data_ptr(), B, R, T, _ptr(warp_params), _ptr(mask_params),
                                       _ptr(extremes), float(mask_value), 1 if spline_f32 else 0, _stream(mel.device)))


def _chain_stream(dev: torch.device, reads, writes) -> int:
    """Stream handle for the epilogue of a front-end call: the call's recorded byte ranges grow by what the epilogue touches,
    so that the NEXT front-end call may still be launched as an independent batch when it stays clear of all of them."""
    st = torch.cuda.current_stream(dev).cuda_stream
    hist = _LAST_CALL.get((dev.index, st))
    if hist:
        hist[-1] = (hist[-1][0] + [_span(t) for t in reads], hist[-1][1] + [_span(t) for t in writes])
    return st


@torch.library.custom_op("wft::augment_drawn_out", mutates_args=("out",), device_types="cuda")
def augment_drawn_out(mel: Tensor, seed: int, clip_offset: int, time_mask_param: int, freq_mask_param: int, time_warp_w: int,
                      p: float, extremes: Optional[Tensor], mask_value: float, spline_f32: bool, out: Tensor) -> None:
    """The augmentation epilogue with the clip parameters drawn inside the kernel (``wft_augment_drawn_f32``): the same draws as
    ``wft::specaug_draw`` + ``wft::time_warp_draw`` for ``(seed, clip_offset + b)``, no launches in front."""
    lib = _lib.load()
    if mel.dim() != 3 or mel.dtype != torch.float32 or not mel.is_contiguous():
        raise ValueError("mel must be a contiguous CUDA float32 tensor of shape [B, R, T]")
    B, R, T = mel.shape
    if extremes is not None and (extremes.dtype != torch.int32 or tuple(extremes.shape) != (B, 2) or not extremes.is_contiguous()
                                 or extremes.device != mel.device):
        raise ValueError(f"extremes must be a contiguous int32 tensor of shape {(B, 2)} on {mel.device}")
    if out.dtype != torch.float32 or tuple(out.shape) != (B, R, T) or not out.is_contiguous() or out.device != mel.device:
        raise ValueError(f"out must be a contiguous CUDA float32 tensor of shape {(B, R, T)}")
    if not 0.0 <= p <= 1.0:
        raise ValueError(f"spec_augment p must be between 0 and 1, got {p}")
    with torch.cuda.device(mel.device):
        _lib.check(lib.wft_augment_drawn_f32(mel.data_ptr(), out.data_ptr(), B, R, T, seed & (2**64 - 1), clip_offset & (2**64 - 1),
                                             int(time_mask_param), int(freq_mask_param), int(time_warp_w), float(p), _ptr(extremes),
                                             float(mask_value), 1 if spline_f32 else 0,
                                             _chain_stream(mel.device, [mel, extremes], [out])))


@augment_drawn_out.register_fake
def _(mel, seed, clip_offset, time_mask_param, freq_mask_param, time_warp_w, p, extremes, mask_value, spline_f32, out):
    return None


@torch.library.custom_op("wft::augment", mutates_args=(), device_types="cuda")
def augment(mel: Tensor, warp_params: Optional[Tensor], mask_params: Optional[Tensor], extremes: Optional[Tensor],
            mask_value: float, spline_f32: bool) -> Tensor:
    out = torch.empty_like(mel)
    _launch_augment(mel, warp_params, mask_params, extremes, mask_value, spline_f32, out)
    return out


@augment.register_fake
def _(mel, warp_params, mask_params, extremes, mask_value, spline_f32):
    return torch.empty_like(mel)


@torch.library.custom_op("wft::augment_out", mutates_args=("out",), device_types="cuda")
def augment_out(mel: Tensor, warp_params: Optional[Tensor], mask_params: Optional[Tensor], extremes: Optional[Tensor],
                mask_value: float, spline_f32: bool, out: Tensor) -> None:
    _launch_augment(mel, warp_params, mask_params, extremes, mask_value, spline_f32, out)


@augment_out.register_fake
def _(mel, warp_params, mask_params, extremes, mask_value, spline_f32, out):
    return None


@torch.library.custom_op("wft::augment_", mutates_args=("mel",), device_types="cuda")
def augment_(mel: Tensor, mask_params: Optional[Tensor], extremes: Optional[Tensor], mask_value: float) -> None:
    """In place (no warp): the masks only overwrite cells."""
    _launch_augment(mel, None, mask_params, extremes, mask_value, False, mel)


@augment_.register_fake
def _(mel, mask_params, extremes, mask_value):
    return None


def _launch_specaug_apply(mel: Tensor, mask_params: Tensor, mask_value: float, out: Tensor) -> None:
    lib = _lib.load()
    if mel.dim() != 3 or mel.dtype != torch.float32 or not mel.is_contiguous():
        raise ValueError("mel must be a contiguous CUDA float32 tensor of shape [B, R, T]")
    B, R, T = mel.shape
    if mask_params.dtype != torch.int32 or tuple(mask_params.shape) != (B, 4) or not mask_params.is_contiguous():
        raise ValueError(f"mask_params must have shape {(B, 4)}")
    with torch.cuda.device(mel.device):
        _lib.check(lib.wft_specaug_apply_f32(mel.data_ptr(), out.data_ptr(), B, R, T, mask_params.data_ptr(), float(mask_value),
                                             _stream(mel.device)))


@torch.library.custom_op("wft::specaug_apply", mutates_args=(), device_types="cuda")
def specaug_apply(mel: Tensor, mask_params: Tensor, mask_value: float) -> Tensor:
    out = torch.empty_like(mel)
    _launch_specaug_apply(mel, mask_params, mask_value, out)
    return out


@specaug_apply.register_fake
def _(mel, mask_params, mask_value):
    return torch.empty_like(mel)


@torch.library.custom_op("wft::specaug_apply_", mutates_args=("mel",), device_types="cuda")
def specaug_apply_(mel: Tensor, mask_params: Tensor, mask_value: float) -> None:
    _launch_specaug_apply(mel, mask_params, mask_value, mel)


@specaug_apply_.register_fake
def _(mel, mask_params, mask_value):
    return None


@torch.library.custom_op("wft::specaug_draw", mutates_args=(), device_types="cuda")
def specaug_draw(like: Tensor, seed: int, clip_offset: int, batch: int, n_mels: int, n_frames: int, time_mask_param: int,
                 freq_mask_param: int, p: float) -> Tensor:
    """``like`` only names the device (custom ops take their device from a tensor argument)."""
    lib = _lib.load()
    out = torch.empty((batch, 4), dtype=torch.int32, device=like.device)
    with torch.cuda.device(like.device):
        _lib.check(lib.wft_specaug_draw(ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_uint64(clip_offset & (2**64 - 1)), batch,
                                        n_mels, n_frames, time_mask_param, freq_mask_param, float(p), out.data_ptr(),
                                        _stream(like.device)))
    return out


@specaug_draw.register_fake
def _(like, seed, clip_offset, batch, n_mels, n_frames, time_mask_param, freq_mask_param, p):
    return like.new_empty((batch, 4), dtype=torch.int32)


@torch.library.custom_op("wft::time_warp_draw", mutates_args=(), device_types="cuda")
def time_warp_draw(like: Tensor, seed: int, clip_offset: int, batch: int, n_frames: int, time_warp_w: int, p: float) -> Tensor:
    lib = _lib.load()
    out = torch.empty((batch, 2), dtype=torch.int32, device=like.device)
    with torch.cuda.device(like.device):
        _lib.check(lib.wft_time_warp_draw(ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_uint64(clip_offset & (2**64 - 1)), batch,
                                          n_frames, time_warp_w, float(p), out.data_ptr(), _stream(like.device)))
    return out


@time_warp_draw.register_fake
def _(like, seed, clip_offset, batch, n_frames, time_warp_w, p):
    return like.new_empty((batch, 2), dtype=torch.int32)


@torch.library.custom_op("wft::mask_bsd", mutates_args=(), device_types="cuda")
def mask_bsd(x: Tensor, t0: int, t1: int, f0: int, f1: int) -> Tensor:
    """``x`` ``[batch, seq, dim]`` (fp32 / fp16 / bf16): ``out[b, s, d] = 0`` for ``s`` in ``[t0, t1)`` or ``d`` in ``[f0, f1)``."""
    lib = _lib.load()
    if x.dtype not in _ELEM_BYTES:
        raise TypeError(f"activations must be float32, float16 or bfloat16, got {x.dtype}")
    if x.dim() != 3:
        raise ValueError(f"activations must be [batch, seq, dim], got shape {tuple(x.shape)}")
    x = x.contiguous()
    B, S, D = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.wft_mask_bsd(x.data_ptr(), out.data_ptr(), _ELEM_BYTES[x.dtype], B, S, D, int(t0), int(t1), int(f0),
                                    int(f1), 0, _stream(x.device)))
    return out


@mask_bsd.register_fake
def _(x, t0, t1, f0, f1):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def _mask_bsd_setup(ctx, inputs, output):
    _, ctx.t0, ctx.t1, ctx.f0, ctx.f1 = inputs


def _mask_bsd_backward(ctx, grad):
    # y = x * m with m in {0, 1}: the gradient is masked by the same spans
    return torch.ops.wft.mask_bsd(grad, ctx.t0, ctx.t1, ctx.f0, ctx.f1), None, None, None, None


mask_bsd.register_autograd(_mask_bsd_backward, setup_context=_mask_bsd_setup)
