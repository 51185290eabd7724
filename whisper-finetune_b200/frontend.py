"""Batched front end: what ``AudioDataset.__getitem__`` + ``collate_fn`` produce for a whole batch, in one launch.

``FrontEnd`` carries the knobs of the reference's YAML (``augmentation.spec_augment.{apply,p,time_mask_param,
freq_mask_param}``, read at ``src/whisper_finetune/data/data_loader.py:109-117``; ``n_mels`` comes from the model,
``scripts/finetune.py:644``) and maps a batch of raw PCM to ``x[B, n_mels, 3000]``:

    zero pad to 480000 (data_loader.py:346) -> log-mel (:278) -> partial-segment cut (:279-280) -> min-value
    pad_or_trim (:281-282) -> time mask, frequency mask (:286-287) -> stacked like collate_fn (:362-367)

The masks of clip ``b`` are a pure function of ``(seed, clip_offset + b)``, so a DistributedSampler shard
reproduces them on any number of GPUs.
"""
from typing import Optional

import torch

from .audio import N_FRAMES, N_SAMPLES, frontend_forward, resolve_device
from . import ops as _ops
from .augment import augment_epilogue, draw_mask_params, draw_warp_params


class FrontEnd:
    """``spec_augment_params`` is the reference's ``augmentation.spec_augment`` block passed verbatim
    (``time_mask_param``, ``freq_mask_param``, ``time_warp_w``, ``p``).  Like the reference
    (data_loader.py:117, 284-287) a positive ``time_warp_w`` turns the time-warp ON whenever the gate passes; it then runs
    with the masks (and the extremes mask) as ONE fused epilogue pass after the front-end kernel.  ``"time_warp": False``
    in the params is the explicit opt-out (a single fused launch, masks only); a missing ``time_warp_w`` means no warp."""

    def __init__(self, n_mels: int = 80, device=None, spec_augment: bool = False,
                 spec_augment_params: Optional[dict] = None, seed: int = 0, n_samples: int = N_SAMPLES,
                 n_frames: int = N_FRAMES, warp_spline: str = "f64"):
        if n_mels not in (80, 128):
            raise ValueError(f"Unsupported n_mels: {n_mels}")
        if warp_spline not in ("f32", "f64"):
            raise ValueError("warp_spline must be 'f32' or 'f64'")
        self.n_mels = n_mels
        self.device = resolve_device(device)
        self.seed = int(seed)
        self.n_samples = int(n_samples)
        self.n_frames = int(n_frames)
        self.warp_spline = warp_spline
        self.spec_augment = bool(spec_augment)
        self.spec_augment_p = 0.0
        self.time_mask_param = self.freq_mask_param = self.time_warp_w = 0
        if spec_augment:
            params = spec_augment_params or {}
            self.spec_augment_p = float(params.get("p", 1.0))
            if not 0.0 <= self.spec_augment_p <= 1.0:
                raise ValueError(f"spec_augment p must be between 0 and 1, got {self.spec_augment_p}")
            self.time_mask_param = int(params["time_mask_param"])
            self.freq_mask_param = int(params["freq_mask_param"])
            if params.get("time_warp", True):
                self.time_warp_w = int(params.get("time_warp_w", 0))
        self._scratch = {}        # (stream, batch, turn % K) -> un-warped features of a time-warped batch
        self._scratch_turn = {}

    def __call__(self, pcm: torch.Tensor, lengths=None, n_valid_frames=None, clip_offset: int = 0,
                 mask_params: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 augment: Optional[torch.Tensor] = None, extremes: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``pcm`` ``[B, N<=480000]`` float32 / int16 (host or device) -> ``[B, n_mels, 3000]`` on the device.

        ``augment`` (optional int / bool ``[B]``): per-clip outcome of the reference's SpecAugment gate
        (data_loader.py:294-301) when it was already decided upstream (the loader's workers draw it from their own RNG):
        clips with 0 get no warp and no masks, clips with 1 get them -- the device-side ``p`` gate is then NOT rolled again.
        ``extremes`` (optional int32 ``[B, 2]``): rows masked from the bottom / top by ExtremesFrequencyMasking
        (data_loader.py:289-290), applied whatever the gate says, like the reference."""
        if not torch.is_tensor(pcm):
            pcm = torch.as_tensor(pcm)
        if pcm.dim() != 2:
            raise ValueError("pcm must be [B, N]")
        if pcm.shape[1] > self.n_samples:
            raise ValueError(f"clips longer than {self.n_samples} samples must be chunked upstream")
        pcm = pcm.to(self.device, non_blocking=True)
        B, N = pcm.shape
        if (mask_params is None and augment is None and extremes is None and self.spec_augment and self.spec_augment_p > 0.0
                and self.time_warp_w == 0 and pcm.dtype in (torch.float32, torch.int16) and pcm.stride(1) == 1):
            # the common augmented batch (masks only): draw + fused kernel behind ONE op call
            from .audio import as_i32_on

            if out is None:
                out = torch.empty((B, self.n_mels, self.n_frames), dtype=torch.float32, device=self.device)
            _ops.run_eager(_ops.frontend_forward_drawn_out, pcm, self.n_mels, self.n_samples - N, as_i32_on(lengths, self.device, (B,), "lengths"),
                                                     self.n_frames, as_i32_on(n_valid_frames, self.device, (B,), "n_valid_frames"),
                                                     self.seed, int(clip_offset), self.time_mask_param, self.freq_mask_param,
                                                     self.spec_augment_p, 0.0, out)
            return out
        if (mask_params is None and augment is None and self.spec_augment and self.spec_augment_p > 0.0 and self.time_warp_w > 0
                and pcm.dtype in (torch.float32, torch.int16) and pcm.stride(1) == 1):
            # the production batch (configs/*: time_warp_w = 80) as ONE call: front-end grid -> ONE epilogue grid that finishes the
            # cells as it loads them (floor / pad / silent tiles) and draws the clip's warp point and mask intervals itself.  The
            # un-warped features go through a rotation of scratch buffers, so that the following batches' front-end grids may run
            # under this batch's epilogue (wft.set_overlap: an independent launch has to stay clear of every call that may still
            # be in flight, see ops._LAST_CALL -- two buffers once the epilogue's grid outgrows the device, which bounds what can
            # be in flight to the call in front; up to 16 (1 GiB at most) for the small batches that fit on the GPU side by side)
            from .audio import as_i32_on

            if torch.cuda.is_current_stream_capturing():
                plain = torch.empty((B, self.n_mels, self.n_frames), dtype=torch.float32, device=self.device)   # the graph's own
            else:
                st = torch.cuda.current_stream(self.device).cuda_stream      # scratch is per stream: stream order protects it
                turn = self._scratch_turn.get(st, 0)
                self._scratch_turn[st] = turn + 1
                nbytes = B * self.n_mels * self.n_frames * 4
                from .ops import _epilogue_outgrows_device

                ring = 2 if _epilogue_outgrows_device(self.device, B, self.n_mels, self.n_frames) else max(2, min(16, (1 << 30) // nbytes))
                key = (st, B, turn % ring)
                plain = self._scratch.get(key)
                if plain is None:
                    for k in [k for k in self._scratch if k[0] == st and k[1] != B]:   # batch size changed on this stream
                        del self._scratch[k]
                    plain = self._scratch[key] = torch.empty((B, self.n_mels, self.n_frames), dtype=torch.float32, device=self.device)
            if out is None:
                out = torch.empty((B, self.n_mels, self.n_frames), dtype=torch.float32, device=self.device)
            ext = None if extremes is None else torch.as_tensor(extremes, dtype=torch.int32).to(self.device).contiguous()
            _ops.run_eager(_ops.frontend_augment_drawn_out, pcm, self.n_mels, self.n_samples - N, as_i32_on(lengths, self.device, (B,), "lengths"),
                                                     self.n_frames, as_i32_on(n_valid_frames, self.device, (B,), "n_valid_frames"),
                                                     self.seed, int(clip_offset), self.time_mask_param, self.freq_mask_param,
                                                     self.time_warp_w, self.spec_augment_p, ext, 0.0, self.warp_spline == "f32",
                                                     plain, out)
            return out
        gate = None
        p_draw = self.spec_augment_p
        if augment is not None:
            gate = torch.as_tensor(augment).to(self.device, torch.int32, non_blocking=True).reshape(B, 1)
            p_draw = 1.0 if self.spec_augment else 0.0   # the gate was rolled upstream: do not roll it a second time
        if mask_params is None and self.spec_augment and p_draw > 0.0:
            mask_params = draw_mask_params(self.seed, clip_offset, B, self.n_mels, self.n_frames,
                                           self.time_mask_param, self.freq_mask_param, p_draw, self.device)
        if gate is not None and mask_params is not None:
            mask_params = (torch.as_tensor(mask_params).to(self.device, torch.int32) * gate).contiguous()   # [0, 0) masks nothing
        warps = None
        if self.time_warp_w > 0 and p_draw > 0.0:
            warps = draw_warp_params(self.seed, clip_offset, B, self.n_frames, self.time_warp_w, p_draw, self.device)
            if gate is not None:   # warp_p = -1: the clip is left alone
                warps = torch.where(gate != 0, warps, torch.tensor([-1, 0], dtype=torch.int32, device=self.device)).contiguous()
        if warps is None and extremes is None:
            return frontend_forward(pcm, self.n_mels, padding=self.n_samples - N, lengths=lengths,
                                    n_frames_out=self.n_frames, n_valid_frames=n_valid_frames,
                                    mask_params=mask_params, mask_value=0.0, out=out)
        if warps is None:
            # no warp: the masks ride in the front-end kernel, the extremes mask is an in-place pass over its output
            x = frontend_forward(pcm, self.n_mels, padding=self.n_samples - N, lengths=lengths,
                                 n_frames_out=self.n_frames, n_valid_frames=n_valid_frames,
                                 mask_params=mask_params, mask_value=0.0, out=out)
            return augment_epilogue(x, None, None, extremes, 0.0, out=x)
        # reference order (data_loader.py:285-290): warp -> time mask -> frequency mask -> extremes mask, one epilogue pass
        plain = frontend_forward(pcm, self.n_mels, padding=self.n_samples - N, lengths=lengths,
                                 n_frames_out=self.n_frames, n_valid_frames=n_valid_frames)
        return augment_epilogue(plain, warps, mask_params, extremes, 0.0, out=out, spline=self.warp_spline)


class HostPipeline:
    """End-to-end path for HOST-resident PCM: pinned host batch -> H2D -> fused front end -> D2H features.

    This is the loader-integration shape of SURVEY 8(f)-4: DataLoader workers hand over raw PCM (``int16`` halves
    the H2D bytes) and the features come back where ``collate_fn`` (data_loader.py:362-367) would have put them.
    The batch is cut into ``n_chunks`` slices that travel on ``n_streams`` CUDA streams so that the H2D copy of
    slice i+1, the kernel of slice i and the D2H copy of slice i-1 overlap (PCIe is full duplex).
    """

    def __init__(self, front_end: FrontEnd, batch: int, n_samples: int = N_SAMPLES, pcm_dtype=torch.float32,
                 n_chunks: int = 4, n_streams: int = 2, readback: str = "features"):
        """``readback="features"``: the whole feature tensor returns to the host (what the reference's CPU path yields);
        ``readback="probe"``: features stay in HBM for the model (``train_step`` moves ``x`` to the device anyway,
        model/model_utils.py:60) and only one float per clip comes back as a liveness probe."""
        self.fe = front_end
        self.batch = int(batch)
        self.n_samples = int(n_samples)
        self.n_chunks = max(1, min(int(n_chunks), self.batch))
        dev = front_end.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, n_streams))]
        bounds = torch.linspace(0, self.batch, self.n_chunks + 1).round().long().tolist()
        self.slices = [(bounds[i], bounds[i + 1]) for i in range(self.n_chunks) if bounds[i + 1] > bounds[i]]
        self.dev_pcm = torch.empty((self.batch, self.n_samples), dtype=pcm_dtype, device=dev)
        self.dev_out = torch.empty((self.batch, front_end.n_mels, front_end.n_frames), dtype=torch.float32, device=dev)
        if readback not in ("features", "probe"):
            raise ValueError("readback must be 'features' or 'probe'")
        self.readback = readback
        self.h2d_bytes = self.dev_pcm.numel() * self.dev_pcm.element_size()
        self.d2h_bytes = self.dev_out.numel() * 4 if readback == "features" else self.batch * 4

    def __call__(self, pcm_host: torch.Tensor, out_host: torch.Tensor, lengths=None, n_valid_frames=None,
                 clip_offset: int = 0) -> torch.Tensor:
        """``pcm_host`` pinned ``[B, N]``, ``out_host`` pinned ``[B, n_mels, 3000]`` (``[B]`` for ``readback="probe"``);
        returns ``out_host`` once every copy has been enqueued: call ``synchronize()`` before reading it on the host, or
        ``join()`` to make the current stream wait for it.

        Consecutive calls overlap: slice ``k`` always travels on stream ``k % n_streams`` and re-uses its own part of the
        device buffers, so stream order alone keeps two batches apart and the H2D copies of batch ``i + 1`` start while the
        D2H copies of batch ``i`` are still draining (the current stream is NOT made to wait here -- that would put a
        full pipeline drain between any two batches)."""
        cur = torch.cuda.current_stream(self.fe.device)
        for k, (a, b) in enumerate(self.slices):
            st = self.streams[k % len(self.streams)]
            st.wait_stream(cur)   # whatever the caller queued before this call (device-side lengths, ...)
            with torch.cuda.stream(st):
                self.dev_pcm[a:b].copy_(pcm_host[a:b], non_blocking=True)
                self.fe(self.dev_pcm[a:b],
                        lengths=None if lengths is None else lengths[a:b],
                        n_valid_frames=None if n_valid_frames is None else n_valid_frames[a:b],
                        clip_offset=clip_offset + a, out=self.dev_out[a:b])
                if self.readback == "features":
                    out_host[a:b].copy_(self.dev_out[a:b], non_blocking=True)
                else:
                    out_host[a:b].copy_(self.dev_out[a:b, 0, 0], non_blocking=True)
        return out_host

    def join(self) -> None:
        """Make the current stream wait for everything enqueued so far (for stream-ordered consumers and event timing)."""
        cur = torch.cuda.current_stream(self.fe.device)
        for st in self.streams:
            cur.wait_stream(st)

    def synchronize(self) -> None:
        for st in self.streams:
            st.synchronize()
