"""The workload run under compute-sanitizer (memcheck / synccheck / racecheck): every kernel of the library, and every path of
the fused kernel and its fix-up grid -- ragged lengths (silent tiles), partial-segment cuts, masks, int16 / 80 mel, capped
grids, one long clip whose floor binds, a config-3 style batch, overlapping independent batches on the workspace ring
(more than one trip round it), intervals drawn inside the kernel, the production call (front-end grid -> staged epilogue that
finishes the cells on load) with its generic instance and under CUDA graph replay, the fused augmentation epilogue, the
draws, pad_or_trim and the activation mask."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whisper_finetune_b200 as w  # noqa: E402

torch.cuda.set_device(0)
lib = w._lib.load()
g = torch.Generator().manual_seed(0)
pcm = (0.1 * torch.randn(3, 480000, generator=g)).clamp(-1, 1).cuda()
lengths = torch.tensor([480000, 100000, 31234], dtype=torch.int32, device="cuda")
nv = torch.tensor([-1, 500, 100], dtype=torch.int32, device="cuda")
masks = w.draw_mask_params(1, 0, 3, 128, 3000, 100, 27, 1.0)
out = w.frontend_forward(pcm, 128, lengths=lengths, n_valid_frames=nv, mask_params=masks, n_frames_out=3000)
out16 = w.frontend_forward((pcm * 32767).round().to(torch.int16), 80, lengths=lengths)
# capped grids: 2 CTAs walk every tile (front-end kernel) and every 32-tile group (fix-up kernel)
lib.wft_debug_set_max_ctas(2)
capped = w.frontend_forward(pcm, 128, lengths=lengths, n_valid_frames=nv, mask_params=masks, n_frames_out=3000)
lib.wft_debug_set_max_ctas(0)
assert torch.equal(capped, out)
# a 2-minute clip: loud second, then faint noise (the floor binds on every later tile), full grid
long = 1e-7 * torch.randn(120 * 16000, generator=g)
long[:16000] += 0.5 * torch.sin(torch.arange(16000) * 0.17)
lm = w.log_mel_spectrogram(long.cuda(), n_mels=128)
# config-3 style batch: 24 ragged clips, cuts, masks, through FrontEnd with the reference's config block (warp + masks)
B = 24
ragged = (0.1 * torch.randn(B, 480000, generator=g)).clamp(-1, 1)
rl = torch.randint(16000, 480001, (B,), generator=g).to(torch.int32)
ragged[torch.arange(480000)[None, :] >= rl[:, None]] = 0
rv = torch.full((B,), -1, dtype=torch.int32)
rv[::4] = torch.randint(2, 3000, (B // 4,), generator=g).to(torch.int32)
fe = w.FrontEnd(n_mels=128, spec_augment=True, seed=3,
                spec_augment_params={"time_mask_param": 100, "freq_mask_param": 27, "time_warp_w": 80, "p": 0.7})
ext = torch.randint(0, 6, (B, 2), generator=g).to(torch.int32)
x = fe(ragged.cuda(), lengths=rl, n_valid_frames=rv, clip_offset=100, extremes=ext)
# independent batches overlapping on the 16-phase workspace ring (WFT_LAUNCH_OVERLAP), 20 calls = more than one trip
old = w.set_overlap(True)
bufs = [torch.empty(3, 128, 3000, device="cuda") for _ in range(3)]
fe_plain = w.FrontEnd(n_mels=128)
fe_draw = w.FrontEnd(n_mels=128, spec_augment=True, seed=9, spec_augment_params={"time_mask_param": 100, "freq_mask_param": 27, "p": 1.0})
lengths_d, nv_d = lengths, nv
for i in range(20):
    if i % 3 == 2:
        fe_draw(pcm, lengths=lengths_d, n_valid_frames=nv_d, clip_offset=3 * i, out=bufs[i % 3])
    else:
        fe_plain(pcm, lengths=lengths_d if i % 3 else None, n_valid_frames=nv_d if i % 3 else None, out=bufs[i % 3])
w.set_overlap(old)
torch.cuda.synchronize()
assert torch.equal(bufs[1], fe_plain(pcm, lengths=lengths_d, n_valid_frames=nv_d))
# the production call on full-length clips (front-end grid -> staged epilogue that finishes the cells on load, floor-only
# instance), the same through the generic epilogue instance, an odd frame count (generic only), and a captured + replayed call
fe_full = w.FrontEnd(n_mels=128, spec_augment=True, seed=4,
                     spec_augment_params={"time_mask_param": 100, "freq_mask_param": 27, "time_warp_w": 80, "p": 1.0})
quiet = pcm.clone()
quiet[1, 100000:] = 0.0                       # the floor binds
y_staged = fe_full(quiet, clip_offset=7)
lib.wft_debug_set_augment_generic(1)
y_generic = fe_full(quiet, clip_offset=7)
lib.wft_debug_set_augment_generic(0)
assert torch.equal(y_staged, y_generic)
odd = w.augment_epilogue(out[:, :, :2999].contiguous(), w.draw_warp_params(1, 0, 3, 2999, 80), masks, None)
static_in, static_out = quiet.clone(), torch.empty_like(y_staged)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    fe_full(static_in, clip_offset=7, out=static_out)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    fe_full(static_in, clip_offset=7, out=static_out)
graph.replay()
graph.replay()
torch.cuda.synchronize()
assert torch.equal(static_out, y_staged)
wp = w.draw_warp_params(1, 0, 3, 3000, 80)
tw = w.time_warp(out, wp)
ep = w.augment_epilogue(out, wp, masks, None, spline="f32")
pt = w.pad_or_trim(out[0, :, :100].contiguous(), 3000)
act = torch.randn(2, 150, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
w.mask_activations(act, (3, 40), (10, 50)).float().sum().backward()
torch.cuda.synchronize()
print("ok", float(out.sum()), float(out16.sum()), float(lm.sum()), float(x.sum()), float(tw.sum()), float(ep.sum()), float(pt.sum()))
