import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import whisper_finetune_b200 as w
torch.cuda.set_device(0)
rng = np.random.default_rng(42)
B = 256
pcm = (0.1*torch.randn(B, 480000, device='cuda')).clamp(-1,1)
lengths = torch.from_numpy(rng.integers(16000, 480001, size=B).astype(np.int32)).cuda()
nv = np.full(B, -1, dtype=np.int32); nv[::4] = (rng.uniform(0.02, 30.0, size=len(nv[::4]))*100).astype(np.int32)
nv = torch.from_numpy(nv).cuda()
masks = w.draw_mask_params(42, 0, B, 128, 3000, 100, 27, 1.0)
out = torch.empty(B, 128, 3000, device='cuda')
for name, kw in (('full-length', {}), ('ragged lengths', dict(lengths=lengths)), ('ragged + cuts + masks (config 3)', dict(lengths=lengths, n_valid_frames=nv, mask_params=masks))):
    for _ in range(3): w.frontend_forward(pcm, 128, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): w.frontend_forward(pcm, 128, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    print(f"{name}: B={B} {ms*1e3:.1f} us/step  {B/ms*1e3:.0f} clips/s  {ms*1e3/B:.3f} us/clip")
