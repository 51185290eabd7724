import sys, json, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import whisper_finetune_b200 as w
torch.cuda.set_device(0)
res = []
for dtype in (torch.float32, torch.int16):
  for nm in (128, 80):
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048):
      if nm == 80 and B not in (64, 256): continue
      pcm = (0.1*torch.randn(min(B,256), 480000, device='cuda')).clamp(-1,1)
      if B > 256: pcm = pcm.repeat(B // 256, 1)
      if dtype == torch.int16: pcm = (pcm*32767).round().to(torch.int16)
      out = torch.empty(B, nm, 3000, device='cuda')
      masks = w.draw_mask_params(42, 0, B, nm, 3000, 100, 43, 1.0)
      for _ in range(3): w.frontend_forward(pcm, nm, mask_params=masks, out=out)
      torch.cuda.synchronize()
      n = max(3, min(50, 4096 // B))
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(n): w.frontend_forward(pcm, nm, mask_params=masks, out=out)
      e1.record(); torch.cuda.synchronize()
      ms = e0.elapsed_time(e1)/n
      byts = B*(480000*pcm.element_size() + nm*3000*4)
      r = dict(dtype=str(dtype).split('.')[-1], n_mels=nm, B=B, us=ms*1e3, clips_per_s=B/ms*1e3, us_per_clip=ms*1e3/B, GBps=byts/ms/1e6, frac=byts/ms/1e6/6551.7)
      res.append(r); print(json.dumps(r), flush=True)
      del pcm, out
json.dump(res, open('gpurun_out/sweep_r1.json','w'))
