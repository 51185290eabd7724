import sys, time, torch
sys.path.insert(0, '/root/repo')
import whisper_finetune_b200 as w
torch.cuda.set_device(0)
for dtype in (torch.float32, torch.int16):
  for B, nm in ((64,128),(256,128),(64,80)):
    pcm = (0.1*torch.randn(B, 480000, device='cuda')).clamp(-1,1)
    if dtype == torch.int16: pcm = (pcm*32767).round().to(torch.int16)
    out = torch.empty(B, nm, 3000, device='cuda')
    for _ in range(3): w.frontend_forward(pcm, nm, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): w.frontend_forward(pcm, nm, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/n
    byts = B*(480000*pcm.element_size() + nm*3000*4)
    print(f"{dtype} B={B} n_mels={nm}: {ms*1e3:.1f} us/step  {B/ms*1e3:.0f} clips/s  {ms*1e3/B:.3f} us/clip  {byts/ms/1e6:.0f} GB/s")
