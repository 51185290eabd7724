import sys, torch
sys.path.insert(0, '/root/repo')
import whisper_finetune_b200 as w
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nm = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.cuda.set_device(0)
pcm = (0.1*torch.randn(B, 480000, device='cuda')).clamp(-1,1)
out = torch.empty(B, nm, 3000, device='cuda')
for _ in range(4): w.frontend_forward(pcm, nm, out=out)
torch.cuda.synchronize()
