import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whisper_finetune_b200 as w
torch.cuda.set_device(0)
B=64
pcm=[(0.1*torch.randn(B,480000,device='cuda')).clamp(-1,1) for _ in range(4)]
outs=[torch.empty(B,128,3000,device='cuda') for _ in range(16)]
masks=w.draw_mask_params(42,0,B,128,3000,100,43,1.0,torch.device('cuda'))
nomask=torch.zeros_like(masks)
w.set_overlap(True)
def run(kw,K=200):
    for i in range(10): w.frontend_forward(pcm[i%4],128,out=outs[i%16],**kw)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): w.frontend_forward(pcm[i%4],128,out=outs[i%16],**kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/K*1e3
for rep in range(2):
    print('no masks      %.2f us'%run({}))
    print('zero masks    %.2f us'%run(dict(mask_params=nomask)))
    print('drawn masks   %.2f us'%run(dict(mask_params=masks)))
    same=[pcm[0]]*4
    pcm_b=pcm; pcm=same
    print('no masks, one PCM set (L2-warm input) %.2f us'%run({}))
    pcm=pcm_b
