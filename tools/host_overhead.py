"""Host time of one front-end call (python -> torch op -> ctypes -> C ABI -> launches): B = 1 so that the GPU is never the
limiter; wall clock per call over 3000 calls, plus a cProfile of where it goes."""
import cProfile, io, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whisper_finetune_b200 as w
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
pcm = [(0.1 * torch.randn(1, 480000, device=dev)).clamp(-1, 1) for _ in range(4)]
outs = [torch.empty(1, 128, 3000, device=dev) for _ in range(16)]
w.set_overlap(True)
for name, params in (("masks", {"time_mask_param": 100, "freq_mask_param": 43, "p": 1.0}),
                     ("masks + time-warp", {"time_mask_param": 100, "freq_mask_param": 43, "time_warp_w": 80, "p": 1.0})):
    fe = w.FrontEnd(n_mels=128, device=dev, spec_augment=True, seed=1, spec_augment_params=params)
    def body(n):
        for i in range(n):
            fe(pcm[i % 4], clip_offset=i, out=outs[i % 16])
    body(200); torch.cuda.synchronize()
    t0 = time.perf_counter(); body(3000); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name}: {(t1 - t0) / 3000 * 1e6:.1f} us of host time per call")
    pr = cProfile.Profile(); pr.enable(); body(1000); pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
