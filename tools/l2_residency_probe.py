#!/usr/bin/env python
"""DRAM bytes of every grid of a production batch (front end -> fix-up -> augmentation epilogue) with the L2 in the state the
previous grid left it in: run under

    ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,\
lts__t_sector_hit_rate.pct --csv --log-file gpurun_out/l2_residency.csv python tools/l2_residency_probe.py

(few metrics = one pass per kernel = no replay, so a kernel sees what its predecessor left in the 126 MB L2).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whisper_finetune_b200 as wft  # noqa: E402

B, NM, T = 64, 128, 3000
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
g = torch.Generator().manual_seed(1)
pcm = [(0.1 * torch.randn(B, 480000, generator=g)).clamp_(-1, 1).to(dev) for _ in range(4)]
outs = [torch.empty(B, NM, T, device=dev) for _ in range(4)]
fe = wft.FrontEnd(n_mels=NM, device=dev, spec_augment=True, seed=42,
                  spec_augment_params={"time_mask_param": 100, "freq_mask_param": 43, "time_warp_w": 80, "p": 1.0})
wft.set_overlap(True)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    fe(pcm[i % 4], clip_offset=i * B, out=outs[i % 4])
torch.cuda.synchronize()
print("ok")
