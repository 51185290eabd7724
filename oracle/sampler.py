"""ORACLE (test infrastructure, never shipped): which clips a rank owns.

Restates the index partition of ``torch.utils.data.DistributedSampler`` as the reference constructs it at
``/root/reference/src/whisper_finetune/scripts/finetune.py:619-629`` (``shuffle=True``, ``seed``,
``drop_last``) and reseeds it through ``set_epoch`` (``model/model_utils.py:209-217``):
``randperm(n, generator=seed+epoch)`` -> pad by wrap-around or drop the tail to a multiple of the world
size -> ``indices[rank::world]``.  PINNED against the real ``DistributedSampler`` in tests.
"""
import math
from typing import List

import torch


def rank_indices(n: int, world: int, rank: int, epoch: int = 0, seed: int = 0, shuffle: bool = True,
                 drop_last: bool = False) -> List[int]:
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).tolist()
    else:
        idx = list(range(n))
    if drop_last and n % world != 0:
        per_rank = math.ceil((n - world) / world)
    else:
        per_rank = math.ceil(n / world)
    total = per_rank * world
    if not drop_last:
        short = total - len(idx)
        if short <= len(idx):
            idx += idx[:short]
        else:
            idx += (idx * math.ceil(short / len(idx)))[:short]
    else:
        idx = idx[:total]
    return idx[rank:total:world]
