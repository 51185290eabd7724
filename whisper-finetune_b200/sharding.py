"""How clips are partitioned over GPUs: the reference's DistributedSampler, and nothing more.

The reference shards every epoch with ``DistributedSampler(dataset, num_replicas=WORLD_SIZE, rank=RANK,
shuffle=True, seed=seed, drop_last=...)`` (``src/whisper_finetune/scripts/finetune.py:619-629``) and reseeds it
with ``set_epoch`` (``model/model_utils.py:209-217``).  Feature extraction needs no cross-GPU traffic: each rank
runs the front end on ``indices[rank::world]`` and, because SpecAugment intervals are keyed by the GLOBAL clip
index, the union over ranks equals the single-GPU result.  ``all_gather_features`` is the one optional
collective (NCCL on GPUs; gloo in CPU tests).
"""
import math
from typing import List, Optional

import torch
import torch.distributed as dist


def shard_indices(n: int, world_size: int, rank: int, epoch: int = 0, seed: int = 0, shuffle: bool = True,
                  drop_last: bool = False) -> List[int]:
    """Global clip indices owned by ``rank`` -- element for element what DistributedSampler yields."""
    if not 0 <= rank < world_size:
        raise ValueError(f"Invalid rank {rank}, rank should be in the interval [0, {world_size - 1}]")
    if shuffle:
        gen = torch.Generator()
        gen.manual_seed(seed + epoch)
        order = torch.randperm(n, generator=gen).tolist()
    else:
        order = list(range(n))
    if drop_last and n % world_size != 0:
        per_rank = math.ceil((n - world_size) / world_size)
        order = order[: per_rank * world_size]
    else:
        per_rank = math.ceil(n / world_size)
        missing = per_rank * world_size - len(order)
        if missing > 0:
            reps = math.ceil(missing / max(len(order), 1))
            order = order + (order * reps)[:missing]
    return order[rank : per_rank * world_size : world_size]


def all_gather_features(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Concatenate every rank's ``[b, n_mels, T]`` block along dim 0 (equal ``b`` on all ranks).

    Row ``r * b + j`` of the result is clip ``j`` of rank ``r``; callers that want DistributedSampler order back
    interleave with ``out.view(world, b, ...).transpose(0, 1)``.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    local = local.contiguous()
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, local, group=group)
    else:
        chunks = list(out.chunk(world, dim=0))
        dist.all_gather(chunks, local, group=group)
    return out
