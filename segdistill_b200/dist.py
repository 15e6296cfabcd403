"""Multi-GPU plumbing for the loss path: the batch shards across ranks, every rank computes its
samples' loss and gradient locally (rows never span samples), and the only collective is ONE
all-reduce of the packed loss scalars for logging.

Reference behaviour (``mmseg/models/segmentors/SD_structure.py:110-144``): one ``all_reduce``
plus one blocking ``.item()`` PER log variable per iteration.  Here: one packed fp32 vector, one
all-reduce (NCCL over NVLink on the B200 box, gloo in the CPU tests), one device->host copy.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous batch shard [lo, hi) of ``rank``; earlier ranks take the remainder."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def parse_losses(losses, group=None):
    """``(total_loss, log_vars)`` with the reference's semantics, using one packed all-reduce.

    total = sum of every entry whose key contains 'loss' (SD_structure.py:121-122); each log
    variable is averaged over ranks (:137-142).
    """
    log_vars = OrderedDict()
    for name, value in losses.items():
        if isinstance(value, torch.Tensor):
            log_vars[name] = value.mean()
        elif isinstance(value, list):
            log_vars[name] = sum(v.mean() for v in value)
        else:
            raise TypeError(f'{name} is not a tensor or list of tensors')
    total = sum(v for k, v in log_vars.items() if 'loss' in k)
    log_vars['loss'] = total
    packed = torch.stack([v.detach().float() for v in log_vars.values()])
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        packed = packed / world
        dist.all_reduce(packed, group=group)
    host = packed.tolist()             # the single device->host synchronisation of the step
    return total, OrderedDict(zip(log_vars.keys(), host))
