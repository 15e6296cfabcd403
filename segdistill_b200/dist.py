"""Multi-GPU plumbing for the loss path: the batch shards across ranks, every rank computes its
samples' loss and gradient locally (rows never span samples), and the only collective is ONE
all-reduce of the packed loss scalars for logging.

Reference behaviour (``mmseg/models/segmentors/SD_structure.py:110-144``): one ``all_reduce``
plus one blocking ``.item()`` PER log variable per iteration.  Here: one packed fp32 vector, one
all-reduce (NCCL over NVLink on the B200 box, gloo in the CPU tests), one device->host copy.
"""
from __future__ import annotations

import os
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


def _as_vector(values):
    """A list of 0-dim fp32 tensors as one vector: a view when they already sit next to each other in one buffer
    (the outputs of a fused two-loss launch do), else a stacked copy."""
    v0 = values[0]
    try:
        base = v0.untyped_storage().data_ptr()
        adjacent = all(isinstance(v, torch.Tensor) and v.dim() == 0 and v.dtype == torch.float32 and
                       v.untyped_storage().data_ptr() == base and v.storage_offset() == v0.storage_offset() + i
                       for i, v in enumerate(values))
    except Exception:
        adjacent = False
    if adjacent:
        return v0.detach().as_strided((len(values),), (1,))
    return torch.stack([v.detach().float().reshape(()) for v in values])


class DeferredLogs:
    """The loss scalars of the last ``interval`` steps, kept on the device; ONE all-reduce and ONE device->host
    copy per ``interval`` steps instead of one collective (and one blocking ``.item()``) per log variable per step.

    The reference reduces and reads every log variable on every iteration (``SD_structure.py:137-142``) although
    its logger consumes them every 50 (``local_configs/_base_/default_runtime.py:2-7``, mmcv ``TextLoggerHook``
    averages the buffered values of the interval).  Here a step only appends its scalars to a device-resident ring
    - on the GPU one 32-thread launch (``sd_log_push``) that a CUDA graph captures along with the loss kernels, so
    no NCCL kernel sits between two steps' loss kernels - and ``flush()`` returns, for every step since the last
    flush, the same rank-averaged values ``parse_losses`` would have produced then.  The differentiable total never
    needs a collective (every rank back-propagates its own shard's loss, as under the reference's DDP).
    """

    def __init__(self, names, interval=50, device=None, group=None):
        self.names = list(names)
        self.interval = int(interval)
        self.group = group
        self.device = torch.device('cpu') if device is None else torch.device(device)
        # one allocation: the ring's slots*n floats followed by the cursor (an int32 in the last word), so that a flush
        # reads both back with ONE device->host copy
        n = len(self.names)
        self._buf = torch.zeros(self.interval * n + 1, dtype=torch.float32, device=self.device)
        self.ring = self._buf[:self.interval * n].view(self.interval, n)
        self.cursor = self._buf[self.interval * n:].view(torch.int32)
        self._flushed = 0                      # value of the cursor at the last flush
        self._pool = []                        # pinned landing buffers of the asynchronous flushes

    def push(self, values, stream=None, in_backward=False):
        """Append one step's scalars (a tensor of ``len(names)`` fp32 values, or a list of 0-dim tensors) - no
        synchronisation, no collective.  Inside a CUDA-graph capture this records the append into the graph.

        ``in_backward`` (CUDA only): do not launch anything now; the append rides on the scaling launch of the step's
        backward (``sd_scale_grad_log``: the launch every backward through a loss node makes anyway), so the step holds
        no launch and no stream fork for its logging.  ``join()`` - call it after ``backward()`` - appends with a launch
        of its own if no backward took the append (a step without gradients, a grouped or a two-factor backward).

        ``stream``: a side stream for the append (CUDA only).  The append is ordered behind the work queued so far on the
        current stream and the caller joins with ``join()`` later - typically after ``backward()``, so that the
        one-warp launch runs next to the backward's kernels instead of between two steps' loss kernels."""
        if not isinstance(values, torch.Tensor):
            values = _as_vector(values)
        values = values.detach()
        if values.is_cuda:
            from . import _cabi
            if in_backward:
                if _cabi.pending_log is not None:
                    self.join()                                   # (an append nobody took: it must keep its place in the ring)
                _cabi.pending_log = (values.float().contiguous(), self.ring, self.cursor)
                return
            if stream is not None:
                stream.wait_stream(torch.cuda.current_stream(values.device))
                with torch.cuda.stream(stream):
                    _cabi.log_push(values.float().contiguous(), self.ring, self.cursor)
                self._side = stream
            else:
                _cabi.log_push(values.float().contiguous(), self.ring, self.cursor)
        else:
            slot = int(self.cursor.item()) % self.interval
            self.ring[slot].copy_(values.float())
            self.cursor += 1

    def join(self):
        """Order the current stream behind an append that ``push(..., stream=side)`` queued on a side stream."""
        side = getattr(self, '_side', None)
        if side is not None:
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._side = None
        if self.device.type == 'cuda':
            from . import _cabi
            log = _cabi.pending_log
            if log is not None and log[1] is self.ring:           # push(in_backward=True) that no backward took
                _cabi.pending_log = None
                _cabi.log_push(log[0], self.ring, self.cursor)

    def flush_start(self):
        """Enqueue the flush without waiting for it: one all-reduce of the ring (mean over ranks), then ONE asynchronous
        device->host copy of ring + cursor into a pinned buffer.  Returns a handle for ``flush_finish``.
        The launch stream is ordered behind the collective, so a CUDA event recorded after this call times it.

        On NCCL the ring is averaged in place (``ReduceOp.AVG``: no clone, no divide - every row a flush reduces is
        either read by that flush or overwritten before a later one reads it; rows reduced a second time hold the same
        value on every rank already).  Other backends (gloo has no AVG) reduce a copy."""
        self.join()
        distributed = dist.is_available() and dist.is_initialized()
        world = dist.get_world_size(self.group) if distributed else 1
        if (distributed and world > 1 and self._buf.is_cuda and dist.get_backend(self.group) == 'nccl'
                and os.environ.get('SD_LOG_FLUSH_COPY', '0') != '1'):
            dist.all_reduce(self.ring, op=dist.ReduceOp.AVG, group=self.group)
            buf = self._buf
        elif distributed and world > 1:
            buf = self._buf.clone()
            ring = buf[:-1]
            ring /= world
            dist.all_reduce(ring, group=self.group)
        else:                              # nothing to reduce: the copy below is ordered on the stream (CUDA)
            buf = self._buf if self._buf.is_cuda else self._buf.clone()
        if buf.is_cuda:
            host = self._pool.pop() if self._pool else torch.empty(buf.shape, dtype=buf.dtype).pin_memory()
            host.copy_(buf, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            return (host, done)
        return (buf, None)

    def flush_finish(self, handle):
        """Wait for a started flush and return, oldest first, one ``OrderedDict(name -> float)`` per step pushed since
        the previous flush (at most ``interval``); each gets ``'loss'`` = the sum of its entries whose name contains
        'loss' unless a variable of that name was pushed."""
        buf, done = handle
        if done is not None:
            done.synchronize()             # the one synchronisation of the interval
        cur = int(buf[-1:].view(torch.int32).item())
        host = buf[:-1].view(self.interval, len(self.names)).tolist()
        if done is not None:
            self._pool.append(buf)         # pinned landing buffer, reused by a later flush
        n = min((cur - self._flushed) & 0x7fffffff, self.interval)
        self._flushed = cur
        out = []
        for k in range(cur - n, cur):
            row = host[k % self.interval]
            rec = OrderedDict(zip(self.names, row))
            if 'loss' not in rec:
                rec['loss'] = sum(v for name, v in rec.items() if 'loss' in name)
            out.append(rec)
        return out

    def flush(self):
        """``flush_finish(flush_start())``: all-reduce and read back the steps pushed since the last flush."""
        return self.flush_finish(self.flush_start())
