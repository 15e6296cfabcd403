"""CPU restatement of the student head's supervised loss (TEST INFRASTRUCTURE ONLY - never imported by the product
path; used by tests/, __graft_entry__.smoke() and bench.py's baseline legs as the checker).

Follows, line by line:
  * ``BaseDecodeHead.losses``             /root/reference/mmseg/models/decode_heads/decode_head.py:217-237
  * ``cross_entropy`` / ``CrossEntropyLoss``  mmseg/models/losses/cross_entropy_loss.py:9-32, :138-198
  * ``weight_reduce_loss``                 mmseg/models/losses/utils.py:25-56
  * ``accuracy`` (top-1)                   mmseg/models/losses/accuracy.py:4-46
  * ``resize``                             mmseg/ops/wrappers.py:8-29 (= F.interpolate)

Pinned: tests/golden/segloss_*.npz are produced by tests/golden/make_golden.py from the UNMODIFIED reference modules
(cross_entropy_loss.py, accuracy.py, wrappers.py loaded by path, called in the order of decode_head.py:217-237);
tests/test_oracle.py checks this restatement against them.
"""
from __future__ import annotations

import torch.nn.functional as F


def decode_head_losses_torch(seg_logit, seg_label, class_weight=None, loss_weight=1.0, ignore_index=255,
                             align_corners=False, reduction='mean', avg_factor=None, seg_weight=None):
    """{'loss_seg': scalar, 'acc_seg': percent} as ``BaseDecodeHead.losses`` computes them (sampler=None unless
    ``seg_weight`` is given)."""
    x = F.interpolate(seg_logit, size=tuple(seg_label.shape[2:]), mode='bilinear',
                      align_corners=align_corners)                                   # decode_head.py:221-225
    label = seg_label.squeeze(1)                                                       # :230
    cw = None if class_weight is None else x.new_tensor(class_weight)                 # cross_entropy_loss.py:184-187
    loss = F.cross_entropy(x, label, weight=cw, reduction='none', ignore_index=ignore_index)   # :19-24
    if seg_weight is not None:
        loss = loss * seg_weight.float()                                              # utils.py:38-42
    if avg_factor is None:                                                            # utils.py:45-55
        if reduction == 'mean':
            loss = loss.mean()
        elif reduction == 'sum':
            loss = loss.sum()
    elif reduction == 'mean':
        loss = loss.sum() / avg_factor
    elif reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    loss = loss_weight * loss                                                         # cross_entropy_loss.py:189
    pred = x.topk(1, dim=1)[1].transpose(0, 1)                                        # accuracy.py:36-38
    correct = pred.eq(label.unsqueeze(0).expand_as(pred))                             # :39
    acc = correct[:1].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / label.numel())   # :44-45
    return {'loss_seg': loss, 'acc_seg': acc[0]}
