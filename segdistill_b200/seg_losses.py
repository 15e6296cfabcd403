"""The student head's supervised loss through the fused kernel: bilinear resize to the label size +
cross-entropy + top-1 accuracy, forward and backward in one pass, the resized logits never materialised.

Mirrors ``BaseDecodeHead.losses`` (``mmseg/models/decode_heads/decode_head.py:217-237``) and the reference's
``CrossEntropyLoss`` (``mmseg/models/losses/cross_entropy_loss.py:138-198``: same constructor, same call), for the
softmax cross-entropy the shipped configs use (``use_sigmoid=False, use_mask=False``).  What the kernel does not
cover raises ``SegDistillUnsupported`` (there is no fallback): ``reduction='none'`` (a per-pixel map), the sigmoid and
mask variants.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from .functional import _finish_backward, _keep, once_differentiable

__all__ = ['CrossEntropyLoss', 'decode_head_losses', 'seg_ce_loss']


class _CEUp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label, scale, class_weight, pixel_weight, ignore_index, loss_weight, denominator):
        ctx.call = dict(class_weight=class_weight, pixel_weight=pixel_weight, ignore_index=ignore_index,
                        loss_weight=loss_weight, denominator=denominator)
        ctx.scale = scale
        loss, acc, dx = _cabi.ce_up(logits, label, scale, **ctx.call)
        _keep(ctx, logits, label, dx)
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, _grad_acc):
        return (_finish_backward(ctx, grad_output, lambda x, y: _cabi.ce_up(x, y, ctx.scale, **ctx.call)[2]),) + (None,) * 7


def seg_ce_loss(seg_logit, seg_label, class_weight=None, pixel_weight=None, ignore_index=255, loss_weight=1.0,
                reduction='mean', avg_factor=None, align_corners=False):
    """(loss, acc) of ``seg_logit`` (B, C, h, w) against ``seg_label`` (B, 1, H, W) or (B, H, W): the logits are
    resized bilinearly to (H, W) - inside the kernel for an integer factor 1, 2, 4, 8 with ``align_corners=False``,
    by ``F.interpolate`` otherwise."""
    if reduction not in ('mean', 'sum'):
        raise _cabi.SegDistillUnsupported(f"fused cross-entropy: reduction='{reduction}' is not supported")
    if avg_factor is not None and reduction != 'mean':
        raise ValueError('avg_factor can not be used with reduction="sum"')          # losses/utils.py:52-55
    label = seg_label if seg_label.dim() == 3 else seg_label.squeeze(1)
    (h, w), (H, W) = seg_logit.shape[2:], label.shape[1:]
    scale = H // h if h and H % h == 0 else 0
    if align_corners or scale not in _cabi.CE_SCALES or W != w * scale:
        seg_logit = F.interpolate(seg_logit, size=(H, W), mode='bilinear', align_corners=align_corners)
        scale = 1
    n_pix = label.numel()
    denom = float(avg_factor) if avg_factor is not None else (float(n_pix) if reduction == 'mean' else 1.0)
    cw = None if class_weight is None else tuple(float(v) for v in class_weight)
    return _CEUp.apply(seg_logit, label, scale, cw, pixel_weight, int(ignore_index), float(loss_weight), denom)


class CrossEntropyLoss(nn.Module):
    """Drop-in for the reference's ``CrossEntropyLoss`` (softmax variant): same constructor, same ``forward`` arguments.
    ``cls_score`` may be at a lower resolution than ``label`` (the resize of ``BaseDecodeHead.losses`` is then fused
    in); ``last_acc`` holds the top-1 accuracy (percent) of the last call."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None, loss_weight=1.0):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        self.use_sigmoid, self.use_mask = use_sigmoid, use_mask
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.class_weight = class_weight
        self.last_acc = None

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=-100,
                align_corners=False, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        if self.use_sigmoid or self.use_mask:
            raise _cabi.SegDistillUnsupported('fused cross-entropy covers the softmax variant only')
        reduction = reduction_override if reduction_override else self.reduction
        loss, acc = seg_ce_loss(cls_score, label, class_weight=self.class_weight, pixel_weight=weight,
                                ignore_index=ignore_index, loss_weight=self.loss_weight, reduction=reduction,
                                avg_factor=avg_factor, align_corners=align_corners)
        self.last_acc = acc
        return loss


def decode_head_losses(seg_logit, seg_label, loss_decode=None, ignore_index=255, align_corners=False, sampler=None):
    """``BaseDecodeHead.losses`` (decode_head.py:217-237): ``{'loss_seg', 'acc_seg'}`` from the head's low-resolution
    logits and the full-resolution labels.  ``loss_decode``: a :class:`CrossEntropyLoss` (default: the reference's
    default ``CrossEntropyLoss()``); ``sampler``: an object with ``sample(seg_logit, seg_label)`` (the reference's
    pixel sampler works on the RESIZED logits, so with a sampler the logits are resized on the host first)."""
    crit = loss_decode if loss_decode is not None else CrossEntropyLoss()
    weight = None
    if sampler is not None:
        seg_logit = F.interpolate(seg_logit, size=tuple(seg_label.shape[2:]), mode='bilinear', align_corners=align_corners)
        weight = sampler.sample(seg_logit, seg_label)
    loss = crit(seg_logit, seg_label, weight=weight, ignore_index=ignore_index, align_corners=align_corners)
    return {'loss_seg': loss, 'acc_seg': crit.last_acc}
