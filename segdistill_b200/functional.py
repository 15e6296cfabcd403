"""Autograd bridges: the fused kernels compute loss AND d(loss)/d(student) in one pass.

Forward launches the kernel and keeps dS; backward hands dS to autograd after a
device-side multiply by grad_output that exits immediately when grad_output == 1
(the distillation loss enters the total as a plain sum, reference
``mmseg/models/segmentors/SD_structure.py:121-122``; 512 under fp16 loss scaling).
The teacher never receives a gradient (reference: frozen, run under ``no_grad``,
``SD_structure.py:44-45,65-67``).
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _cabi


def _finish_backward(ctx, grad_output):
    ds = ctx.ds
    ctx.ds = None                      # drop our reference so autograd can adopt the buffer without a copy
    if ds is None:
        return None
    _cabi.scale_grad_(ds, grad_output)
    if ds.dtype != ctx.in_dtype:
        ds = ds.to(ctx.in_dtype)
    return ds.view(ctx.in_shape)


class _KLRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_student, x_teacher, group, tau, alpha, perm, mse_weight, algo, bchw):
        need_grad = x_student.requires_grad
        loss, ds, _, mse = _cabi.kl_rows(x_student, x_teacher, group=group, tau=tau, alpha=alpha, perm=perm,
                                         mse_weight=mse_weight, algo=algo, bchw=bchw)
        ctx.ds = ds if need_grad else None
        ctx.in_dtype, ctx.in_shape = x_student.dtype, x_student.shape
        if mse is not None:
            total = loss + mse
            ctx.mark_non_differentiable(loss, mse)
            return total, loss, mse
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, *unused):
        return (_finish_backward(ctx, grad_output),) + (None,) * 8


class _KLPixels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_student, x_teacher, tau, alpha, at_weight, algo):
        need_grad = x_student.requires_grad
        loss, ds, _, at = _cabi.kl_pixels(x_student, x_teacher, tau=tau, alpha=alpha, at_weight=at_weight, algo=algo)
        ctx.ds = ds if need_grad else None
        ctx.in_dtype, ctx.in_shape = x_student.dtype, x_student.shape
        if at is not None:
            return loss + at
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return (_finish_backward(ctx, grad_output),) + (None,) * 5


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_student, x_teacher, weight):
        loss, ds = _cabi.mse(x_student, x_teacher, weight=weight)
        ctx.ds = ds if x_student.requires_grad else None
        ctx.in_dtype, ctx.in_shape = x_student.dtype, x_student.shape
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _finish_backward(ctx, grad_output), None, None


class _CGDCorr(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_student, x_teacher, group, alpha):
        loss, ds = _cabi.cgd_corr(x_student, x_teacher, group=group, alpha=alpha)
        ctx.ds = ds if x_student.requires_grad else None
        ctx.in_dtype, ctx.in_shape = x_student.dtype, x_student.shape
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _finish_backward(ctx, grad_output), None, None, None


class _ZeroLoss(torch.autograd.Function):
    """alpha == 0 (before warm-up / after early decay): zero loss, zero (None) gradient, no kernel."""

    @staticmethod
    def forward(ctx, x_student):
        return torch.zeros((), dtype=torch.float32, device=x_student.device)

    @staticmethod
    def backward(ctx, grad_output):
        return None


def kl_rows_loss(x_student, x_teacher, group=1, tau=1.0, alpha=1.0, perm=None, algo='auto', bchw=None):
    """alpha/R * sum_rows KL(softmax(T_row/tau) || softmax(S_row/tau)); rows = ``group`` channels x HW."""
    return _KLRows.apply(x_student, x_teacher, int(group), float(tau), float(alpha), perm, 0.0,
                         _cabi.ALGOS[algo], bchw)


def kl_rows_mse_loss(x_student, x_teacher, group=1, tau=1.0, alpha=1.0, mse_weight=1.0, perm=None, algo='auto'):
    """Fused CWD + feature MSE on the same pair: returns (total, kl_part, mse_part); only total carries grad."""
    return _KLRows.apply(x_student, x_teacher, int(group), float(tau), float(alpha), perm, float(mse_weight),
                         _cabi.ALGOS[algo], None)


def kl_pixels_loss(x_student, x_teacher, tau=1.0, alpha=1.0, at_weight=0.0, algo='auto'):
    """alpha/(B*HW) * sum_pixels KL over channels (+ at_weight * MSE of the channel-mean maps)."""
    return _KLPixels.apply(x_student, x_teacher, float(tau), float(alpha), float(at_weight), _cabi.ALGOS[algo])


def mse_loss(x_student, x_teacher, weight=1.0):
    return _MSE.apply(x_student, x_teacher, float(weight))


def cgd_corr_loss(x_student, x_teacher, group=10, alpha=1.0):
    return _CGDCorr.apply(x_student, x_teacher, int(group), float(alpha))


def zero_loss(x_student):
    return _ZeroLoss.apply(x_student)
