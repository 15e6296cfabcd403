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

from . import _cabi


def once_differentiable(fn):
    """The backward of these nodes launches kernels through the C ABI: it cannot be differentiated again.  torch's
    ``once_differentiable`` enforces that by wrapping every call (~30 us of host time per backward); here the same
    contract is one check: a backward that is itself being recorded (``create_graph=True``) raises."""
    def backward(ctx, *grads):
        if torch.is_grad_enabled():
            raise RuntimeError('segdistill_b200 losses are differentiable once (their backward is a fused kernel)')
        return fn(ctx, *grads)
    return backward


def _keep(ctx, x_student, x_teacher, ds):
    """Forward's bookkeeping: the gradient buffer the kernel already filled, and (only references) the inputs, so
    that a SECOND backward through the same node can rebuild it."""
    ctx.needs = x_student.requires_grad
    ctx.ds = ds if ctx.needs else None
    ctx.in_dtype, ctx.in_shape = x_student.dtype, x_student.shape
    if ctx.needs:
        ctx.save_for_backward(x_student, x_teacher)


def _finish_backward(ctx, grad_output, recompute):
    """dS * grad_output.  The first backward hands out the buffer forward filled (scaled in place; our reference is
    dropped so autograd can adopt it without a copy).  A later backward through the same node - ``retain_graph=True``,
    e.g. the reference trainer's ``log_grad`` mode, SD_structure.py:92-134 - finds it gone and re-runs the kernel on
    the saved inputs (``recompute(x_student, x_teacher) -> dS``).  Without ``retain_graph`` the saved tensors are freed
    and autograd raises its usual error."""
    if not ctx.needs:
        return None
    ds = ctx.ds
    ctx.ds = None
    if ds is None:
        ds = recompute(*ctx.saved_tensors)
    _cabi.scale_grad_(ds, grad_output)
    if ds.dtype != ctx.in_dtype:
        ds = ds.to(ctx.in_dtype)
    return ds.view(ctx.in_shape)


class _KLRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_student, x_teacher, group, tau, alpha, perm, mse_weight, algo, bchw):
        ctx.call = dict(group=group, tau=tau, alpha=alpha, perm=perm, mse_weight=mse_weight, algo=algo, bchw=bchw)
        loss, ds, _, mse = _cabi.kl_rows(x_student, x_teacher, **ctx.call)
        _keep(ctx, x_student, x_teacher, ds)
        if mse is not None:
            total = loss + mse
            ctx.mark_non_differentiable(loss, mse)
            return total, loss, mse
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, *unused):
        return (_finish_backward(ctx, grad_output, lambda s, t: _cabi.kl_rows(s, t, **ctx.call)[1]),) + (None,) * 8


class _KLRowsUp(torch.autograd.Function):
    """Channel-mode KL behind the reference's bilinear resize (losses.py:25-33,101-102), the resize fused in."""

    @staticmethod
    def forward(ctx, x_student, x_teacher, scale, group, tau, alpha, perm):
        ctx.call = dict(group=group, tau=tau, alpha=alpha, perm=perm)
        ctx.scale = scale
        loss, ds, _ = _cabi.kl_rows_up(x_student, x_teacher, scale, **ctx.call)
        _keep(ctx, x_student, x_teacher, ds)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return (_finish_backward(ctx, grad_output, lambda s, t: _cabi.kl_rows_up(s, t, ctx.scale, **ctx.call)[1]),) + (None,) * 6


class _KLPixelsUp(torch.autograd.Function):
    """Pixel-mode KL behind the reference's bilinear resize, the resize fused in."""

    @staticmethod
    def forward(ctx, x_student, x_teacher, scale, tau, alpha):
        ctx.call = (scale, tau, alpha)
        loss, ds = _cabi.kl_pixels_up(x_student, x_teacher, scale, tau=tau, alpha=alpha)
        _keep(ctx, x_student, x_teacher, ds)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        scale, tau, alpha = ctx.call
        return (_finish_backward(ctx, grad_output,
                                 lambda s, t: _cabi.kl_pixels_up(s, t, scale, tau=tau, alpha=alpha)[1]),) + (None,) * 4


PAIR_ALGO = 'auto'      # kernel of the fused two-loss launch: 'auto' | 'cluster' | 'stream' (tests force one)


class _KLRowsMulti(torch.autograd.Function):
    """Two channel-mode KL losses on one pair, one kernel: returns both scalars; dS is their summed gradient.

    dS is built in forward for upstream gradients of 1.  Backward: if both upstream gradients are equal
    (the usual plain sum, possibly times a loss scale) dS is scaled in place on the device; if they differ,
    a flag set by that same kernel lets a second launch of the fused kernel rebuild dS with the individual
    factors - no host synchronisation either way.
    """

    @staticmethod
    def forward(ctx, x_student, x_teacher, g0, tau0, alpha0, g1, tau1, alpha1):
        algo = _cabi.ALGOS[PAIR_ALGO]
        ctx.cfg = cfg = ((g0, g1), (tau0, tau1), (alpha0, alpha1))
        losses, ds = _cabi.kl_rows_multi(x_student, x_teacher, *cfg, algo=algo)
        ctx.algo = algo
        _keep(ctx, x_student, x_teacher, ds)
        return losses.unbind(0)

    @staticmethod
    @once_differentiable
    def backward(ctx, go0, go1):
        if not ctx.needs:
            return (None,) * 8
        ds = ctx.ds
        ctx.ds = None
        groups, taus, alphas = ctx.cfg
        if ds is None:       # a second backward through this node (retain_graph=True): rebuild dS for unit weights
            x_student, x_teacher = ctx.saved_tensors
            ds = _cabi.kl_rows_multi(x_student, x_teacher, groups, taus, alphas, algo=ctx.algo)[1]
        dev = ds.device
        if go0 is go1 or (go0.data_ptr() == go1.data_ptr() and go0.device == go1.device):
            # the two terms entered one sum: a single upstream gradient, one in-place scaling launch
            _cabi.scale_grad_(ds, go0)
        else:
            x_student, x_teacher = ctx.saved_tensors
            go0 = go0.detach().to(device=dev, dtype=torch.float32).reshape(1)
            go1 = go1.detach().to(device=dev, dtype=torch.float32).reshape(1)
            flag = _cabi.scale_grad2_(ds, go0, go1)
            _cabi.kl_rows_multi(x_student, x_teacher, groups, taus, alphas, grad_outputs=(go0, go1), run_if=flag,
                                ds=ds, algo=ctx.algo)
        if ds.dtype != ctx.in_dtype:
            ds = ds.to(ctx.in_dtype)
        return (ds.view(ctx.in_shape),) + (None,) * 7


class _KLRowsGroup(torch.autograd.Function):
    """The channel-mode KL losses of several (student, teacher) pairs from ONE launch (a dispatcher
    step over several layers).  ``apply(cfg, s0, t0, s1, t1, ...)`` with cfg = (groups, taus, alphas) -> one loss per
    pair; backward hands every student the gradient the launch wrote for it, times its upstream gradient."""

    @staticmethod
    def forward(ctx, cfg, *maps):
        students, teachers = maps[0::2], maps[1::2]
        losses, dss = _cabi.kl_rows_group(students, teachers, *cfg)
        ctx.cfg = cfg
        ctx.needs = [s.requires_grad for s in students]
        ctx.dss = dss if any(ctx.needs) else None
        ctx.meta = [(s.dtype, s.shape) for s in students]
        if any(ctx.needs):
            ctx.save_for_backward(*maps)
        return tuple(losses[k] for k in range(len(students)))

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        n = len(ctx.meta)
        if not any(ctx.needs):
            return (None,) * (1 + 2 * n)
        dss = ctx.dss
        ctx.dss = None
        if dss is None:                     # a second backward through this node (retain_graph=True): rebuild
            maps = ctx.saved_tensors
            dss = _cabi.kl_rows_group(maps[0::2], maps[1::2], *ctx.cfg)[1]
        live = [k for k in range(n) if ctx.needs[k]]
        _cabi.scale_grad_group_([dss[k] for k in live], [grads[k] for k in live])      # one launch for every layer
        out = [None]
        for k in range(n):
            g = None
            if ctx.needs[k]:
                g = dss[k]
                dtype, shape = ctx.meta[k]
                if g.dtype != dtype:
                    g = g.to(dtype)
                g = g.view(shape)
            out += [g, None]
        return tuple(out)


class _KLPixels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_student, x_teacher, tau, alpha, at_weight, algo):
        ctx.call = dict(tau=tau, alpha=alpha, at_weight=at_weight, algo=algo)
        loss, ds, _, at = _cabi.kl_pixels(x_student, x_teacher, **ctx.call)
        _keep(ctx, x_student, x_teacher, ds)
        if at is not None:
            return loss + at
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return (_finish_backward(ctx, grad_output, lambda s, t: _cabi.kl_pixels(s, t, **ctx.call)[1]),) + (None,) * 5


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_student, x_teacher, weight):
        ctx.weight = weight
        loss, ds = _cabi.mse(x_student, x_teacher, weight=weight)
        _keep(ctx, x_student, x_teacher, ds)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _finish_backward(ctx, grad_output, lambda s, t: _cabi.mse(s, t, weight=ctx.weight)[1]), None, None


class _IFVDSim(torch.autograd.Function):
    """IFVDLoss's similarity term (losses.py:218-235): class centres, cosine similarity, MSE and their backward."""

    @staticmethod
    def forward(ctx, x_student, x_teacher, cls, weight):
        ctx.call = (cls, weight)
        loss, ds = _cabi.ifvd_sim(x_student, x_teacher, cls, weight=weight)
        _keep(ctx, x_student, x_teacher, ds)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        cls, weight = ctx.call
        return _finish_backward(ctx, grad_output, lambda s, t: _cabi.ifvd_sim(s, t, cls, weight=weight)[1]), None, None, None


class _IFVD(torch.autograd.Function):
    """Whole IFVDLoss (losses.py:213-237): the per-pixel KL kernel writes dS, the similarity term adds its gradient
    to the same buffer - one gradient tensor, one backward node."""

    @staticmethod
    def forward(ctx, x_student, x_teacher, cls, weight, algo):
        ctx.call = (cls, weight, algo)
        loss_pd, ds, _, _ = _cabi.kl_pixels(x_student, x_teacher, tau=1.0, alpha=1.0, algo=algo)
        loss_sim, ds = _cabi.ifvd_sim(x_student, x_teacher, cls, weight=weight, ds=ds)
        _keep(ctx, x_student, x_teacher, ds)
        return loss_sim + loss_pd

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        cls, weight, algo = ctx.call

        def again(s, t):
            ds = _cabi.kl_pixels(s, t, tau=1.0, alpha=1.0, algo=algo)[1]
            return _cabi.ifvd_sim(s, t, cls, weight=weight, ds=ds)[1]
        return _finish_backward(ctx, grad_output, again), None, None, None, None


class _CGDCorr(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_student, x_teacher, group, alpha):
        ctx.call = (group, alpha)
        loss, ds = _cabi.cgd_corr(x_student, x_teacher, group=group, alpha=alpha)
        _keep(ctx, x_student, x_teacher, ds)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        group, alpha = ctx.call
        return _finish_backward(ctx, grad_output, lambda s, t: _cabi.cgd_corr(s, t, group=group, alpha=alpha)[1]), None, None, None


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


def kl_rows_up_loss(x_student, x_teacher, scale, group=1, tau=1.0, alpha=1.0, perm=None):
    """The same on maps up-sampled ``scale`` x (bilinear, align_corners=False) inside the kernel; the gradient
    arrives at the low-resolution ``x_student``."""
    return _KLRowsUp.apply(x_student, x_teacher, int(scale), int(group), float(tau), float(alpha), perm)


def kl_pixels_up_loss(x_student, x_teacher, scale, tau=1.0, alpha=1.0):
    """Per-pixel KL over channels on maps up-sampled ``scale`` x inside the kernel; gradient at low resolution."""
    return _KLPixelsUp.apply(x_student, x_teacher, int(scale), float(tau), float(alpha))


def kl_rows_mse_loss(x_student, x_teacher, group=1, tau=1.0, alpha=1.0, mse_weight=1.0, perm=None, algo='auto'):
    """Fused CWD + feature MSE on the same pair: returns (total, kl_part, mse_part); only total carries grad."""
    return _KLRows.apply(x_student, x_teacher, int(group), float(tau), float(alpha), perm, float(mse_weight),
                         _cabi.ALGOS[algo], None)


def kl_rows_pair_loss(x_student, x_teacher, group0, tau0, alpha0, group1, tau1, alpha1):
    """Two channel-mode KL losses over the same pair in one pass; returns (loss0, loss1)."""
    return _KLRowsMulti.apply(x_student, x_teacher, int(group0), float(tau0), float(alpha0),
                              int(group1), float(tau1), float(alpha1))


def kl_rows_group_loss(pairs, groups, taus, alphas):
    """Channel-mode KL of several (student, teacher) pairs in one launch; returns a tuple of losses, one per pair."""
    maps = [m for pair in pairs for m in pair]
    cfg = (tuple(int(g) for g in groups), tuple(float(v) for v in taus), tuple(float(v) for v in alphas))
    return _KLRowsGroup.apply(cfg, *maps)


def kl_pixels_loss(x_student, x_teacher, tau=1.0, alpha=1.0, at_weight=0.0, algo='auto'):
    """alpha/(B*HW) * sum_pixels KL over channels (+ at_weight * MSE of the channel-mean maps)."""
    return _KLPixels.apply(x_student, x_teacher, float(tau), float(alpha), float(at_weight), _cabi.ALGOS[algo])


def mse_loss(x_student, x_teacher, weight=1.0):
    return _MSE.apply(x_student, x_teacher, float(weight))


def ifvd_sim_loss(x_student, x_teacher, cls, weight=10.0):
    """weight * mean over pixels of (cos(s, centre_s) - cos(t, centre_t))^2, centres per sample and class."""
    return _IFVDSim.apply(x_student, x_teacher, cls, float(weight))


def ifvd_loss(x_student, x_teacher, cls, weight=10.0, algo='auto'):
    """per-pixel KL (tau = alpha = 1) + weight * the class-centre similarity MSE, as one autograd node."""
    return _IFVD.apply(x_student, x_teacher, cls, float(weight), _cabi.ALGOS[algo])


def cgd_corr_loss(x_student, x_teacher, group=10, alpha=1.0):
    return _CGDCorr.apply(x_student, x_teacher, int(group), float(alpha))


def zero_loss(x_student):
    return _ZeroLoss.apply(x_student)
