"""Drop-in loss modules: same names, constructor arguments and 4-argument call as the
reference's ``mmseg/models/distillation/losses.py`` (``KLDLoss`` :9-113, ``PDLoss`` :115,
``CDLoss`` :130, ``CGDLoss`` :145, ``CGDLossWS`` :160, ``ATLoss`` :175), computing through the
sm_100a kernels.  ``criterion(x_student, x_teacher, gt_semantic_seg, step)`` -> 0-dim loss
whose backward delivers the gradient to ``x_student`` only (reference call site:
``mmseg/models/distillation/opts.py:103``).  The 2-argument form is accepted too
(``gt=None`` => no resize, ``step=0``).

Host-side behaviour kept from the reference: the alpha warm-up / early-decay state machine
(:61-92), the channel shuffle drawn from ``torch.randperm`` on the CPU global generator
(:39) every ``interval`` steps, the bilinear resize to the label size (:25-33).  What changed:
the shuffle and the ragged-group pad cost no copies (the kernel gathers / skips), the resize is
skipped when sizes already match (it is an exact identity there) and, for channel-mode losses behind a
2x / 4x / 8x bilinear resize (every shipped preset), done inside the kernel without materialising the
resized maps (``fuse_resize``), when ``alpha == 0`` no
kernel runs at all, and the returned scalar is always fp32 (the kernels accumulate in fp32;
the reference returns the feature dtype, i.e. a bf16/fp16-rounded loss under mixed precision).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from . import functional as SF

__all__ = ['KLDLoss', 'PDLoss', 'CDLoss', 'CGDLoss', 'CGDLossWS', 'ATLoss', 'IFVDLoss',
           'FeatureMSELoss', 'CDMSELoss', 'CGDCorrLoss']


def _ramp(kind, alpha0, frac):
    if kind == 'linear':
        return alpha0 * frac
    if kind == 'exp':
        return alpha0 ** frac
    if kind == 'jump':
        return 0
    return None


class KLDLoss(nn.Module):
    algo = 'auto'          # 'auto' | 'tma' | 'generic' (tests force one)
    fuse_resize = True     # behind a 2x / 4x / 8x bilinear resize: up-sample inside the kernel

    def __init__(self, alpha=1, tau=1, resize_config=None, shuffle_config=None, transform_config=None,
                 warmup_config=None, earlydecay_config=None):
        super().__init__()
        self.alpha_0 = alpha
        self.alpha = alpha
        self.tau = tau
        self.resize_config = resize_config
        self.shuffle_config = shuffle_config
        self.transform_config = transform_config
        self.warmup_config = warmup_config
        self.earlydecay_config = earlydecay_config
        self.last_perm = None

    # -- host-side schedules (reference :61-92); a branch that does not fire keeps the old alpha
    def _update_alpha(self, n_iter):
        wc, dc = self.warmup_config, self.earlydecay_config
        if wc:
            span = wc['warmup_iters']
            if n_iter == span:
                self.alpha = self.alpha_0
            elif n_iter < span:
                v = _ramp(wc['mode'], self.alpha_0, n_iter / span)
                if v is not None:
                    self.alpha = v
        if dc:
            begin, stop = dc['earlydecay_start'], dc['earlydecay_end']
            if n_iter >= stop:
                self.alpha = 0
            elif begin < n_iter < stop:
                frac = (stop - n_iter) / (stop - begin)
                v = _ramp(dc['mode'], self.alpha_0, frac)
                if v is not None:
                    self.alpha = 0.001 * v if dc['mode'] == 'exp' else v

    def _resized(self, x, gt):
        if gt is None or tuple(gt.shape[2:]) == tuple(x.shape[2:]):
            return x                   # bilinear resize to the same size is an exact identity
        return F.interpolate(x, size=tuple(gt.shape[2:]), mode=self.resize_config['mode'],
                             align_corners=self.resize_config['align_corners'])

    # -- the host-side half of a call: schedules, resize, shuffle draw (same order as the reference :96-106)
    def plan(self, x_student, x_teacher, gt=None, n_iter=0, resized=None):
        """Everything ``forward`` decides on the host, without launching: a dict the dispatcher can
        inspect to batch two entries that hook the same tensors into one kernel.  ``resized`` is an
        optional cache {(id(tensor), size): resized tensor} shared between the entries of one step."""
        self._update_alpha(n_iter)
        upscale = self._fusable_upscale(x_student, x_teacher, gt)
        if self.resize_config and not upscale:
            x_student = self._resized_cached(x_student, gt, resized)
            x_teacher = self._resized_cached(x_teacher, gt, resized)
        perm = None
        if self.shuffle_config and n_iter % self.shuffle_config['interval'] == 0:
            perm = torch.randperm(x_student.shape[1])      # same draw as the reference (:39)
        self.last_perm = perm
        tc = self.transform_config
        return {'student': x_student, 'teacher': x_teacher, 'perm': perm, 'alpha': self.alpha, 'tau': self.tau,
                'kind': tc['loss_type'] if tc else None, 'group': tc.get('group_size') if tc else None,
                'algo': self.algo, 'upscale': upscale}

    def _fusable_upscale(self, x_student, x_teacher, gt):
        """Integer factor of the resize when the kernel can do it on the fly (reference :25-33 with the presets'
        bilinear / align_corners=False, channel mode, maps on the GPU), else 0 (resize on the host as before)."""
        from . import _cabi
        rc, tc = self.resize_config, self.transform_config
        if not (self.fuse_resize and rc and tc and gt is not None and self.algo == 'auto'):
            return 0
        if tc['loss_type'] not in ('channel', 'pixel') or rc.get('mode') != 'bilinear' or rc.get('align_corners'):
            return 0
        x = x_student
        if x.dim() != 4 or not x.is_cuda or x.shape != x_teacher.shape or x.dtype != x_teacher.dtype:
            return 0
        if x.dtype not in (torch.float32, torch.bfloat16):
            return 0
        (h, w), (hg, wg) = x.shape[2:], gt.shape[2:]
        scales = _cabi.UP_SCALES if tc['loss_type'] == 'channel' else _cabi.UP_PIXEL_SCALES
        if hg % h or wg % w or hg // h != wg // w or hg // h not in scales:
            return 0
        return hg // h if _cabi.up_supported(h, w) else 0

    def _resized_cached(self, x, gt, cache):
        if cache is None or gt is None:
            return self._resized(x, gt)
        key = (id(x), tuple(gt.shape[2:]), self.resize_config['mode'], self.resize_config['align_corners'])
        if key not in cache:
            cache[key] = self._resized(x, gt)
        return cache[key]

    @staticmethod
    def run(plan):
        x_student, x_teacher = plan['student'], plan['teacher']
        if plan['alpha'] == 0:
            return SF.zero_loss(x_student)
        kind = plan['kind']
        if plan.get('upscale') and kind == 'pixel':
            return SF.kl_pixels_up_loss(x_student, x_teacher, plan['upscale'], tau=plan['tau'], alpha=plan['alpha'])
        if plan.get('upscale'):
            return SF.kl_rows_up_loss(x_student, x_teacher, plan['upscale'], group=plan['group'], tau=plan['tau'],
                                      alpha=plan['alpha'], perm=plan['perm'])
        if kind == 'pixel':
            # softmax over all channels of a pixel: a channel permutation changes nothing
            return SF.kl_pixels_loss(x_student, x_teacher, tau=plan['tau'], alpha=plan['alpha'], algo=plan['algo'])
        if kind == 'channel':
            return SF.kl_rows_loss(x_student, x_teacher, group=plan['group'], tau=plan['tau'],
                                   alpha=plan['alpha'], perm=plan['perm'], algo=plan['algo'])
        # no transform: softmax over the last dim of the 4-D maps; rows = (b, c, h)
        lead = x_student.numel() // x_student.shape[-1]
        return SF.kl_rows_loss(x_student, x_teacher, group=1, tau=plan['tau'], alpha=plan['alpha'],
                               algo=plan['algo'], bchw=(1, lead, x_student.shape[-1]))

    @staticmethod
    def can_fuse(pa, pb):
        """Two planned calls that one two-loss launch can serve: channel mode on the very same (resized)
        tensors, no shuffle this step, non-zero weights, nested rows, default kernel selection."""
        if pa['kind'] != 'channel' or pb['kind'] != 'channel' or pa.get('upscale') or pb.get('upscale'):
            return False
        if pa['student'] is not pb['student'] or pa['teacher'] is not pb['teacher']:
            return False
        if pa['perm'] is not None or pb['perm'] is not None or pa['alpha'] == 0 or pb['alpha'] == 0:
            return False
        if pa['algo'] != 'auto' or pb['algo'] != 'auto':
            return False
        x = pa['student']
        if x.dim() != 4 or x.dtype not in (torch.float32, torch.bfloat16) or not x.is_cuda:
            return False
        from . import _cabi
        return _cabi.multi_supported(x.shape, (pa['group'], pb['group']), x.element_size())

    @staticmethod
    def run_pair(pa, pb):
        """(loss_a, loss_b) of two fusable plans from ONE pass over the maps; falls back to two launches
        when the library declines the layout (rows too long to keep every chunk co-resident)."""
        from . import _cabi
        try:
            return SF.kl_rows_pair_loss(pa['student'], pa['teacher'], pa['group'], pa['tau'], pa['alpha'],
                                        pb['group'], pb['tau'], pb['alpha'])
        except _cabi.SegDistillUnsupported:
            return KLDLoss.run(pa), KLDLoss.run(pb)

    @staticmethod
    def can_group(plan):
        """A planned call that a grouped launch (several pairs, one kernel) can take: channel mode, no shuffle on this
        step, a non-zero weight, no resize left to fuse, default kernel selection, 16-byte aligned rows."""
        x = plan['student']
        if plan['kind'] != 'channel' or plan.get('upscale') or plan['perm'] is not None or plan['alpha'] == 0:
            return False
        if plan['algo'] != 'auto' or x.dim() != 4 or not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16):
            return False
        if x.shape != plan['teacher'].shape:
            return False
        hw = x.shape[2] * x.shape[3]
        return (hw * x.element_size()) % 16 == 0

    @staticmethod
    def run_group(plans):
        """Losses of several groupable plans from ONE launch; falls back to one launch per plan when the library
        declines (SD_ERR_UNSUPPORTED)."""
        from . import _cabi
        try:
            return SF.kl_rows_group_loss([(p['student'], p['teacher']) for p in plans], [p['group'] for p in plans],
                                         [p['tau'] for p in plans], [p['alpha'] for p in plans])
        except _cabi.SegDistillUnsupported:
            return tuple(KLDLoss.run(p) for p in plans)

    def forward(self, x_student, x_teacher, gt=None, n_iter=0):
        return self.run(self.plan(x_student, x_teacher, gt, n_iter))


def _bilinear():
    return {'mode': 'bilinear', 'align_corners': False}


class PDLoss(KLDLoss):
    """Per-pixel logit KL (reference :115-128)."""

    def __init__(self):
        super().__init__(alpha=1, tau=1, resize_config=_bilinear(), transform_config={'loss_type': 'pixel'})


class CDLoss(KLDLoss):
    """Channel-wise spatial-softmax KL, a.k.a. CWD (reference :130-143)."""

    def __init__(self):
        super().__init__(alpha=1, tau=1, resize_config=_bilinear(),
                         transform_config={'loss_type': 'channel', 'group_size': 1})


class CGDLoss(KLDLoss):
    """Channel Group Distillation: the same KL over groups of ``group_size`` channels (reference :145-158)."""

    def __init__(self, group_size=10, alpha=3, tau=2):
        super().__init__(alpha=alpha, tau=tau, resize_config=_bilinear(), shuffle_config={'interval': 1000},
                         transform_config={'loss_type': 'channel', 'group_size': group_size})


class CGDLossWS(KLDLoss):
    """CGD with warm-up and early decay of the weight (reference :160-173)."""

    def __init__(self):
        super().__init__(alpha=3, tau=2, resize_config=_bilinear(), shuffle_config={'interval': 1000},
                         transform_config={'loss_type': 'channel', 'group_size': 10},
                         warmup_config={'mode': 'linear', 'warmup_iters': 2000},
                         earlydecay_config={'mode': 'linear', 'earlydecay_start': 110000,
                                            'earlydecay_end': 120000})


class ATLoss(nn.Module):
    """MSE of the channel-mean maps + per-pixel KL, one fused kernel (reference :175-197)."""
    algo = 'auto'

    def __init__(self):
        super().__init__()

    def forward(self, x_student, x_teacher, gt=None, step=0):
        loss = SF.kl_pixels_loss(x_student, x_teacher, tau=1.0, alpha=1.0, at_weight=1.0, algo=self.algo)
        return loss


class IFVDLoss(nn.Module):
    """Intra-class feature variation distillation (reference :199-238): per-pixel KL + 10 * MSE between the
    student's and the teacher's maps of cosine similarity of every pixel to the centre of its class.

    The KL term runs in the fused pixel kernel.  The similarity term - in the reference a Python loop over all C
    classes (~10 full-size ATen ops per class and tensor), two cosine similarities, an MSE and the autograd
    backward of all of it - is one C-ABI call (csrc/ifvd.cu: deterministic segmented reductions for the class
    centres, one sweep per pixel, the gradient including the path through the centres).
    Pixels whose label matches no class (ignore index 255) keep their own feature as centre, as in the reference.
    """
    algo = 'auto'

    def __init__(self):
        super().__init__()

    @staticmethod
    def _class_map(target, c, h, w):
        """(B, h*w) int32 class of every pixel: the label nearest-resized to the feature size (:218-219) where it
        equals one of the class indices 0..C-1 the reference loops over (:222-224), else C ("no class")."""
        b = target.shape[0]
        lab = F.interpolate(target.float(), size=(h, w), mode='nearest').reshape(b, h * w)
        k = lab.long()
        valid = (k.to(lab.dtype) == lab) & (k >= 0) & (k < c)
        return torch.where(valid, k, torch.full_like(k, c)).to(torch.int32)

    def forward(self, preds_S, preds_T, target, step=0):
        feat_s, feat_t = preds_S, preds_T
        if feat_t.shape[2:] != feat_s.shape[2:]:
            feat_t = F.interpolate(feat_t, size=feat_s.shape[2:], mode='bilinear', align_corners=False)   # :204-209
        feat_t = feat_t.detach()
        b, c, h, w = feat_s.shape
        if c > _cabi.ifvd_max_channels():      # before anything is launched (the KL kernel's work would be thrown away)
            raise _cabi.SegDistillUnsupported(
                f'IFVDLoss: C = {c} classes exceed the {_cabi.ifvd_max_channels()} the class-sum kernels hold per CTA')
        if target.is_cuda and not target.is_floating_point():
            cls = _cabi.ifvd_class_map(target, c, h, w)                                                    # :218-224
        else:
            cls = self._class_map(target, c, h, w).to(feat_s.device)
        return SF.ifvd_loss(feat_s, feat_t, cls, 10.0, algo=self.algo)                                    # :213-237


class FeatureMSELoss(nn.Module):
    """``weight * mean((s - t)^2)`` - the feature-MSE term (reference :178,190 / commented class :812-830)."""

    def __init__(self, weight=1.0):
        super().__init__()
        self.weight = weight

    def forward(self, x_student, x_teacher, gt=None, step=0):
        return SF.mse_loss(x_student, x_teacher, self.weight)


class CDMSELoss(nn.Module):
    """CWD (tau, alpha) + feature MSE on the same pair in ONE pass over the maps (BASELINE config 4)."""
    algo = 'auto'

    def __init__(self, alpha=1, tau=1, mse_weight=1.0, group_size=1):
        super().__init__()
        self.alpha, self.tau, self.mse_weight, self.group_size = alpha, tau, mse_weight, group_size
        self.last_parts = None

    def forward(self, x_student, x_teacher, gt=None, step=0):
        total, kl, mse = SF.kl_rows_mse_loss(x_student, x_teacher, group=self.group_size, tau=self.tau,
                                             alpha=self.alpha, mse_weight=self.mse_weight, algo=self.algo)
        self.last_parts = (kl, mse)
        return total


class CGDCorrLoss(nn.Module):
    """Per-group Gram-matrix (correlation) loss - an extension, NOT in the reference."""

    def __init__(self, group_size=10, alpha=1.0):
        super().__init__()
        self.group_size, self.alpha = group_size, alpha

    def forward(self, x_student, x_teacher, gt=None, step=0):
        return SF.cgd_corr_loss(x_student, x_teacher, self.group_size, self.alpha)
