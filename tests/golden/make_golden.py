#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference's loss module (``/root/reference/mmseg/models/distillation/losses.py``)
is loaded by path.  ``import mmseg`` itself needs mmcv (not installed), so stub
``mmseg`` / ``mmseg.ops`` modules are registered first, with ``resize`` exec'd from the
reference's own ``mmseg/ops/wrappers.py`` (pure torch).  One environment patch: the
reference's ragged-group branch builds its filler with a hard-coded ``.cuda()``
(losses.py:56); this container has no GPU, so ``torch.Tensor.cuda`` is made a no-op
while the fixtures are generated.  No reference source is modified or copied.

The GPU box has no /root/reference: tests only read the committed .npz files.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get('SEGDISTILL_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_losses(ref_root=REF):
    """Return the reference ``losses`` module (unmodified source, stubbed mmseg.ops)."""
    name = '_segdistill_reference_losses'
    if name in sys.modules:
        return sys.modules[name]
    wrappers = types.ModuleType('mmseg.ops.wrappers')
    with open(os.path.join(ref_root, 'mmseg/ops/wrappers.py')) as f:
        exec(compile(f.read(), 'mmseg/ops/wrappers.py', 'exec'), wrappers.__dict__)
    pkg = types.ModuleType('mmseg')
    ops = types.ModuleType('mmseg.ops')
    ops.resize = wrappers.resize
    pkg.ops = ops
    saved = {k: sys.modules.get(k) for k in ('mmseg', 'mmseg.ops')}
    sys.modules['mmseg'], sys.modules['mmseg.ops'] = pkg, ops
    try:
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(ref_root, 'mmseg/models/distillation/losses.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    sys.modules[name] = mod
    return mod


class cuda_is_noop:
    """This container has no GPU: make ``Tensor.cuda()`` return the tensor itself."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig


# (case name, class name, ctor kwargs, attribute overrides, S shape, gt HxW, n_iter, dtype)
CASES = [
    ('cd_2x6x8x8',        'CDLoss',  {}, {}, (2, 6, 8, 8),   (8, 8),   1, 'float32'),
    ('cd_1x3x5x7_odd',    'CDLoss',  {}, {}, (1, 3, 5, 7),   (5, 7),   1, 'float32'),
    ('pd_2x6x8x8',        'PDLoss',  {}, {}, (2, 6, 8, 8),   (8, 8),   1, 'float32'),
    ('pd_2x19x6x10',      'PDLoss',  {}, {}, (2, 19, 6, 10), (6, 10),  1, 'float32'),
    ('cgd_default',       'CGDLoss', {}, {}, (2, 20, 8, 8),  (8, 8),   1, 'float32'),
    ('cgd_g3_pad',        'CGDLoss', dict(group_size=3, alpha=2, tau=4), {}, (2, 7, 6, 10), (6, 10), 7, 'float32'),
    ('cgd_g10_pad_c32',   'CGDLoss', dict(group_size=10), {}, (1, 32, 4, 4), (4, 4), 3, 'float32'),
    ('cgd_g150_one_row',  'CGDLoss', dict(group_size=150, alpha=1, tau=3), {}, (1, 150, 4, 4), (4, 4), 5, 'float32'),
    ('cgd_shuffle_n1000', 'CGDLoss', dict(group_size=5, alpha=3, tau=2), {}, (2, 20, 8, 8), (8, 8), 1000, 'float32'),
    ('cgd_shuffle_pad',   'CGDLoss', dict(group_size=4, alpha=1, tau=1), {}, (2, 10, 4, 6), (4, 6), 2000, 'float32'),
    ('cd_resize_8to32',   'CDLoss',  {}, {}, (2, 6, 8, 8),   (32, 32), 1, 'float32'),
    ('pd_resize_8to24',   'PDLoss',  {}, {}, (1, 5, 8, 8),   (24, 24), 1, 'float32'),
    ('cgd_resize_down',   'CGDLoss', dict(group_size=2), {}, (1, 4, 12, 12), (6, 6), 1, 'float32'),
    ('kld_plain_lastdim', 'KLDLoss', dict(alpha=2, tau=3), {}, (2, 3, 4, 9), (4, 9), 1, 'float32'),
    ('kld_custom_pixel',  'KLDLoss', dict(alpha=0.5, tau=2, transform_config={'loss_type': 'pixel'}), {},
     (2, 6, 4, 4), (4, 4), 1, 'float32'),
    ('cgdws_warm_n500',   'CGDLossWS', {}, {}, (1, 20, 4, 4), (4, 4), 500, 'float32'),
    ('cd_logits_x4',      'CDLoss',  {}, {}, (2, 6, 8, 8),   (8, 8),   1, 'float32*4'),
    ('cd_near_converged', 'CDLoss',  {}, {}, (2, 6, 16, 16), (16, 16), 1, 'near'),
    # nearly converged student (KL ~ 5e-5: lse_t - lse_s cancels), the other layouts; every 'near' fixture also holds
    # the reference run in float64 on the same inputs (loss_f64 / grad_f64)
    ('cgd_near_converged', 'CGDLoss', {}, {}, (2, 20, 16, 16), (16, 16), 1, 'near'),
    ('pd_near_converged',  'PDLoss',  {}, {}, (2, 6, 16, 16),  (16, 16), 1, 'near'),
    ('cd_near_offset',     'CDLoss',  {}, {}, (2, 6, 16, 16),  (16, 16), 1, 'near+3'),
    ('cd_near_64x64',      'CDLoss',  {}, {}, (1, 4, 64, 64),  (64, 64), 1, 'near'),
    ('cgd_near_resize_4x', 'CGDLoss', dict(group_size=3), {}, (1, 6, 8, 8), (32, 32), 1, 'near'),
    ('pd_near_resize_2x',  'PDLoss',  {}, {}, (1, 5, 8, 8),    (16, 16), 1, 'near'),
]


def _inputs(shape, kind, seed):
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(shape, generator=g)
    t = torch.randn(shape, generator=g)
    if kind == 'float32*4':
        s, t = s * 4, t * 4
    elif kind == 'near':
        t = s + 1e-2 * t
    elif kind == 'near+3':                 # ... up to a constant offset, which a softmax does not see
        t = s + 1e-2 * t + 3.0
    return s, t


def run_case(ref, case, seed):
    name, cls, kwargs, attrs, shape, gt_hw, n_iter, kind = case
    s, t = _inputs(shape, kind, seed)
    s.requires_grad_(True)
    gt = torch.zeros(shape[0], 1, *gt_hw, dtype=torch.int64)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        crit = getattr(ref, cls)(**kwargs)
    for k, v in attrs.items():
        setattr(crit, k, v)
    # the reference draws its channel permutation from the CPU global generator
    torch.manual_seed(1234 + seed)
    rng_state = torch.get_rng_state()
    with cuda_is_noop():
        loss = crit(s, t, gt, n_iter)
    loss.backward()
    perm = np.zeros(0, dtype=np.int64)
    sc = crit.shuffle_config
    if sc and n_iter % sc['interval'] == 0:
        torch.set_rng_state(rng_state)
        perm = torch.randperm(shape[1]).numpy()
    rec = dict(S=s.detach().numpy(), T=t.numpy(), gt_hw=np.array(gt_hw), n_iter=np.array(n_iter),
               loss=np.array(loss.item(), dtype=np.float64), grad=s.grad.numpy(), perm=perm,
               alpha_after=np.array(float(crit.alpha)), tau=np.array(float(crit.tau)),
               cls=np.array(cls), kwargs=np.array(repr(kwargs)), manual_seed=np.array(1234 + seed))
    if kind.startswith('near'):
        # the same reference code in float64: what the fp32 reference (and the CUDA path) is measured against
        s64 = s.detach().double().requires_grad_(True)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            crit64 = getattr(ref, cls)(**kwargs)
        torch.set_rng_state(rng_state)
        with cuda_is_noop():
            loss64 = crit64(s64, t.double(), gt, n_iter)
        loss64.backward()
        rec.update(loss_f64=np.array(loss64.item(), dtype=np.float64), grad_f64=s64.grad.numpy())
    return rec


def schedule_table(ref):
    """alpha after each call for the stateful schedules (losses.py:61-92)."""
    steps = [0, 1, 500, 1999, 2000, 2001, 50000, 110000, 110001, 115000, 119999, 120000, 130000]
    out = {}
    x = torch.randn(1, 20, 2, 2)
    gt = torch.zeros(1, 1, 2, 2, dtype=torch.int64)
    crit = ref.CGDLossWS()
    vals = []
    for n in steps:
        crit(x, x.clone(), gt, n)
        vals.append(float(crit.alpha))
    out['ws_steps'] = np.array(steps)
    out['ws_alpha'] = np.array(vals)
    for mode in ('linear', 'exp', 'jump'):
        crit = ref.KLDLoss(alpha=2.0, tau=1, warmup_config={'mode': mode, 'warmup_iters': 10},
                           earlydecay_config={'mode': mode, 'earlydecay_start': 20, 'earlydecay_end': 30})
        seq = list(range(0, 36))
        vals = []
        for n in seq:
            crit(x, x.clone(), gt, n)
            vals.append(float(crit.alpha))
        out[f'{mode}_steps'] = np.array(seq)
        out[f'{mode}_alpha'] = np.array(vals)
    return out


def smoke_values(ref):
    """The survey's cfg1 smoke numbers: seed 0, randn 2x150x64x64, n_iter=1."""
    torch.manual_seed(0)
    s = torch.randn(2, 150, 64, 64)
    t = torch.randn(2, 150, 64, 64)
    gt = torch.zeros(2, 1, 64, 64, dtype=torch.int64)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for cls in ('CDLoss', 'PDLoss', 'CGDLoss', 'ATLoss'):
            x = s.clone().requires_grad_(True)
            loss = getattr(ref, cls)()(x, t, gt, 1)
            loss.backward()
            out[cls + '_loss'] = np.array(loss.item(), dtype=np.float64)
            out[cls + '_grad_abs_sum'] = np.array(x.grad.double().abs().sum().item())
            out[cls + '_grad_probe'] = x.grad[1, 77, 13, 5:13].numpy().copy()
    return out


def at_cases(ref):
    out = {}
    g = torch.Generator().manual_seed(77)
    s = torch.randn(2, 6, 5, 8, generator=g).requires_grad_(True)
    t = torch.randn(2, 6, 5, 8, generator=g)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        loss = ref.ATLoss()(s, t, None, 0)
    loss.backward()
    out.update(S=s.detach().numpy(), T=t.numpy(), loss=np.array(loss.item(), dtype=np.float64),
               grad=s.grad.numpy())
    return out


def ifvd_cases(ref):
    """IFVDLoss (losses.py:199-238): labels at twice the feature resolution (nearest-resized inside), with
    ignore pixels (255) and a class that never occurs."""
    out = {}
    g = torch.Generator().manual_seed(88)
    s = torch.randn(2, 5, 6, 8, generator=g).requires_grad_(True)
    t = torch.randn(2, 5, 6, 8, generator=g)
    target = torch.randint(0, 4, (2, 1, 12, 16), generator=g)
    target[0, 0, :3, :5] = 255
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        loss = ref.IFVDLoss()(s, t, target, 0)
    loss.backward()
    out.update(S=s.detach().numpy(), T=t.numpy(), target=target.numpy(), loss=np.array(loss.item(), dtype=np.float64),
               grad=s.grad.numpy())
    return out


def load_reference_seg_losses(ref_root=REF):
    """The reference's supervised-loss modules, unmodified, loaded by path: ``mmseg/models/losses/cross_entropy_loss.py``
    (+ its ``utils.py``) and ``accuracy.py``, and ``resize`` from ``mmseg/ops/wrappers.py``.  ``import mmseg`` needs mmcv
    (absent), so stub packages stand in for ``mmseg``, ``mmseg.models``, ``mmseg.models.losses`` and
    ``mmseg.models.builder`` (whose ``LOSSES.register_module()`` - an mmcv Registry in the reference - is an identity
    decorator here)."""
    names = ['mmseg', 'mmseg.models', 'mmseg.models.builder', 'mmseg.models.losses', 'mmseg.models.losses.utils',
             'mmseg.models.losses.cross_entropy_loss', 'mmseg.models.losses.accuracy']
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for pkg in names[:4]:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m

        class _Registry:
            def register_module(self, *a, **k):
                return lambda cls: cls
        sys.modules['mmseg.models.builder'].LOSSES = _Registry()
        mods = {}
        for short in ('utils', 'cross_entropy_loss', 'accuracy'):
            full = 'mmseg.models.losses.' + short
            spec = importlib.util.spec_from_file_location(full, os.path.join(ref_root, 'mmseg/models/losses', short + '.py'))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
            mods[short] = mod
        wrappers = types.ModuleType('mmseg.ops.wrappers')
        with open(os.path.join(ref_root, 'mmseg/ops/wrappers.py')) as f:
            exec(compile(f.read(), 'mmseg/ops/wrappers.py', 'exec'), wrappers.__dict__)
        return mods['cross_entropy_loss'].CrossEntropyLoss, mods['accuracy'].accuracy, wrappers.resize
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# (case, logits shape, scale, CrossEntropyLoss kwargs, call kwargs, ignore band, label kind)
SEG_CASES = [
    ('x4_plain',        (2, 7, 6, 8),   4, {}, {}, True),
    ('x1_same_size',    (2, 5, 9, 7),   1, {}, {}, True),
    ('x2_class_weight', (1, 6, 8, 8),   2, dict(class_weight=[0.5, 1.0, 2.0, 1.5, 0.25, 3.0], loss_weight=0.4), {}, True),
    ('x8_avg_factor',   (1, 4, 5, 6),   8, {}, dict(avg_factor=777.0), True),
    ('x4_sum',          (1, 19, 4, 4),  4, dict(reduction='sum', loss_weight=0.01), {}, False),
    ('x2_pixel_weight', (2, 3, 6, 6),   2, {}, 'weight', True),
]


def seg_loss_cases():
    """BaseDecodeHead.losses (decode_head.py:217-237) with the reference's own resize, CrossEntropyLoss and accuracy,
    called in that order; ignore_index=255, align_corners=False (BaseDecodeHead defaults, decode_head.py:52-54)."""
    CE, accuracy, resize = load_reference_seg_losses()
    out = {}
    for n, (name, shape, scale, kw, call_kw, ignore_band) in enumerate(SEG_CASES):
        g = torch.Generator().manual_seed(300 + n)
        b, c, h, w = shape
        logit = (torch.randn(shape, generator=g) * 2.0).requires_grad_(True)
        label = torch.randint(0, c, (b, 1, h * scale, w * scale), generator=g)
        if ignore_band:
            label[:, :, : max(1, h * scale // 5), : w * scale // 2] = 255
        weight = None
        if call_kw == 'weight':
            weight = torch.rand(b, h * scale, w * scale, generator=g)
            call_kw = {}
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            seg_logit = resize(input=logit, size=label.shape[2:], mode='bilinear', align_corners=False)    # :221-225
            seg_label = label.squeeze(1)                                                                  # :230
            loss = CE(**kw)(seg_logit, seg_label, weight=weight, ignore_index=255, **call_kw)            # :231-235
            acc = accuracy(seg_logit, seg_label)                                                          # :236
        loss.backward()
        out[name] = dict(logit=logit.detach().numpy(), label=label.numpy(), scale=np.array(scale),
                         loss=np.array(loss.item(), dtype=np.float64), acc=np.array(float(acc), dtype=np.float64),
                         grad=logit.grad.numpy(), ce_kwargs=np.array(repr(kw)), call_kwargs=np.array(repr(call_kw)),
                         weight=(weight.numpy() if weight is not None else np.zeros(0, dtype=np.float32)))
    return out


def main():
    if sys.argv[1:] == ['segloss']:            # add these fixtures without rewriting the others
        for name, rec in seg_loss_cases().items():
            np.savez_compressed(os.path.join(HERE, f'segloss_{name}.npz'), **rec)
            print(f'segloss_{name:18s} loss={float(rec["loss"]):.9f} acc={float(rec["acc"]):.4f}')
        return
    ref = load_reference_losses()
    if sys.argv[1:] == ['ifvd']:               # add this fixture without rewriting the others
        np.savez_compressed(os.path.join(HERE, 'ifvd_2x5x6x8.npz'), **ifvd_cases(ref))
        return
    for i, case in enumerate(CASES):
        rec = run_case(ref, case, seed=100 + i)
        np.savez_compressed(os.path.join(HERE, f'kld_{case[0]}.npz'), **rec)
        print(f'{case[0]:22s} loss={float(rec["loss"]):.9f}')
    np.savez_compressed(os.path.join(HERE, 'schedules.npz'), **schedule_table(ref))
    np.savez_compressed(os.path.join(HERE, 'smoke_cfg1.npz'), **smoke_values(ref))
    np.savez_compressed(os.path.join(HERE, 'atloss_2x6x5x8.npz'), **at_cases(ref))
    np.savez_compressed(os.path.join(HERE, 'ifvd_2x5x6x8.npz'), **ifvd_cases(ref))
    for name, rec in seg_loss_cases().items():
        np.savez_compressed(os.path.join(HERE, f'segloss_{name}.npz'), **rec)
    print('torch', torch.__version__, 'reference', REF)


if __name__ == '__main__':
    main()
