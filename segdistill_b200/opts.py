"""Dispatcher mirror of the reference's ``mmseg/models/distillation/opts.py``.

``Extractor`` (:13-71) records hooked layer outputs; ``DistillationLoss`` (:74-112) builds one
criterion per ``distillation`` entry from ``loss_name`` + ``loss_config`` and names each result
``loss_{student_layer}<->{teacher_layer}_{loss_info}`` (:105-110).  The reference resolves
``loss_name`` with ``eval`` in a namespace star-imported from its losses module; here the same
bare names resolve through ``LOSS_CLASSES`` (register extra classes there).
"""
from __future__ import annotations

from functools import partial

import torch.nn as nn

from . import losses as _losses

LOSS_CLASSES = {name: getattr(_losses, name) for name in _losses.__all__}


def build_criterion(loss_name, loss_config):
    if isinstance(loss_config, tuple):          # reference unwraps tuple-wrapped configs (:81-82)
        loss_config = loss_config[0]
    try:
        cls = LOSS_CLASSES[loss_name]
    except KeyError:
        raise NameError(f"name '{loss_name}' is not defined (known losses: {sorted(LOSS_CLASSES)})")
    return cls(**loss_config)


class Extractor(nn.Module):
    """Forward hooks on the named student / teacher layers; outputs are kept only in training mode."""

    def __init__(self, student, teacher, distillation):
        super().__init__()
        self.teacher_features = {}
        self.student_features = {}
        want_s, want_t = [], []
        for entry in distillation:
            for names, bucket in ((entry['student_layer'], want_s), (entry['teacher_layer'], want_t)):
                bucket.extend(names if isinstance(names, list) else [names])
        for role, model, wanted in (('teacher', teacher, want_t), ('student', student, want_s)):
            for name, module in model.named_modules():
                if name in wanted:
                    module.register_forward_hook(partial(self._record, name=name, role=role))

    def _record(self, module, inputs, output, name, role):
        if self.training:
            (self.student_features if role == 'student' else self.teacher_features)[name] = output


class _StepRecipe:
    """What one dispatcher step launches, decided once per (shapes, dtypes, pairing) and replayed while it stays valid.

    The slow path re-derives every step what cannot change between steps of a training run: which entries share their
    tensors and fuse into one launch, which kernel serves them, the result keys.  A recipe records that for steps
    whose host-side decisions are static - no alpha schedule (warm-up / early decay), no channel shuffle on this step,
    no host-side resize - so that a step is: fetch the hooked tensors, check they still look the same, one autograd
    node per launch.  Anything else (a shuffle step, a shape change, a stateful criterion) takes the slow path, which
    is always correct."""

    def __init__(self, sig, gt_hw, ops, keys, shuffle_intervals, batch_pairs):
        self.sig, self.gt_hw, self.ops, self.keys = sig, gt_hw, ops, keys
        self.shuffle_intervals, self.batch_pairs = shuffle_intervals, batch_pairs

    def run(self, owner, student_features, teacher_features, gt, step):
        if owner.batch_pairs != self.batch_pairs:
            return None
        for interval in self.shuffle_intervals:
            if step % interval == 0:
                return None                       # the reference draws a channel permutation on this step
        if (None if gt is None else tuple(gt.shape[2:])) != self.gt_hw:
            return None
        xs, xt = [], []
        for s_name, t_name, shape, dtype, device in self.sig:
            a, b = student_features[s_name], teacher_features[t_name]
            if a.shape != shape or a.dtype != dtype or a.device != device or b.shape != shape or b.dtype != dtype:
                return None
            xs.append(a)
            xt.append(b)
        results = [None] * len(self.sig)
        for kind, idx, fn in self.ops:
            if kind == 'pair':
                i, j = idx
                if xs[i] is not xs[j] or xt[i] is not xt[j]:
                    return None
                results[i], results[j] = fn(xs[i], xt[i])
            elif kind == 'group':
                for k, loss in zip(idx, fn([(xs[k], xt[k]) for k in idx])):
                    results[k] = loss
            else:
                results[idx] = fn(xs[idx], xt[idx], gt, step)
        return dict(zip(self.keys, results))


def _static_kld_op(crit, plan):
    """A planned KLD call as a function of the two maps alone, or None when the plan holds per-step host work."""
    from . import functional as SF
    if crit.warmup_config or crit.earlydecay_config or plan['perm'] is not None or crit.algo != 'auto':
        return None
    kind, tau, alpha, group, up = plan['kind'], plan['tau'], plan['alpha'], plan['group'], plan.get('upscale')
    if crit.resize_config and not up and plan['student'].shape[2:] != plan['student_in'].shape[2:]:
        return None                               # resized on the host: stays on the slow path
    if alpha == 0:
        return lambda xs, xt, gt, step: SF.zero_loss(xs)
    if up and kind == 'pixel':
        return lambda xs, xt, gt, step: SF.kl_pixels_up_loss(xs, xt, up, tau=tau, alpha=alpha)
    if up:
        return lambda xs, xt, gt, step: SF.kl_rows_up_loss(xs, xt, up, group=group, tau=tau, alpha=alpha)
    if kind == 'pixel':
        return lambda xs, xt, gt, step: SF.kl_pixels_loss(xs, xt, tau=tau, alpha=alpha)
    if kind == 'channel':
        return lambda xs, xt, gt, step: SF.kl_rows_loss(xs, xt, group=group, tau=tau, alpha=alpha)
    return None


class DistillationLoss(nn.Module):
    batch_pairs = True      # serve two KLD entries on the same tensors with one launch when possible
    cache_steps = True      # replay the step's launch decisions while shapes / pairing stay the same (_StepRecipe)
    batch_groups = True     # serve the channel-mode entries on DIFFERENT tensors with one grouped launch ...
    group_max_elements = 32 * 1024 * 1024   # ... while they are small: above this, one tuned launch per entry wins

    def __init__(self, distillation):
        super().__init__()
        self.distillation = distillation
        crits = []
        for entry in distillation:
            entry['criterion'] = build_criterion(entry['loss_name'], entry['loss_config'])
            crits.append(entry['criterion'])
        self.criteria = nn.ModuleList(crits)
        self._recipe = None

    def forward(self, student_features, teacher_features, gt_semantic_seg, step, student=None, teacher=None):
        """Same results and key names as the reference loop (:87-112).  One difference in HOW: the host-side
        half of every KLD criterion (schedules, resize, shuffle draw) runs first, in entry order, so that two
        entries hooking the very same tensors (e.g. CD + CGD on the logits) can be served by one two-loss
        kernel launch - one read of the maps, one write of the summed gradient."""
        if self.cache_steps and self._recipe is not None:
            try:
                out = self._recipe.run(self, student_features, teacher_features, gt_semantic_seg, step)
            except _losses._cabi.SegDistillUnsupported:      # (a pair the library declined: the slow path splits it)
                out, self.cache_steps = None, False
            if out is not None:
                return out
        out = {}
        plans, resized = {}, {}
        for i, entry in enumerate(self.distillation):
            crit = entry['criterion']
            if isinstance(entry['student_layer'], list) or not isinstance(crit, _losses.KLDLoss):
                continue
            plans[i] = crit.plan(student_features[entry['student_layer']], teacher_features[entry['teacher_layer']],
                                 gt_semantic_seg, step, resized)
        results, pending, pairs, singles = {}, list(plans), [], []
        while pending:
            i = pending.pop(0)
            mate = next((j for j in pending if _losses.KLDLoss.can_fuse(plans[i], plans[j])), None) \
                if self.batch_pairs else None
            if mate is None:
                singles.append(i)
            else:
                pending.remove(mate)
                pairs.append((i, mate))
                results[i], results[mate] = _losses.KLDLoss.run_pair(plans[i], plans[mate])
        # entries on different tensors: one grouped launch for the channel-mode ones (same dtype / device), while small
        group = [i for i in singles if _losses.KLDLoss.can_group(plans[i])] if self.batch_groups else []
        if group:
            x0 = plans[group[0]]['student']
            group = [i for i in group if plans[i]['student'].dtype == x0.dtype and plans[i]['student'].device == x0.device]
            group = group[:_losses._cabi.MAX_GROUP_PAIRS]
        if len(group) < 2 or sum(plans[i]['student'].numel() for i in group) > self.group_max_elements:
            group = []
        if group:
            for i, loss in zip(group, _losses.KLDLoss.run_group([plans[i] for i in group])):
                results[i] = loss
        for i in singles:
            if i not in results:
                results[i] = _losses.KLDLoss.run(plans[i])
        for i, entry in enumerate(self.distillation):
            s_name, t_name = entry['student_layer'], entry['teacher_layer']
            crit = entry['criterion']
            if isinstance(s_name, list):
                # attention/value pair form of the reference (:92-99); no shipped loss uses it
                loss = crit(student_features[s_name[0]], student_features[s_name[1]],
                            teacher_features[t_name[0]], teacher_features[t_name[1]],
                            student, teacher, gt_semantic_seg, step)
                out[f"loss_{s_name[0]}<->{t_name}_{entry['loss_name']}"] = loss
                continue
            if i in results:
                loss = results[i]
            else:
                loss = crit(student_features[s_name], teacher_features[t_name], gt_semantic_seg, step)
            # the reference indexes entry['loss_config'] ITSELF (:105-108, bare except): a tuple-wrapped config, or one
            # without the key, is named 'other' - kept, so that log / checkpoint keys are the same
            cfg = entry['loss_config']
            info = cfg['transform_config'] if isinstance(cfg, dict) and 'transform_config' in cfg else 'other'
            out[f'loss_{s_name}<->{t_name}_{info}'] = loss
        if self.cache_steps:
            self._recipe = self._build_recipe(student_features, teacher_features, gt_semantic_seg, plans, pairs, list(out),
                                              group)
        return out

    def _build_recipe(self, student_features, teacher_features, gt, plans, pairs, keys, group=()):
        """The step just run as a _StepRecipe, or None when some entry needs per-step host work."""
        from . import functional as SF
        sig, ops, intervals = [], [], []
        paired = {i for pr in pairs for i in pr} | set(group)
        for i, entry in enumerate(self.distillation):
            s_name, t_name, crit = entry['student_layer'], entry['teacher_layer'], entry['criterion']
            if isinstance(s_name, list) or len(keys) != len(self.distillation):
                return None
            x = student_features[s_name]
            sig.append((s_name, t_name, x.shape, x.dtype, x.device))
            if not x.is_cuda:
                return None
            if isinstance(crit, _losses.KLDLoss):
                if crit.shuffle_config:
                    intervals.append(crit.shuffle_config['interval'])
                if i in paired:
                    if crit.warmup_config or crit.earlydecay_config:
                        return None
                    continue
                plan = dict(plans[i], student_in=x)
                fn = _static_kld_op(crit, plan)
                if fn is None:
                    return None
                ops.append(('one', i, fn))
            else:
                ops.append(('one', i, lambda xs, xt, gt, step, crit=crit: crit(xs, xt, gt, step)))
        for i, j in pairs:
            pa, pb = plans[i], plans[j]
            args = (int(pa['group']), float(pa['tau']), float(pa['alpha']), int(pb['group']), float(pb['tau']), float(pb['alpha']))
            ops.append(('pair', (i, j), lambda xs, xt, args=args: SF._KLRowsMulti.apply(xs, xt, *args)))
        if group:
            gp = [plans[i] for i in group]
            if any(crit.resize_config and p['student'].shape[2:] != student_features[self.distillation[i]['student_layer']].shape[2:]
                   for i, p, crit in ((i, plans[i], self.distillation[i]['criterion']) for i in group)):
                return None                           # resized on the host: stays on the slow path
            cfg = ([p['group'] for p in gp], [p['tau'] for p in gp], [p['alpha'] for p in gp])
            ops.append(('group', tuple(group), lambda pairs_, cfg=cfg: SF.kl_rows_group_loss(pairs_, *cfg)))
        gt_hw = None if gt is None else tuple(gt.shape[2:])
        return _StepRecipe(sig, gt_hw, ops, keys, intervals, self.batch_pairs)


class ExtractorMT(nn.Module):
    """Multi-teacher hooks (reference :127-168): teacher i's layer `name` is recorded under `name + str(i)`."""

    def __init__(self, student, teachers, distillation):
        super().__init__()
        self.num_teacher = len(teachers)
        self.teacher_features = {}
        self.student_features = {}
        want_s, want_t = [], []
        for entry in distillation:
            for names, bucket in ((entry['student_layer'], want_s), (entry['teacher_layer'], want_t)):
                bucket.extend(names if isinstance(names, list) else [names])
        for i, teacher in enumerate(teachers):
            for name, module in teacher.named_modules():
                if name in want_t:
                    module.register_forward_hook(partial(self._record, name=name + str(i), role='teacher'))
        for name, module in student.named_modules():
            if name in want_s:
                module.register_forward_hook(partial(self._record, name=name, role='student'))

    def _record(self, module, inputs, output, name, role):
        if self.training:
            (self.student_features if role == 'student' else self.teacher_features)[name] = output


class DistillationLossMT(nn.Module):
    """Multi-teacher dispatcher (reference :170-210): entry i pairs the student layer with teacher i's layer
    (`teacher_layer + str(i)`), result key `loss_{student_layer}<->{teacher_layer}{i}_{i}`.  When the number of
    recorded teacher maps differs from the number of entries, the first criterion receives the LIST of all
    teacher maps under the key `loss_random` (:186-198; no shipped loss accepts a list - kept for fidelity)."""

    def __init__(self, distillation):
        super().__init__()
        self.distillation = distillation
        crits = []
        for entry in distillation:
            entry['criterion'] = build_criterion(entry['loss_name'], entry['loss_config'])
            crits.append(entry['criterion'])
        self.criteria = nn.ModuleList(crits)

    def forward(self, student_features, teacher_features, gt_semantic_seg, step):
        out = {}
        if len(teacher_features) != len(self.distillation):
            entry = self.distillation[0]
            x_teacher = [teacher_features[k] for k in teacher_features]
            out['loss_random'] = entry['criterion'](student_features[entry['student_layer']], x_teacher,
                                                    gt_semantic_seg, step)
            return out
        for i, entry in enumerate(self.distillation):
            s_name = entry['student_layer']
            t_name = entry['teacher_layer'] + str(i)
            out[f'loss_{s_name}<->{t_name}_{i}'] = entry['criterion'](student_features[s_name], teacher_features[t_name],
                                                                     gt_semantic_seg, step)
        return out
