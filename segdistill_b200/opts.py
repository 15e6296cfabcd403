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


class DistillationLoss(nn.Module):
    batch_pairs = True      # serve two KLD entries on the same tensors with one launch when possible

    def __init__(self, distillation):
        super().__init__()
        self.distillation = distillation
        crits = []
        for entry in distillation:
            entry['criterion'] = build_criterion(entry['loss_name'], entry['loss_config'])
            crits.append(entry['criterion'])
        self.criteria = nn.ModuleList(crits)

    def forward(self, student_features, teacher_features, gt_semantic_seg, step, student=None, teacher=None):
        """Same results and key names as the reference loop (:87-112).  One difference in HOW: the host-side
        half of every KLD criterion (schedules, resize, shuffle draw) runs first, in entry order, so that two
        entries hooking the very same tensors (e.g. CD + CGD on the logits) can be served by one two-loss
        kernel launch - one read of the maps, one write of the summed gradient."""
        out = {}
        plans, resized = {}, {}
        for i, entry in enumerate(self.distillation):
            crit = entry['criterion']
            if isinstance(entry['student_layer'], list) or not isinstance(crit, _losses.KLDLoss):
                continue
            plans[i] = crit.plan(student_features[entry['student_layer']], teacher_features[entry['teacher_layer']],
                                 gt_semantic_seg, step, resized)
        results, pending = {}, list(plans)
        while pending:
            i = pending.pop(0)
            mate = next((j for j in pending if _losses.KLDLoss.can_fuse(plans[i], plans[j])), None) \
                if self.batch_pairs else None
            if mate is None:
                results[i] = _losses.KLDLoss.run(plans[i])
            else:
                pending.remove(mate)
                results[i], results[mate] = _losses.KLDLoss.run_pair(plans[i], plans[mate])
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
        return out


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
