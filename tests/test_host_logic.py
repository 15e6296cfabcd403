"""Host-side mirror of the reference interface (CPU): schedules, RNG use, dispatcher naming."""
import numpy as np
import pytest
import torch

import segdistill_b200 as sd
from segdistill_b200 import dist as sdist
from helpers import load_golden


def test_constructor_signatures_match_reference():
    import inspect
    sig = inspect.signature(sd.KLDLoss.__init__)
    assert list(sig.parameters)[1:] == ['alpha', 'tau', 'resize_config', 'shuffle_config', 'transform_config',
                                        'warmup_config', 'earlydecay_config']
    assert [p.default for p in list(sig.parameters.values())[1:]] == [1, 1, None, None, None, None, None]
    sig = inspect.signature(sd.CGDLoss.__init__)
    assert [(k, v.default) for k, v in list(sig.parameters.items())[1:]] == [('group_size', 10), ('alpha', 3),
                                                                              ('tau', 2)]
    for cls in (sd.PDLoss, sd.CDLoss, sd.CGDLossWS, sd.ATLoss):
        assert len(inspect.signature(cls.__init__).parameters) == 1
    m = sd.CGDLossWS()
    assert (m.alpha_0, m.tau) == (3, 2)
    assert m.shuffle_config == {'interval': 1000}
    assert m.transform_config == {'loss_type': 'channel', 'group_size': 10}
    assert m.warmup_config == {'mode': 'linear', 'warmup_iters': 2000}
    assert m.earlydecay_config == {'mode': 'linear', 'earlydecay_start': 110000, 'earlydecay_end': 120000}
    assert sd.PDLoss().transform_config == {'loss_type': 'pixel'}
    assert sd.CDLoss().resize_config == {'mode': 'bilinear', 'align_corners': False}


def test_alpha_state_machine_matches_reference_fixture():
    z = load_golden('schedules')
    m = sd.CGDLossWS()
    got = []
    for n in z['ws_steps']:
        m._update_alpha(int(n))
        got.append(float(m.alpha))
    np.testing.assert_allclose(got, z['ws_alpha'], rtol=1e-12, atol=0)
    for mode in ('linear', 'exp', 'jump'):
        m = sd.KLDLoss(alpha=2.0, tau=1, warmup_config={'mode': mode, 'warmup_iters': 10},
                       earlydecay_config={'mode': mode, 'earlydecay_start': 20, 'earlydecay_end': 30})
        got = []
        for n in z[f'{mode}_steps']:
            m._update_alpha(int(n))
            got.append(float(m.alpha))
        np.testing.assert_allclose(got, z[f'{mode}_alpha'], rtol=1e-12, atol=0)


def test_zero_alpha_short_circuit_runs_without_a_kernel():
    m = sd.CGDLossWS()                       # warm-up: alpha(0) = 0
    s = torch.randn(1, 20, 4, 4, requires_grad=True)
    loss = m(s, torch.randn(1, 20, 4, 4), torch.zeros(1, 1, 4, 4, dtype=torch.long), 0)
    assert loss.item() == 0.0 and loss.requires_grad
    loss.backward()
    assert s.grad is None or float(s.grad.abs().sum()) == 0.0
    m2 = sd.CGDLossWS()
    loss = m2(s, s.detach(), None, 125000)   # after early decay
    assert loss.item() == 0.0


def test_shuffle_draws_the_same_permutation_as_the_reference():
    rec = load_golden('kld_cgd_shuffle_n1000')
    m = sd.CGDLossWS()                       # alpha == 0 at n_iter 0 -> no kernel, but the draw happens
    m.warmup_config = None
    m.alpha = m.alpha_0 = 0
    torch.manual_seed(int(rec['manual_seed']))
    s = torch.from_numpy(rec['S'])
    m(s, torch.from_numpy(rec['T']), None, 1000)
    assert m.last_perm is not None
    np.testing.assert_array_equal(m.last_perm.numpy(), rec['perm'])
    m(s, torch.from_numpy(rec['T']), None, 1001)
    assert m.last_perm is None               # only when n_iter % interval == 0


def test_dispatcher_builds_and_names_like_the_reference():
    cfg = [
        {'student_layer': 'decode_head.linear_pred', 'teacher_layer': 'decode_head.linear_pred',
         'loss_name': 'CGDLossWS', 'loss_config': {}},
        {'student_layer': 'a', 'teacher_layer': 'b', 'loss_name': 'KLDLoss',
         'loss_config': ({'alpha': 0, 'tau': 2, 'transform_config': {'loss_type': 'channel', 'group_size': 2}},)},
        {'student_layer': 'a', 'teacher_layer': 'c', 'loss_name': 'KLDLoss',
         'loss_config': {'alpha': 0, 'tau': 2, 'transform_config': {'loss_type': 'channel', 'group_size': 2}}},
    ]
    d = sd.DistillationLoss(cfg)
    assert isinstance(cfg[0]['criterion'], sd.CGDLossWS)
    x = torch.randn(1, 4, 2, 2)
    feats = {'decode_head.linear_pred': x, 'a': x}
    featt = {'decode_head.linear_pred': x, 'b': x, 'c': x}
    out = d(feats, featt, None, 0, None, None)       # all alphas are 0 -> no kernel needed on CPU
    # a tuple-wrapped loss_config is unwrapped for the constructor only (opts.py:81-82); the key lookup indexes the
    # tuple itself and lands in the bare except (:105-108) -> 'other'
    assert list(out) == ['loss_decode_head.linear_pred<->decode_head.linear_pred_other',
                         'loss_a<->b_other',
                         "loss_a<->c_{'loss_type': 'channel', 'group_size': 2}"]
    with pytest.raises(NameError):
        sd.build_criterion('NoSuchLoss', {})
    with pytest.raises(TypeError):
        sd.build_criterion('CDLoss', {'alpha': 1})     # CDLoss() takes no kwargs, as in the reference


def test_extractor_records_only_in_training_mode():
    import torch.nn as nn
    stu = nn.Sequential(nn.Conv2d(3, 4, 1), nn.Conv2d(4, 5, 1))
    tea = nn.Sequential(nn.Conv2d(3, 4, 1), nn.Conv2d(4, 5, 1))
    ex = sd.Extractor(stu, tea, [{'student_layer': '1', 'teacher_layer': '1'}])
    x = torch.randn(1, 3, 4, 4)
    ex.train()
    stu(x), tea(x)
    assert ex.student_features['1'].shape == (1, 5, 4, 4) and '1' in ex.teacher_features
    ex.student_features.clear()
    ex.eval()
    stu(x)
    assert ex.student_features == {}


def test_multi_teacher_dispatcher_pairs_entry_i_with_teacher_i():
    """opts.py:127-210: teacher i's hooked layer is recorded as name+str(i); entry i is computed against it and
    named loss_{student}<->{teacher}{i}_{i}; a mismatching number of teacher maps goes to `loss_random`."""
    import torch.nn as nn
    stu = nn.Sequential(nn.Conv2d(3, 4, 1), nn.Conv2d(4, 5, 1))
    teachers = [nn.Sequential(nn.Conv2d(3, 4, 1), nn.Conv2d(4, 5, 1)) for _ in range(2)]
    cfg = [{'student_layer': '1', 'teacher_layer': '1', 'loss_name': 'KLDLoss', 'loss_config': {'alpha': 0}}
           for _ in range(2)]
    ex = sd.ExtractorMT(stu, teachers, cfg)
    ex.train()
    x = torch.randn(1, 3, 4, 4)
    stu(x)
    for t in teachers:
        t(x)
    assert sorted(ex.teacher_features) == ['10', '11'] and list(ex.student_features) == ['1']
    d = sd.DistillationLossMT(cfg)
    out = d(ex.student_features, ex.teacher_features, None, 0)          # alpha 0 -> no kernel needed on CPU
    assert list(out) == ['loss_1<->10_0', 'loss_1<->11_1']
    seen = {}

    class Spy(nn.Module):
        def forward(self, s, t, gt, step):
            seen['teacher'] = t
            return s.sum() * 0

    cfg[0]['criterion'] = Spy()
    out = d(ex.student_features, {'10': ex.teacher_features['10']}, None, 0)
    assert list(out) == ['loss_random'] and isinstance(seen['teacher'], list) and len(seen['teacher']) == 1


def test_shard_bounds_cover_the_batch():
    for batch, world in ((128, 8), (16, 2), (7, 4), (3, 8)):
        pieces = [sdist.shard_bounds(batch, r, world) for r in range(world)]
        assert pieces[0][0] == 0 and pieces[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))


def test_parse_losses_single_process():
    losses = {'loss_seg': torch.tensor(1.5), 'acc': torch.tensor(80.0), 'loss_kd': [torch.tensor(0.25), torch.tensor(0.5)]}
    total, logs = sdist.parse_losses(losses)
    assert total.item() == pytest.approx(2.25)
    assert logs == {'loss_seg': 1.5, 'acc': 80.0, 'loss_kd': 0.75, 'loss': 2.25}


def test_deferred_logs_single_process_ring_wraps_and_reads_back_in_one_buffer():
    """dist.DeferredLogs without a process group: the ring and its cursor are ONE allocation (a flush reads both back
    with one copy), a flush returns the steps since the previous one oldest first, at most ``interval`` of them, and
    'loss' = the sum of the entries whose name contains 'loss' (SD_structure.py:121-122)."""
    dl = sdist.DeferredLogs(['loss_a', 'loss_b', 'acc'], interval=4)
    assert dl.ring.untyped_storage().data_ptr() == dl.cursor.untyped_storage().data_ptr()
    assert dl.cursor.dtype == torch.int32 and dl.ring.shape == (4, 3)
    assert dl.flush() == []
    for i in range(6):                                 # six steps into four slots: the two oldest are overwritten
        dl.push(torch.tensor([float(i), 0.5, 10.0 + i]))
    recs = dl.flush()
    assert [r['loss_a'] for r in recs] == [2.0, 3.0, 4.0, 5.0]
    assert all(r['loss'] == r['loss_a'] + 0.5 and r['acc'] == 10.0 + r['loss_a'] for r in recs)
    dl.push([torch.tensor(7.0), torch.tensor(1.0), torch.tensor(0.25)])    # a list of 0-dim tensors
    assert dl.flush() == [{'loss_a': 7.0, 'loss_b': 1.0, 'acc': 0.25, 'loss': 8.0}]
    assert dl.flush() == []
    assert int(dl.cursor.item()) == 7


def test_autograd_adopts_the_gradient_buffer_without_copy():
    """The pattern functional._finish_backward relies on: dropping ctx's reference lets autograd keep dS."""
    holder = {}

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            ctx.ds = torch.full_like(x, 2.0)
            holder['ptr'] = ctx.ds.data_ptr()
            return x.sum() * 0 + 1.0

        @staticmethod
        def backward(ctx, g):
            ds = ctx.ds
            ctx.ds = None
            return ds

    x = torch.randn(64, requires_grad=True)
    Fn.apply(x).backward()
    assert x.grad.data_ptr() == holder['ptr']


def test_ifvd_class_map_equals_the_reference_masks():
    """IFVDLoss hands the kernel one class index per pixel instead of the reference's C float masks
    (losses.py:218-224): index i where `Upsample(nearest)(target.float()) == i` for an i in range(C), else C."""
    import torch.nn.functional as F
    from segdistill_b200.losses import IFVDLoss
    g = torch.Generator().manual_seed(3)
    b, c, h, w = 2, 7, 5, 6
    target = torch.randint(0, 9, (b, 1, 2 * h, 2 * w), generator=g)       # classes 7, 8 match no channel index
    target[1, 0, 4:, :3] = 255
    cls = IFVDLoss._class_map(target, c, h, w)
    assert cls.dtype == torch.int32 and cls.shape == (b, h * w)
    tar = F.interpolate(target.float(), size=(h, w), mode='nearest').reshape(b, h * w)
    covered = torch.zeros(b, h * w, dtype=torch.bool)
    for i in range(c):
        m = tar == i
        assert bool((cls[m] == i).all())
        covered |= m
    assert bool((cls[~covered] == c).all()) and int((~covered).sum()) > 0


def test_second_backward_through_the_same_node_rebuilds_the_gradient(monkeypatch):
    """retain_graph=True and two backward passes (the reference trainer's log_grad mode, SD_structure.py:92-134):
    the first hands out the buffer forward filled, the second must re-run the kernel - never return None.
    (The C ABI is mocked on the CPU: this is host logic.)"""
    from segdistill_b200 import _cabi
    from segdistill_b200 import functional as SF
    calls = []

    def fake_mse(s, t, weight=1.0, grad_scale=1.0):
        calls.append('mse')
        d = s.detach() - t.detach()
        return weight * (d * d).mean(), 2.0 * weight * d / d.numel()

    def fake_scale(ds, go):
        return ds.mul_(go.detach().reshape(()).to(ds.dtype))

    monkeypatch.setattr(_cabi, 'mse', fake_mse)
    monkeypatch.setattr(_cabi, 'scale_grad_', fake_scale)
    torch.manual_seed(3)
    w = torch.randn(4, 5, requires_grad=True)
    t = torch.randn(4, 5)
    x = w * 2.0                                    # non-leaf student, as in training
    loss = SF.mse_loss(x, t, 0.5)
    want = 2.0 * 2.0 * 0.5 * (2.0 * w.detach() - t) / t.numel()
    (g1,) = torch.autograd.grad(loss, w, retain_graph=True)
    (g2,) = torch.autograd.grad(3.0 * loss, w, retain_graph=True)      # a different upstream weight
    loss.backward()
    assert torch.allclose(g1, want, rtol=1e-6, atol=1e-8)
    assert torch.allclose(g2, 3.0 * want, rtol=1e-6, atol=1e-8)
    assert torch.allclose(w.grad, want, rtol=1e-6, atol=1e-8)
    assert calls == ['mse', 'mse', 'mse']          # forward + one re-run per extra backward
    with pytest.raises(RuntimeError):              # graph freed: autograd's own error, not a silent None
        loss.backward()
