#!/usr/bin/env python
"""Kernel-level timings through the C ABI (no autograd, no modules): CUDA events over back-to-back launches.

    python scripts/kbench.py [--iters 50] [--only cd_f32,fused_f32]
Prints one line per case: microseconds per launch, achieved GB/s on the algorithmic bytes
(read S + read T + write dS) and the fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segdistill_b200 import _cabi  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=50)
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    pk = peak()

    def pair(shape, dtype):
        g = torch.Generator(device=dev).manual_seed(0)
        return (torch.randn(shape, device=dev, generator=g).to(dtype),
                torch.randn(shape, device=dev, generator=g).to(dtype))

    L = (16, 150, 128, 128)
    cases = {
        'cd_f32': (L, torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=1)),
        'cd_bf16': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows(s, t, group=1)),
        'cgd10_f32': (L, torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0)),
        'cgd10_bf16': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0)),
        'fused_f32': (L, torch.float32, lambda s, t: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0))),
        'fused_bf16': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0))),
        'cgd10_f32_cluster': (L, torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0, algo=4)),
        'cgd10_bf16_cluster': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0, algo=4)),
        'cd_f32_cluster': (L, torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=1, algo=4)),
        'cd_bf16_cluster': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows(s, t, group=1, algo=4)),
        'fused_f32_cluster': (L, torch.float32,
                              lambda s, t: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0), algo=4)),
        'fused_bf16_cluster': (L, torch.bfloat16,
                               lambda s, t: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0), algo=4)),
        'fused_f32_grid': (L, torch.float32,
                           lambda s, t: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0), algo=6)),
        'fused_bf16_grid': (L, torch.bfloat16,
                            lambda s, t: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0), algo=6)),
        'cgd10_f32_grid': (L, torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0, algo=6)),
        'cgd10_bf16_grid': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0, algo=6)),
        'cd_f32_grid': (L, torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=1, algo=6)),
        'cd_bf16_grid': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows(s, t, group=1, algo=6)),
        'cgd10_f32_stream': (L, torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0, algo=3)),
        'cgd10_bf16_stream': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0, algo=3)),
        'fused_bf16_stream': (L, torch.bfloat16,
                              lambda s, t: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0), algo=3)),
        'fused_f32_stream': (L, torch.float32,
                             lambda s, t: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0), algo=3)),
        'corr10_f32': (L, torch.float32, lambda s, t: _cabi.cgd_corr(s, t, group=10)),
        'corr10_bf16': (L, torch.bfloat16, lambda s, t: _cabi.cgd_corr(s, t, group=10)),
        'corr150_bf16': (L, torch.bfloat16, lambda s, t: _cabi.cgd_corr(s, t, group=150)),
        'corr256_512ch_bf16': ((16, 512, 64, 64), torch.bfloat16, lambda s, t: _cabi.cgd_corr(s, t, group=256)),
        'corr256_512ch_f32': ((16, 512, 64, 64), torch.float32, lambda s, t: _cabi.cgd_corr(s, t, group=256)),
        'up4_cgd10_2x150x128_f32': ((2, 150, 128, 128), torch.float32, lambda s, t: _cabi.kl_rows_up(s, t, 4, group=10, tau=2.0, alpha=3.0)),
        'up4_cgd10_16x150x128_f32': (L, torch.float32, lambda s, t: _cabi.kl_rows_up(s, t, 4, group=10, tau=2.0, alpha=3.0)),
        'up8_cd_16x150x64_f32': ((16, 150, 64, 64), torch.float32, lambda s, t: _cabi.kl_rows_up(s, t, 8, group=1)),
        'up4_cgd10_16x150x128_bf16': (L, torch.bfloat16, lambda s, t: _cabi.kl_rows_up(s, t, 4, group=10, tau=2.0, alpha=3.0)),
        'pd_f32': (L, torch.float32, lambda s, t: _cabi.kl_pixels(s, t)),
        'pd_bf16': (L, torch.bfloat16, lambda s, t: _cabi.kl_pixels(s, t)),
        'cd_512ch_f32': ((16, 512, 64, 64), torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=1, tau=4.0)),
        'cd+mse_512ch_f32': ((16, 512, 64, 64), torch.float32,
                             lambda s, t: _cabi.kl_rows(s, t, group=1, tau=4.0, mse_weight=1.0)),
        'mse_512ch_f32': ((16, 512, 64, 64), torch.float32, lambda s, t: _cabi.mse(s, t)),
        'cd_cfg1_f32': ((2, 150, 64, 64), torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=1)),
        'cgd150_f32': ((16, 150, 64, 64), torch.float32, lambda s, t: _cabi.kl_rows(s, t, group=150, tau=2.0)),
    }
    def blocky(shape, block=8):
        b, c, h, w = shape
        g = torch.Generator(device=dev).manual_seed(1)
        coarse = torch.randint(0, c, (b, 1, h // block, w // block), device=dev, generator=g)
        return coarse.repeat_interleave(block, 2).repeat_interleave(block, 3).reshape(b, h * w).to(torch.int32)

    def torch_ifvd_sim(s, t, cls):
        """the similarity term as ATen ops under autograd (scatter-add centres) - what ifvd.cu replaced"""
        import torch.nn.functional as F
        b, c = s.shape[:2]
        k = cls.long()
        x = s.reshape(b, c, -1).detach().requires_grad_(True)

        def sim(f):
            gi = k.unsqueeze(1).expand_as(f)
            sums = f.new_zeros(b, c, c + 1).scatter_add_(2, gi, f)
            cnt = f.new_zeros(b, c + 1).scatter_add_(1, k, torch.ones_like(k, dtype=f.dtype))
            return F.cosine_similarity(f, torch.gather(sums / (cnt.unsqueeze(1) + 1e-6), 2, gi), dim=1)

        (10 * F.mse_loss(sim(x), sim(t.reshape(b, c, -1)))).backward()

    S2 = (2, 150, 128, 128)
    cls2, cls16, cls16c = blocky(S2), blocky(L), blocky(L, 32)   # label blocks of 8x8 / 32x32 pixels
    cases.update({
        'ifvd_sim_2x150x128_f32': (S2, torch.float32, lambda s, t: _cabi.ifvd_sim(s, t, cls2)),
        'ifvd_sim_2x150x128_aten': (S2, torch.float32, lambda s, t: torch_ifvd_sim(s, t, cls2)),
        'ifvd_sim_16x150x128_f32': (L, torch.float32, lambda s, t: _cabi.ifvd_sim(s, t, cls16)),
        'ifvd_sim_16x150x128_f32_coarse': (L, torch.float32, lambda s, t: _cabi.ifvd_sim(s, t, cls16c)),
        'ifvd_sim_16x150x128_bf16': (L, torch.bfloat16, lambda s, t: _cabi.ifvd_sim(s, t, cls16)),
        'ifvd_sim_16x150x128_aten': (L, torch.float32, lambda s, t: torch_ifvd_sim(s, t, cls16)),
    })
    # the student head's supervised loss: logits at 1/4 resolution, 512x512 labels
    lab2 = torch.randint(0, 150, (2, 512, 512), device=dev)
    lab16 = torch.randint(0, 150, (16, 512, 512), device=dev)

    def torch_seg_loss(s, lab):
        """decode_head.py:217-237 as ATen ops under autograd - what ce_up.cu replaced"""
        import torch.nn.functional as F
        x = s.detach().requires_grad_(True)
        up = F.interpolate(x, size=lab.shape[1:], mode='bilinear', align_corners=False)
        loss = F.cross_entropy(up, lab, reduction='none', ignore_index=255).mean()
        (up.argmax(1) == lab).float().mean()
        loss.backward()

    cases.update({
        'ce_up4_2x150x128_f32': (S2, torch.float32, lambda s, t: _cabi.ce_up(s, lab2, 4)),
        'ce_up4_2x150x128_aten': (S2, torch.float32, lambda s, t: torch_seg_loss(s, lab2)),
        'ce_up4_16x150x128_f32': (L, torch.float32, lambda s, t: _cabi.ce_up(s, lab16, 4)),
        'ce_up4_16x150x128_bf16': (L, torch.bfloat16, lambda s, t: _cabi.ce_up(s, lab16, 4)),
    })
    # BASELINE config 2: CGD on the four MiT-B0 stage maps, one grouped launch vs four launches
    st_shapes = [(16, 32, 128, 128), (16, 64, 64, 64), (16, 160, 32, 32), (16, 256, 16, 16)]
    st = [pair(sh, torch.float32) for sh in st_shapes]
    st_bytes = sum(3 * s.numel() * 4 for s, _ in st)

    def cfg2_grouped(s, t):
        _cabi.kl_rows_group([x for x, _ in st], [y for _, y in st], (10,) * 4, (2.0,) * 4, (3.0,) * 4)

    def cfg2_separate(s, t):
        for x, y in st:
            _cabi.kl_rows(x, y, group=10, tau=2.0, alpha=3.0)
    cases.update({'cfg2_grouped_f32': ((1, 1, 4, 4), torch.float32, cfg2_grouped),
                  'cfg2_separate_f32': ((1, 1, 4, 4), torch.float32, cfg2_separate)})
    only = [x for x in a.only.split(',') if x]
    for name, (shape, dtype, fn) in cases.items():
        if only and name not in only:
            continue
        s, t = pair(shape, dtype)
        try:
            for _ in range(5):
                fn(s, t)
        except _cabi.SegDistillUnsupported:
            print(f'{name:20s} (the forced kernel does not take this layout)', flush=True)
            continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn(s, t)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / a.iters * 1e3
        nbytes = st_bytes if name.startswith('cfg2_') else 3 * s.numel() * s.element_size()
        gbs = nbytes / us / 1e3
        print(f'{name:20s} {us:9.1f} us  {gbs:8.1f} GB/s  {gbs / pk:6.3f} of measured peak   [{_cabi.last_kernel()}]',
              flush=True)
        assert _cabi.workspace_error_flag() == 0
        del s, t


if __name__ == '__main__':
    main()
