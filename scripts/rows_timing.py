#!/usr/bin/env python
"""Where kl_rows_tma_kernel's time goes: per-phase cycle counters of threads 0 and 480 of every CTA, written by a
build with -DSD_ROWS_TIMING (SD_NVCC_EXTRA=-DSD_ROWS_TIMING python -m segdistill_b200.build --force).

    python scripts/rows_timing.py [bf16]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segdistill_b200 import _cabi  # noqa: E402

dev = torch.device('cuda', 0)
shape = (int(os.environ.get('SD_B', '16')), 150, 128, 128)
dtype = torch.bfloat16 if 'bf16' in sys.argv else torch.float32
g = torch.Generator(device=dev).manual_seed(0)
s = torch.randn(shape, device=dev, generator=g).to(dtype)
t = torch.randn(shape, device=dev, generator=g).to(dtype)
for _ in range(3):
    _cabi.kl_rows(s, t, group=1)
torch.cuda.synchronize()
print(_cabi.last_kernel())
ws = next(iter(_cabi._workspaces.values()))
off = 256 + 4 * 3 * 1024
n_cta = 148
v = ws[off:off + n_cta * 16 * 8].view(torch.int64).view(n_cta, 16).cpu().double()
names = ['ring->regs+max', 'barrier 1 (+issue)', 'exponentials', 'warp reduce', 'barrier 2 + merge', 'gradient + stores']
if _cabi.last_kernel() == 'kl_rows_rm_kernel':
    names = ['wait next row', 'exponentials + loads', 'warp reduce, publish', 'barrier (+issue)', 'merge', 'gradient + stores']
for base, who in ((0, 'thread 0'), (8, 'thread 480')):
    rows = v[:, base + 6].clamp(min=1)
    print(f'-- {who}: rows per CTA {rows.mean():.1f}, total cycles {v[:, base + 7].mean():.0f}')
    for i, n in enumerate(names):
        per = v[:, base + i] / rows
        print(f'   {n:20s} {per.mean():8.0f} cycles per row  (min {per.min():6.0f} max {per.max():6.0f})')
    print(f'   {"sum":20s} {(v[:, base:base + 6].sum(1) / rows).mean():8.0f}')
