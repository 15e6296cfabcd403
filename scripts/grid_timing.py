#!/usr/bin/env python
"""Where the grid-resident kernel's time goes: per-CTA cycle counters written by a build with -DSD_GRID_TIMING
(SD_NVCC_EXTRA=-DSD_GRID_TIMING python -m segdistill_b200.build --force); unit size via SEGDISTILL_GRID_UNIT.

park warp 0   : wait for a TMEM slot | wait for the ring | total
grad warp 8   : wait for the row statistics | total
gather warp 0 : wait own park | poll packets | max + Z -> fin | KL terms | latency park done -> fin ready | units
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
nl1 = 'cgd' in sys.argv
g = torch.Generator(device=dev).manual_seed(0)
s = torch.randn(shape, device=dev, generator=g).to(dtype)
t = torch.randn(shape, device=dev, generator=g).to(dtype)
for _ in range(3):
    if nl1:
        _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0, algo=6)
    else:
        _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0), algo=6)
torch.cuda.synchronize()
ws = next(iter(_cabi._workspaces.values()))
off = 256 + 4 * 3 * 1024
n_cta = 148
v = ws[off:off + n_cta * 16 * 8].view(torch.int64).view(n_cta, 16).cpu()
names = ['pk:wait_tmem', 'pk:wait_ring', 'pk:total', 'gr:wait_fin', 'gr:total', 'ga:wait_own', 'ga:poll', 'ga:fin', 'ga:kl',
         'ga:latency', 'ga:units', 'pk:decode', 'pk:raise_refs', 'pk:accumulate', 'pk:tmem_st', 'pk:close']
names = names[:11]
t = v[:, 11:16].double()
t0 = t[:, 0].min()
for i, nme in enumerate(['T:cta_start', 'T:first_data', 'T:park_end', 'T:grad_end', 'T:cta_end']):
    col = (t[:, i] - t0) / 1e3
    print(f'{nme:14s} min {col.min():8.2f}  mean {col.mean():8.2f}  max {col.max():8.2f} us after the first CTA started')
units = v[:, 10].float().clamp(min=1)
for i, nme in enumerate(names):
    col = v[:, i].float()
    per = (col / units).mean() if nme.startswith('ga:') else float('nan')
    print(f'{nme:14s} mean {col.mean():10.0f}  min {col.min():10.0f}  max {col.max():10.0f} cycles   per gathered unit {per:8.0f}')
