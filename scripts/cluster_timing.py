#!/usr/bin/env python
"""Where the cluster-resident kernel's time goes: per-CTA cycle counters written by a build with
-DSD_CLUSTER_TIMING (SD_NVCC_EXTRA=-DSD_CLUSTER_TIMING python -m segdistill_b200.build --force).

stats warp : wait records | records -> summaries pushed | wait exchange | merge -> row statistics
park warp 0: all rows (incl. the two waits) | wait for a TMEM slot | wait for the ring | total
grad warp 8: wait row statistics | gradient
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
    _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0))
torch.cuda.synchronize()
ws = next(iter(_cabi._workspaces.values()))
R = shape[0] * shape[1]
off = 256 + 4 * 3 * 1024 + 4 * R * 2
off = (off + 127) // 128 * 128
n_cta = min(120, shape[0] * 15 * 8)
v = ws[off:off + n_cta * 16 * 8].view(torch.int64).view(n_cta, 16).cpu()
names = ['st:wait_rec', 'st:summ', 'st:wait_xch', 'st:merge', 'pk:rows', 'pk:wait_tmem', 'pk:wait_ring', 'pk:total', 'gr:wait_fin', 'gr:grad', 'pk:lds+max', 'pk:redux', 'pk:math', 'pk:chunk_after_lds', 'T:fin_last', 'T:grad_done']
for i, nme in enumerate(names):
    col = v[:, i].float()
    print(f'{nme:12s} mean {col.mean():10.0f}  min {col.min():10.0f}  max {col.max():10.0f} cycles  (per row: {col.mean() / max(1, shape[0] * 15 // 15):8.0f})')
