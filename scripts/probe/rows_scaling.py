"""fused CGD+CD kernel time vs batch size: fixed cost (launch, pipeline fill/drain) vs per-row cost."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segdistill_b200 import _cabi
dev = torch.device('cuda', 0)
for B in (1, 2, 4, 8, 16, 32, 48):
    g = torch.Generator(device=dev).manual_seed(0)
    s = torch.randn((B, 150, 128, 128), device=dev, generator=g)
    t = torch.randn((B, 150, 128, 128), device=dev, generator=g)
    f = lambda: _cabi.kl_rows_multi(s, t, (10, 1), (2.0, 1.0), (3.0, 1.0))
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30): f()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 30 * 1e3
    rows = B * 15
    print(f'B={B:3d} super-rows={rows:4d} rows/cluster={-(-rows // 15):3d}  {us:8.1f} us   {12 * s.numel() / us / 1e3:7.0f} GB/s  [{_cabi.last_kernel()}]')
