#!/usr/bin/env python
"""BASELINE config 2 (CGD g=10 on the four MiT-B0 stage maps, B=16): GPU time of every stage on its own per kernel
variant (CUDA-graph replay of the C-ABI call, so the host is out of the picture), and of the four back to back."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from segdistill_b200 import _cabi  # noqa: E402

dev = torch.device('cuda', 0)
stages = [(16, 32, 128, 128), (16, 64, 64, 64), (16, 160, 32, 32), (16, 256, 16, 16)]
pairs = [(torch.randn(sh, device=dev), torch.randn(sh, device=dev)) for sh in stages]


def timed(fn, n=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (s, t), sh in zip(pairs, stages):
    for algo in ('auto', 'tma', 'stream', 'cluster'):
        try:
            us = timed(lambda: _cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0, algo=_cabi.ALGOS[algo]))
            print(f'{sh} {algo:8s}: {us:7.1f} us  {12 * s.numel() / us / 1e3:7.1f} GB/s  [{_cabi.last_kernel()}]', flush=True)
        except Exception as e:
            print(f'{sh} {algo:8s}: {type(e).__name__} {str(e)[:60]}', flush=True)
            torch.cuda.synchronize()
print(f'all four, one stream: {timed(lambda: [_cabi.kl_rows(s, t, group=10, tau=2.0, alpha=3.0) for s, t in pairs]):7.1f} us')
