#!/usr/bin/env python
"""IFVD similarity term, bf16 vs fp32, on a power-of-two plane (128x128) and next to it (120x136, 136x136): is the slow
bf16 class-sum kernel a matter of the 32 KB channel stride?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from segdistill_b200 import _cabi  # noqa: E402

dev = torch.device('cuda', 0)
for hw in ((128, 128), (120, 136), (136, 136), (128, 256)):
    for dtype in (torch.float32, torch.bfloat16):
        shape = (16, 150) + hw
        s = torch.randn(shape, device=dev).to(dtype)
        t = torch.randn(shape, device=dev).to(dtype)
        cls = torch.randint(0, 150, (16, 1, hw[0] // 8, hw[1] // 8), device=dev).repeat_interleave(8, 2).repeat_interleave(8, 3)
        cls = cls.reshape(16, -1).to(torch.int32)
        for _ in range(3):
            _cabi.ifvd_sim(s, t, cls)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            _cabi.ifvd_sim(s, t, cls)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        print(f'{hw} {str(dtype):15s} {us:8.1f} us   {us / (s.numel() / 1e6):6.2f} us per Melem', flush=True)
