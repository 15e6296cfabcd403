"""Fused bilinear resize (kl_rows_up) vs host-side F.interpolate + the ordinary kernels, training shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import segdistill_b200 as sd
from segdistill_b200 import _cabi
dev = torch.device('cuda', 0)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

for shape, hw in (((2, 150, 128, 128), (512, 512)), ((2, 150, 64, 64), (512, 512)), ((16, 150, 128, 128), (512, 512))):
    g = torch.Generator(device=dev).manual_seed(0)
    s = torch.randn(shape, device=dev, generator=g).requires_grad_(True)
    t = torch.randn(shape, device=dev, generator=g)
    gt = torch.zeros(shape[0], 1, *hw, dtype=torch.long, device=dev)
    for cls in (sd.CGDLoss, sd.CDLoss, sd.PDLoss):
        res = {}
        for fuse in (True, False):
            crit = cls(); crit.fuse_resize = fuse
            def f():
                s.grad = None
                crit(s, t, gt, 1).backward()
            res[fuse] = timeit(f)
        hi = shape[0] * shape[1] * hw[0] * hw[1]
        print(f'{cls.__name__:8s} {shape} -> {hw}: fused {res[True]:9.1f} us   host resize + kernels {res[False]:9.1f} us   x{res[False] / res[True]:.2f}   ({hi / res[True] / 1e3:.1f} G up-sampled elem/s)')
