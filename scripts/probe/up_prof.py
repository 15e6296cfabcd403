"""one CGD and one PD call behind a 4x resize (training shape) for ncu"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segdistill_b200 import _cabi
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
s = torch.randn((16, 150, 128, 128), device=dev, generator=g)
t = torch.randn((16, 150, 128, 128), device=dev, generator=g)
for _ in range(3):
    _cabi.kl_rows_up(s, t, 4, group=10, tau=2.0, alpha=3.0)
    _cabi.kl_pixels_up(s, t, 4)
torch.cuda.synchronize()
