"""Host cost of the eager module-API step against PyTorch's own floor for the same graph shape:
(a) the benchmarked step (dispatcher forward of CD + CGD fused, sum, backward);
(b) the same with the library's C call and launches replaced by nothing (an autograd.Function that returns two
    preallocated scalars and a preallocated gradient) - what torch.autograd itself costs for one custom node with two
    outputs, an add and a backward;
(c) two torch.empty + one ctypes call of 20 arguments + one tiny kernel launch (scale_grad).
Host time per step (perf_counter over 300 steps, GPU drained afterwards)."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import segdistill_b200 as sd
from segdistill_b200 import _cabi
dev = torch.device('cuda', 0)
S = torch.randn(16, 150, 128, 128, device=dev).requires_grad_(True)
T = torch.randn(16, 150, 128, 128, device=dev)
gt = torch.zeros(16, 1, 128, 128, dtype=torch.long, device=dev)
dl = sd.DistillationLoss([
    {'student_layer': 'a', 'teacher_layer': 'a', 'loss_name': 'CGDLoss', 'loss_config': dict(group_size=10, alpha=3, tau=2)},
    {'student_layer': 'b', 'teacher_layer': 'b', 'loss_name': 'CDLoss', 'loss_config': {}}])
fS, fT = {'a': S, 'b': S}, {'a': T, 'b': T}

def step_ours():
    S.grad = None
    l1, l2 = dl(fS, fT, gt, 1, None, None).values()
    (l1 + l2).backward()

out2 = torch.zeros(2, device=dev)
dsbuf = torch.zeros_like(S)

class Floor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t):
        return out2[0], out2[1]
    @staticmethod
    def backward(ctx, g0, g1):
        return dsbuf, None

def step_floor():
    S.grad = None
    l1, l2 = Floor.apply(S, T)
    (l1 + l2).backward()

def timeit(fn, n=300):
    for _ in range(30): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6

print(f'(a) library step, host: {timeit(step_ours):7.1f} us')
print(f'(b) torch.autograd floor (one custom node, two outputs, add, backward): {timeit(step_floor):7.1f} us')
def calls():
    a = torch.empty(S.shape, dtype=S.dtype, device=dev); b = torch.empty(2, dtype=torch.float32, device=dev)
    _cabi.scale_grad_(dsbuf, out2[0])
print(f'(c) two torch.empty + one C call with a launch: {timeit(calls):7.1f} us')
