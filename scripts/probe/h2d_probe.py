"""H2D bandwidth of the e2e step's two input copies (2 x 157 MB pinned): one stream vs two streams vs chunked."""
import torch, time
dev = torch.device('cuda', 0)
shape = (16, 150, 128, 128)
hS = torch.randn(shape).pin_memory(); hT = torch.randn(shape).pin_memory()
dS = torch.empty(shape, device=dev); dT = torch.empty(shape, device=dev)
nbytes = 2 * hS.numel() * 4
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def one():
    dS.copy_(hS, non_blocking=True); dT.copy_(hT, non_blocking=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): dS.copy_(hS, non_blocking=True)
    with torch.cuda.stream(s2): dT.copy_(hT, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
for name, fn in (('one stream', one), ('two streams', two)):
    ms = t(fn)
    print(f'{name:12s} {ms:.3f} ms  {nbytes / ms / 1e6:.1f} GB/s')
