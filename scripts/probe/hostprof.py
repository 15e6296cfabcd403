import cProfile, pstats, sys, time, torch
sys.path.insert(0, '/root/repo')
import segdistill_b200 as sd
dev = torch.device('cuda', 0)
S = torch.randn(16, 150, 128, 128, device=dev).requires_grad_(True)
T = torch.randn(16, 150, 128, 128, device=dev)
gt = torch.zeros(16, 1, 128, 128, dtype=torch.long, device=dev)
dl = sd.DistillationLoss([
    {'student_layer': 'a', 'teacher_layer': 'a', 'loss_name': 'CGDLoss', 'loss_config': dict(group_size=10, alpha=3, tau=2)},
    {'student_layer': 'b', 'teacher_layer': 'b', 'loss_name': 'CDLoss', 'loss_config': {}}])
fS, fT = {'a': S, 'b': S}, {'a': T, 'b': T}
def step():
    S.grad = None
    out = dl(fS, fT, gt, 1, None, None)
    l1, l2 = out.values()
    (l1 + l2).backward()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('host us/step', (t1 - t0) / 200 * 1e6, 'with drain', (t2 - t0) / 200 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
