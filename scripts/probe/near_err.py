"""Near-converged accuracy of the row kernels vs the float64 closed form (relative error of the loss), per kernel."""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
import oracle  # noqa: E402
from segdistill_b200 import _cabi  # noqa: E402

dev = torch.device('cuda', 0)


def near(shape, seed, eps=1e-2, offset=0.0):
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(shape, generator=g)
    return s, s + eps * torch.randn(shape, generator=g) + offset


for shape, g, tau, off in (((2, 150, 64, 64), 1, 1.0, 0.0), ((2, 150, 64, 64), 10, 2.0, 0.0), ((2, 150, 128, 128), 10, 2.0, 0.0),
                           ((2, 150, 128, 128), 1, 1.0, 0.3), ((4, 64, 16, 16), 1, 1.0, -2.0)):
    errs = []
    for seed in range(3):
        s, t = near(shape, 100 + seed, offset=off)
        f64 = oracle.kld_closed_form_f64(s.numpy(), t.numpy(), 'channel', g, tau, 3.0)[0]
        sd_, td_ = s.to(dev), t.to(dev)
        row = []
        for algo in ('tma', 'stream', 'cluster', 'generic'):
            try:
                l = _cabi.kl_rows(sd_, td_, group=g, tau=tau, alpha=3.0, algo=_cabi.ALGOS[algo])[0].item()
                row.append(abs(l - f64) / f64)
            except _cabi.SegDistillUnsupported:
                row.append(float('nan'))
        if g == 10 or g == 1:
            try:
                out, _ = _cabi.kl_rows_multi(sd_, td_, (10, 1), (2.0, 1.0), (3.0, 1.0))
                fa = oracle.kld_closed_form_f64(s.numpy(), t.numpy(), 'channel', 10, 2.0, 3.0)[0]
                fb = oracle.kld_closed_form_f64(s.numpy(), t.numpy(), 'channel', 1, 1.0, 1.0)[0]
                row += [abs(out[0].item() - fa) / fa, abs(out[1].item() - fb) / fb]
            except _cabi.SegDistillUnsupported:
                row += [float('nan')] * 2
        errs.append(row)
    e = np.array(errs)
    print(shape, 'g', g, 'tau', tau, 'off', off, 'KL %.2e' % f64, ' max rel err [tma stream cluster generic | pair cgd, pair cd]:',
          ' '.join('%.1e' % v for v in np.nanmax(e, 0)))
