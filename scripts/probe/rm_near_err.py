#!/usr/bin/env python
"""Loss error vs float64 of the bf16 register-capacity row kernels on nearly converged pairs (KL ~ 4e-5):
run with SEGDISTILL_ROWS_RM=0 / 1 to compare kl_rows_tma_kernel and kl_rows_rm_kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import segdistill_b200 as sd  # noqa: E402
from segdistill_b200 import _cabi  # noqa: E402

dev = torch.device('cuda', 0)
for shape, g, tau in [((3, 128, 16, 16), 64, 2.0), ((2, 32, 32, 32), 16, 1.0), ((2, 8, 64, 64), 4, 2.0),
                      ((3, 150, 128, 128), 1, 1.0)]:
    errs = []
    for seed in range(6):
        for offset in (0.0, 0.5):
            gen = torch.Generator().manual_seed(seed)
            s = torch.randn(shape, generator=gen)
            t = (s + 1e-2 * torch.randn(shape, generator=gen) + offset).to(torch.bfloat16)
            s = s.to(torch.bfloat16)
            f64_loss, f64_grad, _ = oracle.kld_closed_form_f64(s.float().numpy(), t.float().numpy(), 'channel', g, tau, 3.0)
            x = s.to(dev).requires_grad_(True)
            loss = sd.CGDLoss(group_size=g, alpha=3, tau=tau)(x, t.to(dev), None, 1)
            loss.backward()
            errs.append(abs(loss.item() - f64_loss) / f64_loss)
    print(f'{_cabi.last_kernel():20s} {str(shape):22s} g={g:3d} KL~{f64_loss:.2e}  rel err: max {max(errs):.2e}  mean {np.mean(errs):.2e}')
