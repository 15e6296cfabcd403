"""Shared helpers for the test-suite (fixtures loading, seeded inputs)."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_cases(prefix='kld_'):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    rec = {k: z[k] for k in z.files}
    if 'kwargs' in rec:
        rec['kwargs'] = ast.literal_eval(str(rec['kwargs']))
        rec['cls'] = str(rec['cls'])
    return rec


def seeded_pair(shape, seed=0, scale=1.0, dtype=torch.float32):
    """SURVEY.md §8(d) synthetic inputs: manual_seed, S = randn, then T = randn."""
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(shape, generator=g) * scale
    t = torch.randn(shape, generator=g) * scale
    return s.to(dtype), t.to(dtype)


def rel_err(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(b), 1e-30)
