"""Build libsegdistill_sm100.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU).

    python -m segdistill_b200.build [--verbose] [--force]
"""
from __future__ import annotations

import argparse
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libsegdistill_sm100.so')

NVCC_FLAGS = [
    '-std=c++17', '-O3', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
    '--shared', '-cudart', 'static',
]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the CUDA extension cannot be built')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh'))
    deps.append(os.path.join(os.path.dirname(PKG), 'include', 'segdistill.h'))
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    extra = os.environ.get('SD_NVCC_EXTRA', '').split()      # e.g. -DSD_CLUSTER_TIMING (scripts/cluster_timing.py)
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + sources()
    if verbose:
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libsegdistill_sm100.so')
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--verbose', action='store_true')
    ap.add_argument('--force', action='store_true')
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
    sys.exit(0)
