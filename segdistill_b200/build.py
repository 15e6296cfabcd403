"""Build libsegdistill_sm100.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU).

    python -m segdistill_b200.build [--verbose] [--force]

Every ``csrc/*.cu`` is compiled to its own object under ``build/`` (in parallel, only when the source or a header
changed) and the objects are linked into the shared library.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libsegdistill_sm100.so')
OBJ_DIR = os.path.join(os.path.dirname(PKG), 'build', 'obj')

NVCC_FLAGS = [
    '-std=c++17', '-O3', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
]
LINK_FLAGS = ['--shared', '-cudart', 'static', '-gencode', 'arch=compute_100a,code=sm_100a']


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the CUDA extension cannot be built')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def headers():
    deps = glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh'))
    deps.append(os.path.join(os.path.dirname(PKG), 'include', 'segdistill.h'))
    return deps


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in sources() + headers())


def _obj_path(src, extra):
    tag = hashlib.sha1(' '.join(NVCC_FLAGS + extra).encode()).hexdigest()[:8]
    return os.path.join(OBJ_DIR, f'{os.path.splitext(os.path.basename(src))[0]}.{tag}.o')


def _compile(nvcc, src, obj, extra, verbose):
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, res.returncode, res.stdout, ' '.join(cmd)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = find_nvcc()
    extra = os.environ.get('SD_NVCC_EXTRA', '').split()      # e.g. -DSD_CLUSTER_TIMING (scripts/cluster_timing.py)
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in headers())
    jobs, objs = [], []
    for src in sources():
        obj = _obj_path(src, extra)
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time)
        if stale:
            jobs.append((src, obj))
    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for src, rc, out, cmd in pool.map(lambda j: _compile(nvcc, j[0], j[1], extra, verbose), jobs):
            if verbose:
                print(cmd, flush=True)
            if verbose or rc != 0:
                print(out)
            failed = failed or rc != 0
    if failed:
        raise RuntimeError('nvcc failed building libsegdistill_sm100.so')
    cmd = [nvcc] + LINK_FLAGS + ['-o', LIB] + objs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(' '.join(cmd))
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed linking libsegdistill_sm100.so')
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--verbose', action='store_true')
    ap.add_argument('--force', action='store_true')
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
    sys.exit(0)
