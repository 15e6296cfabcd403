#!/usr/bin/env python
"""Warp-stall samples of an ncu report aggregated per CUDA SOURCE LINE (needs -lineinfo in the build).

The report's SASS page gives samples per instruction; `nvdisasm -g` on the cubin of the same build gives
the source line of every instruction; both list the kernel's instructions in address order.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'segdistill_b200', 'libsegdistill_sm100.so')


def sass_rows(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Address"') or not lines[i].strip()),
               len(lines))
    return list(csv.DictReader(io.StringIO('\n'.join(lines[start:end]))))


def kernel_name(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[2][rows[0].index('Kernel Name')]


def line_table():
    """{mangled function: [(offset, file, line)]} from nvdisasm -g over every cubin of the library."""
    out = {}
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(['cuobjdump', '-xelf', 'all', LIB], cwd=d, capture_output=True)
        for cubin in glob.glob(os.path.join(d, '*.cubin')):
            txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
            fn, cur = None, (None, 0)
            for l in txt.splitlines():
                m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
                if m:
                    fn = m.group(1)
                    out.setdefault(fn, [])
                    continue
                m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
                if m:
                    cur = (m.group(1), int(m.group(2)))
                    continue
                m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);', l)
                if m and fn:
                    out[fn].append((int(m.group(1), 16), cur[0], cur[1], m.group(2).strip()))
    return out


def demangle(name):
    return subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = sass_rows(rep)
    kname = kernel_name(rep)
    table = line_table()
    key = re.sub(r'\s+', '', kname.replace('void ', '').split('(')[0])
    cands = [fn for fn in table if re.sub(r'\s+', '', demangle(fn).replace('void ', '').replace('sd::', '').split('(')[0]) == key.replace('sd::', '')]
    cands = [fn for fn in cands if len(table[fn]) == len(rows)] or cands
    if not cands:
        # template arguments print differently in ncu and c++filt (bools, dependent types): same base name, same length
        base = key.replace('sd::', '').split('<')[0]
        cands = [fn for fn in table if base in fn and len(table[fn]) == len(rows)]
    if not cands:
        print('no cubin function matches', kname)
        return
    fn = cands[0]
    insts = table[fn]
    if len(insts) != len(rows):
        print(f'warning: {len(insts)} instructions in the cubin vs {len(rows)} in the report (stale build?)')
    stall_cols = [c for c in rows[0] if c.startswith('stall_') and 'Not Issued' not in c]
    per_line = defaultdict(lambda: [0, 0, Counter()])
    tot = 0
    for r, ins in zip(rows, insts):
        n = int(r['# Samples'] or 0)
        tot += n
        e = per_line[(os.path.basename(ins[1] or '?'), ins[2])]
        e[0] += n
        e[1] += int(r['Instructions Executed'] or 0)
        for c in stall_cols:
            e[2][c[6:]] += int(r[c] or 0)
    src_cache = {}

    def src(f, ln):
        if f not in src_cache:
            p = os.path.join(ROOT, 'segdistill_b200', 'csrc', f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        L = src_cache[f]
        return L[ln - 1].strip() if 0 < ln <= len(L) else ''

    print(f'{kname}\n{tot} samples; per source line: samples %, warp-instructions executed, top stalls, source')
    for (f, ln), (n, ex, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        sts = ' '.join(f'{k}:{v * 100 // max(n, 1)}' for k, v in st.most_common(3))
        print(f'{n * 100.0 / max(tot, 1):5.1f}% {ex:>10d}  {f}:{ln:<4d} [{sts:36s}] {src(f, ln)[:90]}')


if __name__ == '__main__':
    main()
