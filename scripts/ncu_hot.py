#!/usr/bin/env python
"""Summarise an ncu report's SASS source page: top stall lines and per-opcode instruction counts.
usage: python scripts/ncu_hot.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Address"') or not lines[i].strip()), len(lines))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:end]))))  # first kernel of the report only
tot_samples = sum(int(r['# Samples'] or 0) for r in rows)
tot_inst = sum(int(r['Instructions Executed'] or 0) for r in rows)
print(f'{len(rows)} SASS lines, {tot_samples} samples, {tot_inst} warp-instructions')
ops = Counter()
for r in rows:
    op = r['Source'].split()[0] if not r['Source'].strip().startswith('@') else r['Source'].split()[1]
    ops[op.split('.')[0]] += int(r['Instructions Executed'] or 0)
print('opcode mix:', ', '.join(f'{k}:{v * 100 // max(tot_inst, 1)}%' for k, v in ops.most_common(18)))
stall_cols = [c for c in rows[0] if c.startswith('stall_') and 'Not Issued' not in c]
agg = Counter()
for r in rows:
    for c in stall_cols:
        agg[c] += int(r[c] or 0)
print('stall totals:', ', '.join(f'{k[6:]}:{v * 100 // max(tot_samples, 1)}%' for k, v in agg.most_common(10)))
print('--- hottest lines (samples, % , instr executed, top stall, SASS)')
for i, r in sorted(enumerate(rows), key=lambda x: -int(x[1]['# Samples'] or 0))[:top]:
    n = int(r['# Samples'] or 0)
    st = max(stall_cols, key=lambda c: int(r[c] or 0))
    print(f'{i:5d} {n:6d} {n * 100.0 / tot_samples:5.1f}% {r["Instructions Executed"]:>8s} {st[6:]:14s} {r["Source"].strip()[:90]}')
