#!/usr/bin/env python
"""Condense an ncu report (.ncu-rep, captured with --set full) into the few numbers DESIGN.md and
bench.py quote: per-launch duration, DRAM bytes, DRAM / L2 / SM throughput, pipe and issue utilisation,
registers, occupancy, and the top warp-stall reasons of the SASS source page.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/rNN_x.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

RAW = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput % of ncu peak'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 (lts) throughput %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots active %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('sm__inst_executed_pipe_tensor.sum', 'tensor instructions'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
    ('sm__cycles_active.avg', 'SM active cycles'),
    ('smsp__inst_executed.sum', 'warp instructions'),
]


def ncu(rep, page):
    return subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout


def raw_page(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, 'raw'))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"kernel: {d.get('Kernel Name', '?')}   [launch id {d.get('ID', '?')}]")
        for key, label in RAW:
            if key in d and d[key] != '':
                print(f'  {label:34s} {d[key]:>16s} {units[hdr.index(key)]}')
        try:
            rd = float(d['dram__bytes_read.sum'].replace(',', ''))
            wr = float(d['dram__bytes_write.sum'].replace(',', ''))
            u = units[hdr.index('dram__bytes_read.sum')]
            print(f'  {"dram traffic (read + write)":34s} {rd + wr:16.3f} {u}')
        except Exception:
            pass


def source_page(rep, top=12):
    lines = ncu(rep, 'source').splitlines()
    try:
        start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    except StopIteration:
        return
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Address"') or not lines[i].strip()),
               len(lines))
    rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:end]))))
    tot_samples = sum(int(r['# Samples'] or 0) for r in rows)
    tot_inst = sum(int(r['Instructions Executed'] or 0) for r in rows)
    ops = Counter()
    for r in rows:
        src = r['Source'].split()
        op = src[1] if src and src[0].startswith('@') and len(src) > 1 else (src[0] if src else '?')
        ops[op.split('.')[0]] += int(r['Instructions Executed'] or 0)
    stall_cols = [c for c in rows[0] if c.startswith('stall_') and 'Not Issued' not in c]
    agg = Counter()
    for r in rows:
        for c in stall_cols:
            agg[c] += int(r[c] or 0)
    print(f'  source page (first launch): {len(rows)} SASS lines, {tot_samples} samples, {tot_inst} warp-instructions')
    print('  opcode mix: ' + ', '.join(f'{k} {v * 100 // max(tot_inst, 1)}%' for k, v in ops.most_common(14)))
    print('  stall reasons: ' + ', '.join(f'{k[6:]} {v * 100 // max(tot_samples, 1)}%' for k, v in agg.most_common(9)))
    proof = [k for k in ops if k.startswith(('UTC', 'LDTM', 'STTM', 'UTMA', 'UBLKCP', 'SYNCS', 'HMMA'))]
    print('  Blackwell/TMA opcodes present: ' + (', '.join(f'{k} x{ops[k]}' for k in sorted(proof)) or 'none'))
    print(f'  hottest SASS lines:')
    for i, r in sorted(enumerate(rows), key=lambda x: -int(x[1]['# Samples'] or 0))[:top]:
        n = int(r['# Samples'] or 0)
        st = max(stall_cols, key=lambda c: int(r[c] or 0))
        print(f'    {n * 100.0 / max(tot_samples, 1):5.1f}%  {st[6:]:16s} {r["Source"].strip()[:80]}')


for rep in sys.argv[1:]:
    print(f'==== {rep}')
    raw_page(rep)
    source_page(rep)
    print()
