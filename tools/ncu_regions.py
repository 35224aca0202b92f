#!/usr/bin/env python3
"""Aggregates the per-line output of tools/ncu_lines.py into code regions (acsf.cuh / cells.cuh)."""
import re, collections, sys
reg = collections.Counter(); smp = collections.Counter()
def region(f, l, text):
    if 'sm_30_intrinsics' in f: return 'shfl intrinsics'
    t = text
    if f == 'cells.cuh':
        if 'floor_div' in t or 'int q = a / b' in t or 'a % b' in t: return 'neighbor_cell'
        if 'nc.' in t or 'p.w' in t or 'S.D[' in t or 'S.nb' in t: return 'neighbor_cell'
        return 'cells: staging / candidate iteration'
    if f == 'acsf.cuh':
        return 'acsf.cuh'
    return f
for line in open(sys.argv[1]):
    m = re.match(r'\s*([\d.]+)% inst\s+([\d.]+)% samples\s+(\S+):(\d+)\s*(.*)', line)
    if not m: continue
    reg[(m.group(3), int(m.group(4)) // 1)] += float(m.group(1))
    smp[(m.group(3), int(m.group(4)))] += float(m.group(2))
# ranges given on the command line: file:lo-hi=name
rng = []
for a in sys.argv[2:]:
    spec, name = a.split('=')
    f, r = spec.split(':'); lo, hi = r.split('-')
    rng.append((f, int(lo), int(hi), name))
out = collections.Counter(); outs = collections.Counter()
for (f, l), v in reg.items():
    nm = f
    for (rf, lo, hi, name) in rng:
        if rf == f and lo <= l <= hi: nm = name; break
    out[nm] += v; outs[nm] += smp[(f, l)]
for k, v in out.most_common(): print('%6.1f%% inst %6.1f%% samples  %s' % (v, outs[k], k))
