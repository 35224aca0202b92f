#!/usr/bin/env python3
"""Dynamic instruction profile of one kernel from an ncu source-page CSV, by SASS address blocks:
   python tools/ncu_regions2.py <source.csv> <kernel-substring> [block]
prints per block: executed warp instructions (share), stall samples, FP64 share, first source text."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]; blk = int(sys.argv[3]) if len(sys.argv) > 3 else 64
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; blocks.append((r[1], cur))
    elif cur is not None:
        cur.append(r)
name, rs = [b for b in blocks if want in b[0]][0]
hdr = rs[0]; ci, si, so = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
data = [(int(r[ci]), int(r[si]), r[so].strip()) for r in rs[1:] if r[ci].isdigit()]
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print(name[:90]); print("total", tot, "samples", tots)
for b in range(0, len(data), blk):
    seg = data[b:b + blk]
    n = sum(d[0] for d in seg); s = sum(d[1] for d in seg)
    f64 = sum(d[0] for d in seg if re.match(r'(@!?U?P\d+\s+)?D(FMA|MUL|ADD|SETP)', d[2]))
    ops = collections.Counter()
    for d in seg:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', d[2]); ops[m.group(2) if m else '?'] += d[0]
    print("%5d-%5d  %5.1f%% inst  %5.1f%% samples  fp64 %4.0f%%  maxexec %9d  %s" % (b, b + len(seg) - 1, 100.0 * n / tot, 100.0 * s / max(tots, 1), 100.0 * f64 / max(n, 1), max(d[0] for d in seg), " ".join("%s:%.0f%%" % (k, 100.0 * v / max(n, 1)) for k, v in ops.most_common(4))))
