#!/usr/bin/env python3
"""Run-length view of a kernel's dynamic instruction counts (ncu source page joined with nvdisasm line info):
   python tools/ncu_runs.py <report.ncu-rep> <lib.so> <mangled-substring> <demangled-substring> [minshare]
consecutive SASS instructions with the same execution count form one run: count, instructions, share, source lines."""
import csv, re, subprocess, sys, tempfile, os, collections
rep, so, kern, want = sys.argv[1:5]
minshare = float(sys.argv[5]) if len(sys.argv) > 5 else 0.3
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = sass.splitlines()
start = None
for i, l in enumerate(lines):
    if l.strip().startswith(".section") and ".text." in l and kern in l:
        start = i; break
cur = ("?", 0); inst = []
for l in lines[start + 1:]:
    s = l.strip()
    if s.startswith(".section"): break
    m = re.match(r'//## File "([^"]+)", line (\d+)(.*)', s)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
    if m: inst.append((cur[0], cur[1], m.group(2)))
src = rep if rep.endswith(".csv") else None
out = open(src).read() if src else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; c = None
for r in rows:
    if r and r[0] == "Kernel Name": c = []; blocks.append((r[1], c))
    elif c is not None: c.append(r)
name, rs = [b for b in blocks if want in b[0]][0]
hdr = rs[0]; ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(int(r[ci]), int(r[si])) for r in rs[1:] if r[ci].isdigit()]
n = min(len(data), len(inst)); tot = sum(d[0] for d in data)
print(name[:100], "instructions", len(data), "disassembled", len(inst), "total executed", tot)
i = 0
while i < n:
    j = i
    while j + 1 < n and abs(data[j + 1][0] - data[i][0]) <= 0.02 * max(data[i][0], 1): j += 1
    cnt = sum(data[k][0] for k in range(i, j + 1)); smp = sum(data[k][1] for k in range(i, j + 1))
    if 100.0 * cnt / tot >= minshare:
        f64 = sum(1 for k in range(i, j + 1) if re.match(r'(@!?U?P\d+\s+)?D(FMA|MUL|ADD|SETP)', inst[k][2]))
        srcs = collections.Counter((inst[k][0], inst[k][1]) for k in range(i, j + 1))
        top = ", ".join("%s:%d(%d)" % (a, b, v) for (a, b), v in srcs.most_common(5))
        print("%5d-%5d exec %9d x %3d inst (fp64 %3d) = %5.2f%%  samples %5.2f%%  %s" % (i, j, data[i][0], j - i + 1, f64, 100.0 * cnt / tot, 100.0 * smp / max(1, sum(d[1] for d in data)), top))
    i = j + 1
