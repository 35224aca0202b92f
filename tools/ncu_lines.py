#!/usr/bin/env python3
"""Joins an ncu SASS source page with nvdisasm line info and prints the hottest source lines.

  python tools/ncu_lines.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [top]

(There is no GPU here: the report comes back from gpurun; the .so is the one that ran.)
"""
import csv
import re
import subprocess
import sys
import tempfile
import os
import collections


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # locate the function's text section
    lines = sass.splitlines()
    start = None
    for i, l in enumerate(lines):
        if l.strip().startswith(".section") and ".text." in l and kern in l:
            start = i
            break
    assert start is not None, "kernel not found in SASS"
    cur = ("?", 0)
    inst_lines = []   # (addr, file, line, text)
    stack = []
    for l in lines[start + 1:]:
        s = l.strip()
        if s.startswith(".section"):
            break
        m = re.match(r'//## File "([^"]+)", line (\d+)(.*)', s)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            # inlined-at chains: keep the innermost location (first of the chain)
            continue
        m = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
        if m:
            inst_lines.append((int(m.group(1), 16), cur[0], cur[1], m.group(2)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # several kernels may be in the report; take the first block whose name matches
    blocks, cur_rows = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur_rows = []
            blocks.append((r[1], cur_rows))
        elif cur_rows is not None:
            cur_rows.append(r)
    # pick the report block whose instruction count matches the disassembly (or env NCU_BLOCK)
    bi = int(os.environ.get("NCU_BLOCK", "-1"))
    want = os.environ.get("NCU_NAME")          # substring of the demangled kernel name, e.g. "k_bpnn_mma<(int)0"
    if want:
        cand = [k for k, (nm, rr) in enumerate(blocks) if want in nm]
        exact = [k for k in cand if len(blocks[k][1]) - 1 == len(inst_lines)]
        bi = (exact or cand or [0])[0]
    elif bi < 0:
        bi = 0
        for k, (nm, rr) in enumerate(blocks):
            if len(rr) - 1 == len(inst_lines):
                bi = k; break
    name, rs = blocks[bi]
    hdr = rs[0]
    ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = rs[1:]
    if len(data) != len(inst_lines):
        print("warning: %d profiled instructions vs %d disassembled" % (len(data), len(inst_lines)))
    agg = collections.defaultdict(lambda: [0, 0, 0])
    n = min(len(data), len(inst_lines))
    tot_i = tot_s = 0
    for k in range(n):
        try:
            ni, ns = int(data[k][ci]), int(data[k][si])
        except ValueError:
            continue
        key = (inst_lines[k][1], inst_lines[k][2])
        agg[key][0] += ni
        agg[key][1] += ns
        agg[key][2] += 1
        tot_i += ni
        tot_s += ns
    print("kernel:", name[:100])
    print("warp instructions executed: %d, stall samples: %d" % (tot_i, tot_s))
    srcs = {}
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
        print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100.0 * v[0] / tot_i, 100.0 * v[1] / max(tot_s, 1), f, ln, text))


if __name__ == "__main__":
    main()
