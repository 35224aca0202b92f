#!/usr/bin/env python3
"""Static SASS instruction count per source line for one kernel:
   python tools/sass_lines.py <lib.so> <mangled-kernel-substring> [top]"""
import re, subprocess, sys, tempfile, os, collections
so, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = sass.splitlines()
start = None
for i, l in enumerate(lines):
    if l.strip().startswith(".section") and ".text." in l and kern in l:
        start = i; break
assert start is not None
cur = ("?", 0); cnt = collections.Counter(); tot = 0
for l in lines[start + 1:]:
    s = l.strip()
    if s.startswith(".section"): break
    m = re.match(r'//## File "([^"]+)", line (\d+)', s)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'/\*[0-9a-f]{4,}\*/', s): cnt[cur] += 1; tot += 1
print("total", tot)
for (f, ln), c in cnt.most_common(top): print("%5d  %s:%d" % (c, f, ln))
