#!/usr/bin/env python3
"""Static SASS opcode histogram of the hot kernels of libfnetgpu.so (cuobjdump -sass):
   python tools/sass_hist.py [lib.so] > profiles/r02_sass_opcode_histogram.txt"""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fortnet_b200", "libfnetgpu.so")
want = [("k_acsf_lean<2,1,STRUCT,unsorted,G=4,f64>  (C2 values)", "_Z11k_acsf_leanILi2ELi1ELi2ELb0ELi4ELb0E"),
        ("k_acsf_lean<2,1,STRUCT,sorted,G=2,f64>  (C3 values)", "_Z11k_acsf_leanILi2ELi1ELi2ELb1ELi2ELb0E"),
        ("k_acsf_lean<1,4,DIRECT,unsorted,G=1,f64>  (C5 values)", "_Z11k_acsf_leanILi1ELi4ELi0ELb0ELi1ELb0E"),
        ("k_acsf_force_lean<2,1,STRUCT,sorted,G=2,local>  (C4 forces)", "_Z17k_acsf_force_leanILi2ELi1ELi2ELb1ELi2ELb1E"),
        ("k_bpnn_mma<0,4,1,fused>  (C2 training gradient)", "_Z10k_bpnn_mmaILi0ELi4ELi1ELi1E"),
        ("k_bpnn_mma<0,9,2,cluster-fused>  (C3 training gradient)", "_Z10k_bpnn_mmaILi0ELi9ELi2ELi2E"),
        ("k_bpnn_mma<1,1,2>  (C4 input gradients)", "_Z10k_bpnn_mmaILi1ELi1ELi2ELi0E")]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
print("static SASS opcode histogram, sm_100a, %s" % os.path.basename(lib))
print("(tcgen05 / TMA / UTMA opcodes: none -- the path is FP64; the tensor-core instruction is DMMA.8x8x4, see DESIGN.md section 4)")
for label, mangled in want:
    body = [f for f in funcs[1:] if f.startswith(mangled)]
    if not body:
        print("\n%s: not found" % label); continue
    ops = collections.Counter()
    for l in body[0].split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_x]+)*)", l)
        if m:
            full = m.group(1)
            op = full.split(".")[0]
            ops["DMMA.8x8x4" if full.startswith("DMMA") else op] += 1
    tot = sum(ops.values())
    print("\n%s: %d instructions" % (label, tot))
    print("  " + "  ".join("%s %d" % (k, v) for k, v in ops.most_common(18)))
    print("  FP64 (DFMA+DMUL+DADD+DSETP) %d, DMMA %d, LDS %d, STS %d, LDG %d, IMAD %d, MUFU %d, BAR %d" % (
        sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP")), ops["DMMA.8x8x4"], ops["LDS"], ops["STS"], ops["LDG"], ops["IMAD"], ops["MUFU"], ops["BAR"]))
