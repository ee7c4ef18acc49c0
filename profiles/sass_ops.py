#!/usr/bin/env python3
"""Per-kernel SASS opcode summary of the built library (evidence that the kernels are Blackwell-native and of what they
spend their instructions on): python profiles/sass_ops.py [lib.so] > profiles/r02_sass_ops.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "upmix_b200", "csrc", "libupmix_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
cols = ["total", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "SYNCS", "BAR", "LDL", "STL"]
print(f"# {os.path.basename(lib)}: arch {arch}; static instruction counts per kernel (cuobjdump -sass)")
print("# FFMA2/FMUL2/FADD2 = packed FP32x2 (sm_100), LDGSTS = cp.async, UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier, LDL/STL = spills")
print(" ".join(f"{c:>7s}" for c in cols) + "  kernel")
tot = collections.Counter()
for blk in re.split(r"\n\s*Function : ", out)[1:]:
    name = blk.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(upmix::BandDev.*", "", dem).replace("upmix::", "").replace("(int)", "").replace("(bool)", "")
    c = collections.Counter()
    for l in blk.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            c[m.group(1)] += 1
            c["total"] += 1
    tot.update(c)
    print(" ".join(f"{c[k]:7d}" for k in cols) + "  " + dem[:110])
print(" ".join(f"{tot[k]:7d}" for k in cols) + "  ALL KERNELS")
