#!/usr/bin/env python3
"""Static SASS instruction mix of a kernel, split at BAR.SYNC (one row per barrier-delimited phase).
usage: sass_phases.py <lib.so> <kernel-substring>"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if pat not in name:
        continue
    ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", l).split(";")[0].strip() for l in blk.split("\n") if re.match(r"\s*/\*[0-9a-f]{4}\*/", l)]
    print("==", name, len(ops), "instructions")
    seg, cur = [], []
    for o in ops:
        cur.append(o)
        if "BAR.SYNC" in o:
            seg.append(cur)
            cur = []
    seg.append(cur)
    tot = collections.Counter()
    for i, sg in enumerate(seg):
        c = collections.Counter()
        for o in sg:
            c[re.sub(r"^@!?U?P\d+\s+", "", o).split()[0].split(".")[0]] += 1
        tot.update(c)
        fp = c["FADD"] + c["FMUL"] + c["FFMA"]
        mem = {k: c[k] for k in ("LDS", "STS", "LDG", "STG", "LDL", "STL") if c[k]}
        other = len(sg) - fp - sum(mem.values())
        print(f"  phase {i:2d}: {len(sg):5d}  FP {fp:4d}  mem {mem}  other {other:4d}  top-other {[(k, v) for k, v in c.most_common(12) if k not in ('FADD','FMUL','FFMA','LDS','STS','LDG','STG')][:6]}")
    print("  total:", dict(tot.most_common(14)))
