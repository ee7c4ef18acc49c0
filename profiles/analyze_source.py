#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: stall-reason totals, per-opcode shares and the hottest
SASS sites per kernel.   usage: ncu -i X.ncu-rep --page source --csv > src.csv; analyze_source.py src.csv [filter] [ntop]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
filt = sys.argv[2] if len(sys.argv) > 2 else ""
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 12
kern, data, hdr = None, {}, None
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]
        data[kern] = []
    elif r and r[0] == "Address":
        hdr = r
    elif kern and len(r) > 10:
        data[kern].append(r)
for k, rs in data.items():
    if filt not in k:
        continue
    si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot = sum(int(r[si]) for r in rs)
    totex = sum(int(r[ei]) for r in rs)
    print("=====", k[:70], "| SASS instrs", len(rs), "| samples", tot, "| warp instr executed", totex)
    st = {n: sum(int(r[hdr.index(n)]) for r in rs) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
    print("  stalls:", ", ".join(f"{n[6:]} {100 * v / tot:.1f}%" for n, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
    op, opex = collections.Counter(), collections.Counter()
    for r in rs:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += int(r[si])
        opex[o] += int(r[ei])
    print("  opcode samples%/executed%:", ", ".join(f"{o} {100 * v / tot:.1f}/{100 * opex[o] / totex:.1f}" for o, v in op.most_common(14)))
    top = sorted(range(len(rs)), key=lambda i: -int(rs[i][si]))[:ntop]
    names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for t in top:
        best = max(names, key=lambda n: int(rs[t][hdr.index(n)]))
        j = t
        while j > 0 and not re.search(r"LDG|LDS|BAR|MUFU", rs[j - 1][1]) and t - j < 60:
            j -= 1
        prev = rs[j - 1][1].strip()[:50] if j > 0 else ""
        print(f"  {t:5d} samples={int(rs[t][si]):5d} ({100 * int(rs[t][si]) / tot:4.1f}%) exec={int(rs[t][ei]):8d} {best[6:]:10s} {rs[t][1].strip()[:44]:44s} <- {prev} (-{t - j + 1})")
