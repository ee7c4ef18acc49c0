#!/usr/bin/env python3
"""Timings of single-band plans, decimated path: python profiles/dec_perf.py [seconds] [N:ratio ...]  (UPMIX_FORCE_ACCUM=1 for the accumulating variants)"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upmix_b200.center_extraction as ce
from upmix_b200 import _native
sr = 48000
seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 600
specs = sys.argv[2:] or ["8192:10", "8192:4", "4096:4", "2048:2", "65536:10", "65536:4", "16384:4"]
n = seconds * sr
g = torch.Generator(device="cuda").manual_seed(1)
L = 0.1 * torch.randn(n, device="cuda", generator=g)
R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
out = torch.zeros((3, 1, n), dtype=torch.float32, device="cuda")
for spec in specs:
    N, ratio = int(spec.split(":")[0]), float(spec.split(":")[1])
    f_low = 32.0 * sr / N
    f_high = min(ratio * f_low, sr / 2)
    e = ce.MultiBandExtractorAccu(N, 0.75, ce.make_blackman_harris, f_low, f_high, sr, "raised_cosine", f_low / 4, f_high / 4)
    plan = ce.plan_for([e], _native.OUT_LSCRS, 0)
    best = 1e9
    for rep in range(3):
        for _ in range(2):
            plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 3)
    print(f"N={N:6d} ratio={ratio:5} {best * 3600 / seconds:7.3f} ms/band-hour ({50 * math.log2(N) * n / (best * 1e-3) / 1e12:5.1f} TF nominal)", flush=True)
    plan.release_workspace()
