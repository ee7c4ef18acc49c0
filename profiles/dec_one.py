#!/usr/bin/env python3
"""One single-band plan over a short track (for ncu launch lists): python profiles/dec_one.py N ratio seconds [full]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upmix_b200.center_extraction as ce
from upmix_b200 import _native
N, ratio, seconds = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
flags = _native.PLAN_NO_DECIMATE if len(sys.argv) > 4 and sys.argv[4] == "full" else 0
sr = 48000
f_low = 32.0 * sr / N
f_high = min(ratio * f_low, sr / 2)
e = ce.MultiBandExtractorAccu(N, 0.75, ce.make_blackman_harris, f_low, f_high, sr, "raised_cosine", f_low / 4, f_high / 4)
n = seconds * sr
g = torch.Generator(device="cuda").manual_seed(1)
L = 0.1 * torch.randn(n, device="cuda", generator=g)
R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
out = torch.empty((3, 1, n), dtype=torch.float32, device="cuda")
plan = ce.plan_for([e], _native.OUT_LSCRS, flags)
for _ in range(3):
    plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
torch.cuda.synchronize()
