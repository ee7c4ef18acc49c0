#!/usr/bin/env python3
"""One dense single-band plan over a short track (for ncu): python profiles/fb_one.py N seconds"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upmix_b200.center_extraction as ce
N, seconds = int(sys.argv[1]), int(sys.argv[2])
sr = 48000
f_low = 32.0 * sr / N
e = ce.MultiBandExtractorAccu(N, 0.75, ce.make_blackman_harris, f_low, sr / 2, sr, "raised_cosine", f_low / 4, 0.0)
n = seconds * sr
g = torch.Generator(device="cuda").manual_seed(1)
L = 0.1 * torch.randn(n, device="cuda", generator=g)
R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
out = torch.empty((3, 1, n), dtype=torch.float32, device="cuda")
plan = ce.plan_for([e])
for _ in range(3):
    plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
torch.cuda.synchronize()
