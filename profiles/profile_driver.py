#!/usr/bin/env python3
"""Small driver for ncu captures: the bench workload (3 bands: 65536 / 8192 / 1024) on a shorter
track, a few passes, nothing else.  Usage on the GPU box:
    python profiles/profile_driver.py [seconds] [passes]
"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import upmix_b200.center_extraction as ce

seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sr = 48000
n = seconds * sr
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands([0.0, 200.0, 2000.0], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
plan = ce.plan_for(ext)
g = torch.Generator(device="cuda").manual_seed(1)
L = 0.1 * torch.randn(n, device="cuda", generator=g)
R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
out = torch.empty((3, 1, n), dtype=torch.float32, device="cuda")
for _ in range(passes):
    plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
torch.cuda.synchronize()
print("done", float(out[0, 0, n // 2]))
