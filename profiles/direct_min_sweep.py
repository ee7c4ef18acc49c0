#!/usr/bin/env python3
"""Staged vs direct band summation against the track length (bench workload bands).  Run twice:
UPMIX_DIRECT_MIN=1 (always direct) and UPMIX_DIRECT_MIN=1000000000000 (always staged)."""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import upmix_b200.center_extraction as ce

sr = 48000
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands([0.0, 200.0, 2000.0], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
plan = ce.plan_for(ext)
res = []
for seconds in (5, 10, 20, 40, 80, 160, 320):
    n = seconds * sr
    L = 0.1 * torch.randn(n, device="cuda")
    R = 0.5 * L + 0.05 * torch.randn(n, device="cuda")
    out = torch.empty((3, 1, n), dtype=torch.float32, device="cuda")
    for _ in range(3):
        plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
    b.record()
    torch.cuda.synchronize()
    res.append(f"{seconds}s {a.elapsed_time(b) / 10:.3f}")
    plan.release_workspace()
print(os.environ.get("UPMIX_DIRECT_MIN"), " | ".join(res))
