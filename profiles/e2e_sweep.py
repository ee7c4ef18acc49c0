#!/usr/bin/env python3
"""End-to-end (pinned host in, pinned host out) time of the bench workload against the segment length of
the pipelined host path, with the raw PCIe copy times beside it.  usage: python profiles/e2e_sweep.py [seconds]"""
import contextlib
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import upmix_b200.center_extraction as ce

seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
sr = 48000
n = seconds * sr
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands([0.0, 200.0, 2000.0], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
plan = ce.plan_for(ext)
g = torch.Generator().manual_seed(1)
L = (0.1 * torch.randn(n, generator=g)).pin_memory()
R = (0.5 * L + 0.05 * torch.randn(n, generator=g)).pin_memory()
d = torch.empty((3, n), dtype=torch.float32, device="cuda")
h = torch.empty((3, n), dtype=torch.float32).pin_memory()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def h2d():
    d[0].copy_(L, non_blocking=True)
    d[1].copy_(R, non_blocking=True)


def d2h():
    h.copy_(d, non_blocking=True)


s2 = torch.cuda.Stream()


def both():
    h2d()
    with torch.cuda.stream(s2):
        h.copy_(d, non_blocking=True)


print(f"H2D {2 * n * 4 / 1e9:.2f} GB: {timed(h2d):.2f} ms; D2H {3 * n * 4 / 1e9:.2f} GB: {timed(d2h):.2f} ms; both at once: {timed(both):.2f} ms")
for seg in [float(x) for x in (sys.argv[2:] or [0, 450, 225, 120, 60, 30, 15])]:
    ms = timed(lambda: plan.process_host_tensors(L, R, segment_seconds=seg, sample_rate=sr))
    print(f"segment_seconds={seg:6.1f}: {ms:7.2f} ms end to end -> {seconds / ms * 1e3:9.0f} audio-s/s")
