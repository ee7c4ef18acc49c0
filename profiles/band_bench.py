#!/usr/bin/env python3
"""Times single-band plans (CUDA events) for a list of STFT sizes on a synthetic track.
usage: python profiles/band_bench.py [seconds] [N[:ratio] ...]
ratio = f_high / f_low of the band (default 4, the main.py crossovers; "d" = dense, up to Nyquist --
the top band of every crossover set; 10 = the 200-2000 Hz band of the bench workload)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import upmix_b200.center_extraction as ce

seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 600
specs = sys.argv[2:] or ["256", "1024", "4096", "8192", "16384", "65536"]
sr = 48000
n = seconds * sr
g = torch.Generator(device="cuda").manual_seed(1)
L = 0.1 * torch.randn(n, device="cuda", generator=g)
R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
out = torch.empty((3, 1, n), dtype=torch.float32, device="cuda")
for spec in specs:
    N = int(spec.split(":")[0])
    ratio = spec.split(":")[1] if ":" in spec else "4"
    f_low = 32.0 * sr / N
    f_high = sr / 2 if ratio == "d" else min(float(ratio) * f_low, sr / 2)
    e = ce.MultiBandExtractorAccu(N, 0.75, ce.make_blackman_harris, f_low, f_high, sr, "raised_cosine",
                                  f_low / 4, f_high / 4)
    plan = ce.plan_for([e])
    for _ in range(2):
        plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    import math
    tf = 50 * math.log2(N) * n / (ms * 1e-3) / 1e12
    print(f"N={spec:>8s}  {ms:8.3f} ms per {seconds} s  -> {seconds / (ms * 1e-3):10.0f} audio-s/s  {tf:6.2f} TFLOP/s nominal")
    plan.release_workspace()
