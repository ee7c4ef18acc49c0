#!/usr/bin/env python3
"""Device-resident timing of every BASELINE.json config shape (CUDA events, best of 3 after warm-up).
usage: python profiles/config_bench.py [fp32_tflops]"""
import contextlib
import io
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import upmix_b200.center_extraction as ce
from upmix_b200 import _native, bela


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def synth(tracks, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    L = 0.1 * torch.randn((tracks, n), device="cuda", generator=g)
    R = 0.5 * L + 0.05 * torch.randn((tracks, n), device="cuda", generator=g)
    return L, R


def timeit(plan, L, R, reps=3):
    n_out = 3 if plan.out_mode == _native.OUT_LSCRS else 2
    out = torch.empty((n_out, L.shape[0], L.shape[1]), dtype=torch.float32, device="cuda")
    for _ in range(2):
        plan.process_segment(L, R, 0, L.shape[1], 0, L.shape[1], out=out)
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        plan.process_segment(L, R, 0, L.shape[1], 0, L.shape[1], out=out)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    del out
    return best


peak = float(sys.argv[1]) if len(sys.argv) > 1 else _native.measure_fp32_tflops(0)[0]
print(f"FP32 peak used: {peak:.1f} TFLOP/s")
rows = []


def report(name, ext, plan, tracks, seconds, sr, T):
    n = int(seconds * sr)
    L, R = synth(tracks, n, 1)
    ms = timeit(plan, L, R)
    w = 10 * T * sum(math.log2(e.block_size) for e in ext)
    audio = tracks * seconds
    rtf = audio / (ms * 1e-3)
    tf = rtf * sr * w / 1e12
    print(f"{name:34s} sizes {[e.block_size for e in ext]}  {tracks:4d} x {seconds:6.0f} s  {ms:9.3f} ms  {rtf:11.0f} audio-s/s  "
          f"W={w:5.0f} flop/sample  {tf:6.2f} TFLOP/s = {100 * tf / peak:5.1f} % of FP32 peak")
    plan.release_workspace()
    del L, R
    torch.cuda.empty_cache()


e1 = quiet(ce.chain_bands, [0, 30, 120, 480, 1920, 7680], 0.75, ce.make_blackman_harris, 48000, "raised_cosine")
report("cfg1 default 6 bands, 10 s", e1, ce.plan_for(e1), 1, 10, 48000, 5)
report("cfg1 shape, 1 hour", e1, ce.plan_for(e1), 1, 3600, 48000, 5)
e2 = quiet(ce.chain_bands, [0, 200, 2000], 0.75, ce.make_blackman_harris, 48000, "raised_cosine")
report("cfg2 3 bands, 1 hour", e2, ce.plan_for(e2), 1, 3600, 48000, 5)
e3 = quiet(bela.bela_chain_bands, [0.0, 500.0, 2000.0, 8000.0, 24000.0], 48000.0, 2048)
report("cfg3 Bela bands, fold-down, 600 s", e3, ce.plan_for(e3, _native.OUT_FOLD), 1, 600, 48000, 5)
e4 = quiet(ce.chain_bands, [0, 100, 200, 400, 800, 1600, 3200, 6400], 0.75, ce.make_blackman_harris, 96000, "raised_cosine",
           max_block_size=8192)
report("cfg4 8 bands 96 kHz, 600 s", e4, ce.plan_for(e4), 1, 600, 96000, 5)
report("cfg5 shape: 32 x 5-min tracks, 6 bands", e1, ce.plan_for(e1), 32, 300, 48000, 5)

# Bela streaming latency: wall time per 2048-sample hardware block
up = bela.MultiBandUpmix()
quiet(up.setup, 2048, 48000.0, 4, [0.0, 500.0, 2000.0, 8000.0, 24000.0])
L, R = synth(1, 2048 * 96, 2)
import time
for i in range(32):                      # (start-up blocks, then one graph capture per block phase)
    up.process(L[0, i * 2048:(i + 1) * 2048], R[0, i * 2048:(i + 1) * 2048])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(32, 96):
    up.process(L[0, i * 2048:(i + 1) * 2048], R[0, i * 2048:(i + 1) * 2048])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 64
print(f"cfg3 streaming: {dt * 1e6:.0f} us wall per 2048-sample block (42.7 ms of audio; algorithmic latency 3*2048 samples = 128 ms)")
