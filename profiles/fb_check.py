#!/usr/bin/env python3
"""Frame-batched dense-band kernel (upmix_fb.cuh) against the oracle and the one-frame kernel, then timings.
usage: python profiles/fb_check.py [seconds_for_timing]"""
import math, os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import upmix_b200.center_extraction as ce
from upmix_b200 import _native
from oracle import upmix_oracle as uo
sr = 48000

def band(N):
    f_low = 32.0 * sr / N
    e = ce.MultiBandExtractorAccu(N, 0.75, ce.make_blackman_harris, f_low, sr / 2, sr, "raised_cosine", f_low / 4, 0.0)
    b = uo.make_band(N, 0.75, uo.blackman_harris, f_low, sr / 2, sr, "raised_cosine", f_low / 4, 0.0)
    return e, b

def check(N, n):
    e, b = band(N)
    L, R = uo.synth_stereo(n, N, stress=True)
    ref = uo.process_band_batched(b, L.astype(np.float64), R.astype(np.float64))
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    res = {}
    for name, flags in (("fb", 0), ("one", _native.PLAN_NO_BATCH)):
        plan = ce.plan_for([e], _native.OUT_LSCRS, flags)
        out = [o.cpu().numpy() for o in plan.process(dl, dr)]
        res[name] = out
        rep = " ".join(f"{nm}:{uo.snr_db(a, o):6.1f}dB/{np.max(np.abs(a - o)):.1e}" for nm, a, o in zip("CLR", ref, out))
        print(f"N={N:5d} n={n:7d} {name:4s} {rep}", flush=True)
    d = [float(np.max(np.abs(a - o))) for a, o in zip(res["fb"], res["one"])]
    w = [int(np.argmax(np.abs(a - o))) for a, o in zip(res["fb"], res["one"])]
    print(f"          fb vs one-frame max diff {d} at {w}", flush=True)

def timing(N, seconds):
    e, _ = band(N)
    n = seconds * sr
    g = torch.Generator(device="cuda").manual_seed(1)
    L = 0.1 * torch.randn(n, device="cuda", generator=g)
    R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
    out = torch.zeros((3, 1, n), dtype=torch.float32, device="cuda")
    row = []
    for name, flags in (("fb", 0), ("one", _native.PLAN_NO_BATCH)):
        plan = ce.plan_for([e], _native.OUT_LSCRS, flags)
        best = 1e9
        for rep in range(2):
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
        row.append(f"{name} {best * 3600 / seconds:7.3f} ms/band-hour ({50 * math.log2(N) * n / (best * 1e-3) / 1e12:5.1f} TF nominal)")
    print(f"N={N:5d} " + " | ".join(row), flush=True)

if __name__ == "__main__":
    seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    for N in (256, 512, 1024):
        for n in (20 * N + 1237, 5 * N + 3, 100001):
            try:
                check(N, n)
            except Exception:
                traceback.print_exc()
    for N in (256, 512, 1024):
        try:
            timing(N, seconds)
        except Exception:
            traceback.print_exc()
