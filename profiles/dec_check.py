#!/usr/bin/env python3
"""Decimated path (upmix_dec.cu) against the oracle and against the full-size kernels, then timings.
usage: python profiles/dec_check.py [seconds_for_timing]"""
import math
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import upmix_b200.center_extraction as ce
from upmix_b200 import _native
from oracle import upmix_oracle as uo

sr = 48000


def band(N, ratio, f_low=None):
    f_low = 32.0 * sr / N if f_low is None else f_low
    f_high = min(ratio * f_low, sr / 2)
    e = ce.MultiBandExtractorAccu(N, 0.75, ce.make_blackman_harris, f_low, f_high, sr, "raised_cosine", f_low / 4, f_high / 4)
    b = uo.make_band(N, 0.75, uo.blackman_harris, f_low, f_high, sr, "raised_cosine", f_low / 4, f_high / 4)
    return e, b


def check(N, ratio, f_low=None):
    e, b = band(N, ratio, f_low)
    n = max(6 * N + 1237, 30011)
    L, R = uo.synth_stereo(n, N, stress=True)
    ref = uo.process_band_batched(b, L.astype(np.float64), R.astype(np.float64))
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    top = int(np.nonzero(b.gain)[0].max())
    res = {}
    for name, flags in (("dec", 0), ("full", _native.PLAN_NO_DECIMATE)):
        plan = ce.plan_for([e], _native.OUT_LSCRS, flags)
        out = [o.cpu().numpy() for o in plan.process(dl, dr)]
        res[name] = out
        rep = " ".join(f"{nm}:{uo.snr_db(a, o):6.1f}dB/{np.max(np.abs(a - o)):.1e}" for nm, a, o in zip("CLR", ref, out))
        print(f"N={N:6d} ratio={ratio:5} top_bin={top:4d} {name:5s} {rep}", flush=True)
    d = max(float(np.max(np.abs(a - o))) for a, o in zip(res["dec"], res["full"]))
    print(f"          dec vs full max diff {d:.2e}", flush=True)


def timing(N, ratio, seconds, f_low=None):
    e, _ = band(N, ratio, f_low)
    n = seconds * sr
    g = torch.Generator(device="cuda").manual_seed(1)
    L = 0.1 * torch.randn(n, device="cuda", generator=g)
    R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
    out = torch.empty((3, 1, n), dtype=torch.float32, device="cuda")
    row = []
    for name, flags in (("dec", 0), ("full", _native.PLAN_NO_DECIMATE)):
        plan = ce.plan_for([e], _native.OUT_LSCRS, flags)
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
        row.append(f"{name} {ms * 3600 / seconds:7.3f} ms/band-hour ({50 * math.log2(N) * n / (ms * 1e-3) / 1e12:5.1f} TF nominal)")
        plan.release_workspace()
    print(f"N={N:6d} ratio={ratio:5} " + " | ".join(row), flush=True)


if __name__ == "__main__":
    seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    cases = [(8192, 10), (8192, 4), (4096, 4), (2048, 2), (65536, 10), (65536, 4), (16384, 4), (32768, 2), (8192, 1.5, 20.0)]
    for c in cases:
        try:
            check(*c)
        except Exception:
            traceback.print_exc()
    for c in cases:
        try:
            timing(c[0], c[1], seconds, *c[2:])
        except Exception:
            traceback.print_exc()
