#!/usr/bin/env python3
"""One of the BASELINE config shapes, two passes, for an ncu launch list.  usage: python profiles/cfg_driver.py cfg1|cfg3|cfg4 [seconds]"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import upmix_b200.center_extraction as ce
from upmix_b200 import _native

which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 600
sr, edges, mb, mode = {"cfg1": (48000, [0, 30, 120, 480, 1920, 7680], 65536, _native.OUT_LSCRS),
                       "cfg3": (48000, [0, 500, 2000, 8000], 8192, _native.OUT_FOLD),
                       "cfg4": (96000, [0, 100, 200, 400, 800, 1600, 3200, 6400], 8192, _native.OUT_LSCRS)}[which]
n = seconds * sr
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands(edges, 0.75, ce.make_blackman_harris, sr, "raised_cosine" if which != "cfg3" else "hard_zero", max_block_size=mb)
plan = ce.plan_for(ext, mode)
g = torch.Generator(device="cuda").manual_seed(1)
L = 0.1 * torch.randn(n, device="cuda", generator=g)
R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
for _ in range(2):
    plan.process(L, R)
torch.cuda.synchronize()
print("done", [e.block_size for e in ext])
