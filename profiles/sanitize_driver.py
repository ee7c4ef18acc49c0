#!/usr/bin/env python3
"""Tiny run of every kernel (fused sizes 64..8192, four-step 16384, band sum, streaming, export) for
compute-sanitizer.  usage: compute-sanitizer --tool memcheck python profiles/sanitize_driver.py"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import upmix_b200.center_extraction as ce
from upmix_b200 import _native, bela

sr = 48000
rng = np.random.default_rng(0)
for n_fft in (64, 256, 1024, 2048, 4096, 8192, 16384):
    n = 3 * n_fft + 77
    L = (0.1 * rng.standard_normal(n)).astype(np.float32)
    R = (0.1 * rng.standard_normal(n)).astype(np.float32)
    f_low = 32.0 * sr / n_fft
    e = ce.MultiBandExtractorAccu(n_fft, 0.75, ce.make_blackman_harris, f_low, min(4 * f_low, sr / 2), sr, "raised_cosine",
                                  f_low / 4, f_low)
    out = e.process_all_blocks(L, R)
    assert all(np.isfinite(o).all() for o in out)
with contextlib.redirect_stdout(io.StringIO()):
    up = bela.MultiBandUpmix()
    up.setup(512, 48000.0, 4, [0.0, 500.0, 2000.0, 8000.0, 24000.0])
for i in range(6):
    up.process((0.1 * rng.standard_normal(512)).astype(np.float32), (0.1 * rng.standard_normal(512)).astype(np.float32))
e = ce.MultiBandExtractorAccu(1024, 0.75, ce.make_blackman_harris, 200.0, 2000.0, sr, "raised_cosine", 50.0, 500.0)
for f in range(5):
    e.process_stereo_chunk(L[f * 256:f * 256 + 1024], R[f * 256:f * 256 + 1024])
e.flush_final()
c, l, r = (torch.randn(5001, device="cuda") for _ in range(3))
_native.peak3(c, l, r)
_native.export_mix("split", 0.5, c, l, r)
torch.cuda.synchronize()
print("sanitize driver done")
