#!/usr/bin/env python3
"""Wall time per hardware block of the Bela-equivalent streaming mode (upmix_stream_block): python profiles/stream_latency.py [hw]"""
import os, sys, time, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upmix_b200 import bela
hw = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
up = bela.MultiBandUpmix()
with contextlib.redirect_stdout(io.StringIO()):
    up.setup(hw, 48000.0, 4, [0.0, 500.0, 2000.0, 8000.0, 24000.0])
g = torch.Generator(device="cuda").manual_seed(2)
x = 0.1 * torch.randn(2, hw * 400, device="cuda", generator=g)
for i in range(50):
    up.process(x[0, i * hw:(i + 1) * hw], x[1, i * hw:(i + 1) * hw])
torch.cuda.synchronize()
# device time per block (events), blocks queued back to back
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(50, 350):
    up.process(x[0, i * hw:(i + 1) * hw], x[1, i * hw:(i + 1) * hw])
b.record()
torch.cuda.synchronize()
print(f"hw={hw}: {a.elapsed_time(b) / 300 * 1e3:.1f} us per block, back to back (device timeline)")
# wall latency of one block: submit, wait for the result
t = []
for i in range(350, 400):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o = up.process(x[0, i * hw:(i + 1) * hw], x[1, i * hw:(i + 1) * hw])
    torch.cuda.synchronize()
    t.append(time.perf_counter() - t0)
t.sort()
print(f"hw={hw}: wall per block incl. python + sync: median {t[len(t) // 2] * 1e6:.1f} us, min {t[0] * 1e6:.1f} us; block = {hw / 48000 * 1e3:.1f} ms of audio")
