#!/usr/bin/env python3
"""Does running the whole plan over L2-sized time segments (all bands on segment i, then segment i+1) pay?  The band-order
sum is a read-modify-write of the outputs by every band after the first; with segments of a few M samples those
re-reads hit the 126 MB L2 instead of HBM.  python profiles/l2_segments.py [seconds]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import upmix_b200.center_extraction as ce
sr = 48000
seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
n = seconds * sr
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands([0.0, 200.0, 2000.0], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
plan = ce.plan_for(ext)
g = torch.Generator(device="cuda").manual_seed(1)
L = 0.1 * torch.randn(n, device="cuda", generator=g)
R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
out = torch.empty((3, n), device="cuda")
lib, h = plan._lib, plan._h
st = torch.cuda.current_stream().cuda_stream
def run(S):
    wsb = plan.workspace_bytes(min(S, n), 1)
    ws = plan._workspace(wsb)
    for a in range(0, n, S):
        b = min(n, a + S)
        rc = lib.upmix_process_segment(h, L.data_ptr(), R.data_ptr(), 0, n, n, a, b, 1, n, out[0, a:].data_ptr(), out[1, a:].data_ptr(),
                                       out[2, a:].data_ptr(), n, ws.data_ptr(), wsb, st)
        assert rc == 0, rc
ref = None
for S in (n, 64 << 20, 16 << 20, 8 << 20, 4 << 20, 3 << 20, 2 << 20):
    S = min(n, S // 65536 * 65536) if S < n else n
    for _ in range(2):
        run(S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run(S)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    same = bool(torch.equal(ref, out))
    print(f"segment {S / 2**20:8.2f} Mi samples ({S * 20 / 1e6:7.1f} MB of in+out): {e0.elapsed_time(e1) / 5:7.3f} ms per track-hour   bit-identical to one call: {same}", flush=True)
