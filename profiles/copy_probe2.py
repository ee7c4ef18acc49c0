#!/usr/bin/env python3
"""Bare copies of one 1-hour track's bytes in segment-sized pieces: three 1-D device-to-host copies per segment (one per output
channel) against ONE 2-D copy (3 rows) per segment; uploads as two 1-D copies per segment.  python profiles/copy_probe2.py"""
import ctypes, time, torch
n = 3600 * 48000
dev = torch.device("cuda")
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
hin = torch.empty((2, n), pin_memory=True)
hout = torch.empty((3, n), pin_memory=True)
din = torch.empty((2, n), device=dev)
dout = torch.empty((3, n), device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for seg in (4320 * 1024, 2 * 4320 * 1024, 4 * 4320 * 1024):
    for mode in ("1-D x3", "2-D"):
        def copies():
            for a in range(0, n, seg):
                b = min(n, a + seg)
                with torch.cuda.stream(s1):
                    din[0, a:b].copy_(hin[0, a:b], non_blocking=True)
                    din[1, a:b].copy_(hin[1, a:b], non_blocking=True)
                if mode == "2-D":
                    rc = rt.cudaMemcpy2DAsync(hout[0, a:].data_ptr(), n * 4, dout[0, a:].data_ptr(), n * 4, (b - a) * 4, 3, 2, s2.cuda_stream)
                    assert rc == 0, rc
                else:
                    with torch.cuda.stream(s2):
                        for ch in range(3):
                            hout[ch, a:b].copy_(dout[ch, a:b], non_blocking=True)
            torch.cuda.synchronize()
        copies()
        t0 = time.perf_counter()
        for _ in range(4):
            copies()
        print(f"segments of {seg * 4 / 1e6:6.1f} MB per channel, D2H {mode:7s}: {(time.perf_counter() - t0) / 4 * 1e3:7.2f} ms", flush=True)
