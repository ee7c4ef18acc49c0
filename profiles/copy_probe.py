#!/usr/bin/env python3
"""Bare pinned H2D + D2H of one 1-hour track's bytes (1.38 GB up, 2.07 GB down) by copy size and by the number of streams per
direction: what the per-copy set-up costs and whether a second stream hides it.  python profiles/copy_probe.py"""
import time, torch
n = 3600 * 48000
dev = torch.device("cuda")
for piece in (1 << 28, 4320 * 1024, 1 << 20):
    for ns in (1, 2, 3):
        hi, ho = torch.empty(piece, pin_memory=True), torch.empty(piece, pin_memory=True)
        di, do = torch.empty(piece, device=dev), torch.empty(piece, device=dev)
        up = [torch.cuda.Stream() for _ in range(ns)]
        dn = [torch.cuda.Stream() for _ in range(ns)]
        def copies():
            for k in range(-(-2 * n // piece)):
                with torch.cuda.stream(up[k % ns]):
                    di.copy_(hi, non_blocking=True)
            for k in range(-(-3 * n // piece)):
                with torch.cuda.stream(dn[k % ns]):
                    ho.copy_(do, non_blocking=True)
            torch.cuda.synchronize()
        copies()
        t0 = time.perf_counter()
        for _ in range(4):
            copies()
        dt = (time.perf_counter() - t0) / 4
        moved = (-(-2 * n // piece) + -(-3 * n // piece)) * piece * 4
        print(f"pieces of {piece * 4 / 1e6:8.1f} MB, {ns} stream(s) per direction: {dt * 1e3 * (20 * n) / moved:7.2f} ms (scaled to the exact bytes)", flush=True)
        del hi, ho, di, do
