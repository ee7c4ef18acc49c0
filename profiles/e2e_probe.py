#!/usr/bin/env python3
"""Where the pinned e2e leg loses against the bare-copy ceiling: the host pipeline with and without its kernels, by segment
length, against bare copies in the same granularity.  python profiles/e2e_probe.py"""
import contextlib, io, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    import upmix_b200.center_extraction as ce
    sr, n = 48000, 3600 * 48000
    with contextlib.redirect_stdout(io.StringIO()):
        ext = ce.chain_bands([0.0, 200.0, 2000.0], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
    g = torch.Generator().manual_seed(1)
    hl = (0.1 * torch.randn(n, generator=g)).pin_memory()
    hr = (0.05 * torch.randn(n, generator=g)).pin_memory()
    def run():
        r = ce.extract_center_left_right_multi_band_in_memory(hl, hr, sr, ext)
        return float(r[0][n // 2])
    run(); run()
    t0 = time.perf_counter()
    for _ in range(4):
        run()
    dt_pinned = (time.perf_counter() - t0) / 4
    import numpy as np
    wave = np.empty((n, 2), dtype=np.float64)
    wave[:, 0] = hl.numpy(); wave[:, 1] = hr.numpy()
    def run_np():
        r = ce.extract_center_left_right_multi_band_in_memory(wave[:, 0], wave[:, 1], sr, ext)
        x = float(r[0][n // 2]); del r
        return x
    run_np(); run_np()
    t0 = time.perf_counter()
    for _ in range(3):
        run_np()
    print(f"{sys.argv[1]:50s} pinned {dt_pinned * 1e3:7.2f} ms   numpy float64 views {(time.perf_counter() - t0) / 3 * 1e3:7.2f} ms", flush=True)
    if sys.argv[1] == "bare":
        dev = torch.device("cuda")
        for piece in (1 << 28, 4320 * 1024, 1 << 20):
            hi, ho = torch.empty(piece, pin_memory=True), torch.empty(piece, pin_memory=True)
            di, do = torch.empty(piece, device=dev), torch.empty(piece, device=dev)
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
            def copies():
                with torch.cuda.stream(s1):
                    for _ in range(-(-2 * n // piece)):
                        di.copy_(hi, non_blocking=True)
                with torch.cuda.stream(s2):
                    for _ in range(-(-3 * n // piece)):
                        ho.copy_(do, non_blocking=True)
                s1.synchronize(); s2.synchronize()
            copies()
            t0 = time.perf_counter()
            for _ in range(4):
                copies()
            dt = (time.perf_counter() - t0) / 4
            moved = (-(-2 * n // piece) + -(-3 * n // piece)) * piece * 4
            print(f"bare copies in pieces of {piece * 4 / 1e6:8.1f} MB: {dt * 1e3 * (20 * n) / moved:7.2f} ms (scaled to the exact bytes)")
        # D2H alone / H2D alone
        for name, fn in (("D2H alone", lambda: ho.copy_(do, non_blocking=True)), ("H2D alone", lambda: di.copy_(hi, non_blocking=True))):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(200): fn()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(f"{name}: {200 * piece * 4 / dt / 1e9:.1f} GB/s")
else:
    legs = []
    for rep in range(3):
        legs += [("x2 up to 90 s", {}),
                 ("x1.3 up to 360 s", {"UPMIX_HOST_MAXSEG": str(360 * 48000), "UPMIX_HOST_GROWTH": "1.3"}),
                 ("x1.3 up to 180 s", {"UPMIX_HOST_MAXSEG": str(180 * 48000), "UPMIX_HOST_GROWTH": "1.3"}),
                 ("x1.5 up to 360 s", {"UPMIX_HOST_MAXSEG": str(360 * 48000), "UPMIX_HOST_GROWTH": "1.5"})]
    for label, env in legs:
        subprocess.run([sys.executable, __file__, label], env=dict(os.environ, **env))
