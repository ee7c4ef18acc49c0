#!/usr/bin/env python3
"""End-to-end timings of the host-buffer pipeline (upmix_process_host_ex): python profiles/host_e2e.py [seconds]
legs: pinned float32 tensors; pageable float32 numpy; float64 strided views of an interleaved array (main.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, contextlib, io
import upmix_b200.center_extraction as ce
sr = 48000
seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
n = seconds * sr
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands([0.0, 200.0, 2000.0], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
rng = np.random.default_rng(1)
L = (0.1 * rng.standard_normal(n)).astype(np.float32)
R = (0.5 * L + 0.05 * rng.standard_normal(n)).astype(np.float32)
wave = np.stack([L, R], axis=1).astype(np.float64)
hl, hr = torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()
def timeit(name, fn, reps=3):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
        _ = float(r[0][n // 2])
        del r                                   # drop the result before the next call (steady state of the allocators)
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:50s} {dt * 1e3:8.1f} ms  {seconds / dt:10.0f} audio-s/s", flush=True)
timeit("pinned float32 tensors", lambda: ce.extract_center_left_right_multi_band_in_memory(hl, hr, sr, ext))
for th in os.environ.get("THREADS", "8").split(","):
  for pin in ("0", "1"):
    os.environ["UPMIX_HOST_THREADS"] = th
    os.environ["UPMIX_NUMPY_PINNED_OUT"] = pin
    print("numpy outputs:", "pinned (cached blocks)" if pin == "1" else "pageable (np.empty)")
    timeit(f"pageable float32 numpy, {th} threads", lambda: ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext))
    timeit(f"float64 interleaved views, {th} threads", lambda: ce.extract_center_left_right_multi_band_in_memory(wave[:, 0], wave[:, 1], sr, ext))
t0 = time.perf_counter(); x = wave[:, 0].astype(np.float32); y = wave[:, 1].astype(np.float32); dt = time.perf_counter() - t0
print(f"numpy astype(float32) of both columns (1 thread): {dt * 1e3:.1f} ms")
