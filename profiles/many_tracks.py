import contextlib, io, os, sys
sys.path.insert(0, os.getcwd())
import torch
import upmix_b200.center_extraction as ce
sr = 48000
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands([0, 30, 120, 480, 1920, 7680], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
plan = ce.plan_for(ext)
for tracks, secs in [(int(a), int(b)) for a, b in (x.split("x") for x in (sys.argv[1:] or ["32x300", "256x37", "512x19"]))]:
    n = secs * sr
    g = torch.Generator(device="cuda").manual_seed(1)
    L = 0.1 * torch.randn((tracks, n), device="cuda", generator=g); R = 0.5 * L + 0.05 * torch.randn((tracks, n), device="cuda", generator=g)
    out = torch.empty((3, tracks, n), dtype=torch.float32, device="cuda")
    for _ in range(2): plan.process_segment(L, R, 0, n, 0, n, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(3): plan.process_segment(L, R, 0, n, 0, n, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"{tracks} tracks x {secs} s: {ms:.2f} ms -> {tracks * secs / ms * 1e3:.0f} audio-s/s")
    del L, R, out; plan.release_workspace()
