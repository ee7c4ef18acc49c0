import contextlib, io, os, sys
sys.path.insert(0, os.getcwd())
import torch
import upmix_b200.center_extraction as ce
sr = 48000; n = 1200 * sr
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands([0, 30, 120, 480, 1920, 7680], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
plan = ce.plan_for(ext)
g = torch.Generator(device="cuda").manual_seed(1)
L = 0.1 * torch.randn(n, device="cuda", generator=g); R = 0.5 * L + 0.05 * torch.randn(n, device="cuda", generator=g)
out = torch.empty((3, 1, n), dtype=torch.float32, device="cuda")
for _ in range(2): plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
torch.cuda.synchronize()
