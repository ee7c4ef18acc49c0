#!/usr/bin/env python3
"""End-to-end time per rank under torchrun (every rank its own 1-hour track, pinned host buffers), for a few
segment policies of the host pipeline.  usage: torchrun --nproc-per-node N profiles/e2e_multi.py"""
import contextlib
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import upmix_b200.center_extraction as ce

rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
sr, n = 48000, 3600 * 48000
with contextlib.redirect_stdout(io.StringIO()):
    ext = ce.chain_bands([0.0, 200.0, 2000.0], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
plan = ce.plan_for(ext)
g = torch.Generator().manual_seed(1 + rank)
L = (0.1 * torch.randn(n, generator=g)).pin_memory()
R = (0.5 * L + 0.05 * torch.randn(n, generator=g)).pin_memory()
for seg in [float(x) for x in (sys.argv[1:] or [0, 90, 0, 90])]:
    plan.process_host_tensors(L, R, segment_seconds=seg, sample_rate=sr)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        plan.process_host_tensors(L, R, segment_seconds=seg, sample_rate=sr)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / 3], device="cuda", dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"segment_seconds={seg:5.1f}: {dt.item() * 1e3:7.2f} ms (max over {dist.get_world_size()} ranks)", flush=True)
dist.destroy_process_group()
