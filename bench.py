#!/usr/bin/env python3
"""Benchmark of the multi-band STFT centre-extraction path (BASELINE.json metric: realtime factor,
audio-seconds processed per second, 48 kHz stereo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- a 1-hour 48 kHz stereo track, 3 bands
(crossovers 0/200/2000 Hz -> STFT sizes 65536, 8192, 1024 by the dynamic-resolution rule), Ls/C/Rs
out.  One step = one pass of the whole track through every band, the bands summed in band order.  With N > 1 every
rank runs its own 1-hour track (independent shards, no collective; "scaling": "weak") and `value`
is the total audio-seconds of all ranks over the slowest rank's device time.

`value`: inputs resident in HBM, CUDA events on the launching stream.  `e2e`: the same metric through
the public call `extract_center_left_right_multi_band_in_memory` with pinned HOST tensors (H2D and D2H
inside the timed region).  `roofline`: this path is FP32-bound (north_star) --
algorithmic flops = 10*T*sum(log2 N_b) per stereo sample (T = 5 real transforms per frame) over
the measured FMA throughput of this GPU; the HBM view (20 B per stereo sample over the measured copy
bandwidth in MEASURED_PEAKS.json) is reported beside it.  `cpu_baseline`: the oracle port of the
reference prototype (frame loop, one thread per band) on a bounded excerpt, on this box's host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000
EDGES = [0.0, 200.0, 2000.0]
TRACK_SECONDS = 3600
T_REAL_FFTS = 5
BYTES_PER_SAMPLE = 20           # 2 x 4 B in, 3 x 4 B out
FP32_FALLBACK_TFLOPS = 74.4     # 148 SM x 128 lanes x 2 x 1.965 GHz (SURVEY.md 8d planning figure)
HBM_FALLBACK_GBS = 6650.0       # B200_PROFILING.md fallback


def quiet_chain(ce, **kw):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return ce.chain_bands(EDGES, 0.75, ce.make_blackman_harris, SR, "raised_cosine", **kw)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """Start of the timed region: only samples taken after this count."""
        self.t_mark = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if len(r) < 6 or ts < getattr(self, "t_mark", 0.0):
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(sample_seconds=240.0):
    """Oracle port of the reference prototype (frame loop + ThreadPoolExecutor over bands,
    center_extraction.py:449-460, 499-501) on the first `sample_seconds` of the workload."""
    from oracle import upmix_oracle as uo
    n = int(sample_seconds * SR)
    L, R = uo.synth_stereo(n, 1)
    L64, R64 = L.astype(np.float64), R.astype(np.float64)
    bands = uo.chain(EDGES, 0.75, uo.blackman_harris, SR)
    t0 = time.perf_counter()
    uo.upmix_multiband(bands, L64, R64, batched=False, threads=True)
    dt = time.perf_counter() - t0
    return {"value": sample_seconds / dt, "unit": "audio-s/s", "cores": min(len(bands), os.cpu_count() or 1),
            "kind": "port", "sample": f"first {int(sample_seconds)} s of the 1-hour track, frame-loop oracle, "
            f"one thread per band ({len(bands)} bands), numpy {np.__version__}, {os.cpu_count()} host cpus",
            "seconds": dt}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (oracle port: the Python prototype cannot
    travel to the GPU box) on a bounded sample per step."""
    if rank != 0:
        return
    sample = 60.0
    vals = []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(sample)
        if i >= args.warmup:
            vals.append(b)
    dt = float(np.mean([v["seconds"] for v in vals]))
    value = sample / dt
    cb = dict(vals[-1], value=value)
    cb.pop("seconds", None)
    line = {"impl": "reference", "metric": "realtime factor (audio-s/s, 48 kHz stereo)", "value": value,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, sample_note=f"each step = {int(sample)} s excerpt on host cores"),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, sample_note=None):
    cfg = {"workload": "cfg2: 1-hour 48 kHz stereo track per GPU, 3 bands (crossovers 0/200/2000 Hz; STFT 65536/8192/1024, "
                       "75% overlap Blackman-Harris WOLA), Ls/C/Rs out",
           "sample_rate": SR, "track_seconds": TRACK_SECONDS, "bands": 3, "stft_sizes": [65536, 8192, 1024],
           "tracks_per_gpu": 1, "sharding": f"independent tracks x{n_gpus}, no collective",
           "l2": "inputs (1.38 GB) and outputs (2.07 GB) per step exceed the 126 MB L2"}
    if sample_note:
        cfg["note"] = sample_note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=int, default=TRACK_SECONDS, help="track length (default: the 1-hour workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200): upmix_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"      # NCCL's version banner goes to stdout; keep it to the JSON line
        dist.init_process_group("nccl", device_id=dev)

    import upmix_b200.center_extraction as ce
    from upmix_b200 import _native

    n = args.seconds * SR
    ext = quiet_chain(ce)
    sizes = [e.block_size for e in ext]
    plan = ce.plan_for(ext)

    # synthetic track of this rank, generated on the device (SURVEY.md 8d formula)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    L = 0.1 * torch.randn(n, device=dev, generator=g)
    R = 0.5 * L + 0.05 * torch.randn(n, device=dev, generator=g)
    out = torch.empty((3, 1, n), dtype=torch.float32, device=dev)

    def step():
        plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()                     # nvidia-smi needs a few hundred ms to come up: start before the warm-up
    for _ in range(args.warmup):
        step()
    barrier()
    time.sleep(0.3)
    sampler.mark()
    _native.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _native.launch_count()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * args.seconds / (ms_per_step * 1e-3)

    # ---- per-band timing (single-band plans) for the dominant-kernel roofline -------------------
    band_ms = []
    if rank == 0:
        for e in ext:
            p1 = ce.plan_for([e])
            o1 = out
            for _ in range(2):
                p1.process_segment(L[None], R[None], 0, n, 0, n, out=o1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a.record()
            for _ in range(3):
                p1.process_segment(L[None], R[None], 0, n, 0, n, out=o1)
            b.record()
            torch.cuda.synchronize(dev)
            band_ms.append(a.elapsed_time(b) / 3)
            p1.release_workspace()

    # ---- end to end through the public API with pinned host tensors -----------------------------
    e2e = None
    if not args.no_e2e:
        hl = torch.empty(n, dtype=torch.float32, pin_memory=True)
        hr = torch.empty(n, dtype=torch.float32, pin_memory=True)
        hl.copy_(L)
        hr.copy_(R)
        ce.extract_center_left_right_multi_band_in_memory(hl, hr, SR, ext)      # warm-up (allocations)
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 3))
        for _ in range(reps):
            res = ce.extract_center_left_right_multi_band_in_memory(hl, hr, SR, ext)
            _ = float(res[0][n // 2])                                          # touch the result on the host
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * args.seconds / float(tt.item()), "unit": "audio-s/s", "h2d_bytes_per_step": 8 * n,
               "d2h_bytes_per_step": 12 * n, "ms_per_step": float(tt.item()) * 1e3,
               "api": "upmix_b200.center_extraction.extract_center_left_right_multi_band_in_memory(pinned CPU tensors)"}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        try:
            fp32_tflops, sms = _native.measure_fp32_tflops(local_rank)
            fp32_src = "measured in this run (FMA probe kernel)"
        except Exception as ex:  # pragma: no cover
            fp32_tflops, sms, fp32_src = FP32_FALLBACK_TFLOPS, 148, f"fallback ({ex})"
        samples_per_s = value / world * SR               # per GPU
        w_all = 10 * T_REAL_FFTS * sum(math.log2(s) for s in sizes)
        # dominant kernel: the fused kernel of the slowest band that is served by a single kernel
        # (one launch per step); the four-step band (three kernels per wave) is listed beside it
        fused = [i for i, s_ in enumerate(sizes) if s_ <= 8192]
        dom = max(fused, key=lambda i: band_ms[i]) if (band_ms and fused) else 0
        dom_n = sizes[dom]
        dom_flops = 10 * T_REAL_FFTS * math.log2(dom_n) * n
        dom_tflops = dom_flops / (band_ms[dom] * 1e-3) / 1e12 if band_ms else None
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tr.get(f"band_fused_kernel<{dom_n}>")
            if ent:      # DRAM bytes per stereo sample of this kernel from the committed ncu --set full capture
                traffic = ent["dram_bytes_per_sample"] * n
        except Exception:
            pass
        roofline = {"bound": "fp32", "kernel": f"band_fused_kernel<{dom_n}> (timed as the single-band plan of the N={dom_n} band: "
                                               "one launch of this kernel per step, writing its hops straight to the outputs)",
                    "achieved": dom_tflops, "peak": fp32_tflops, "unit": "TFLOP/s",
                    "frac": (dom_tflops / fp32_tflops) if dom_tflops else None, "traffic": traffic,
                    "traffic_note": "dram__bytes_read+write per launch scaled from profiles/traffic.json (ncu --set full)",
                    "peak_source": fp32_src, "sm_count": sms,
                    "algorithmic": f"10*T*log2(N) = {10 * T_REAL_FFTS * math.log2(dom_n):.0f} flops per stereo sample for this band "
                                   f"(T={T_REAL_FFTS} real FFTs/frame, 75% overlap), {n} samples per launch; 20 B/sample compulsory HBM bytes",
                    "band_ms": dict(zip([str(s) for s in sizes], band_ms)),
                    "whole_path": {"flops_per_sample": w_all, "achieved_tflops": samples_per_s * w_all / 1e12,
                                   "frac": samples_per_s * w_all / 1e12 / fp32_tflops},
                    "hbm_view": {"bound": "hbm", "achieved": samples_per_s * BYTES_PER_SAMPLE / 1e9,
                                 "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": samples_per_s * BYTES_PER_SAMPLE / 1e9 / peaks.get("hbm_gbs", HBM_FALLBACK_GBS),
                                 "peak_source": peak_src + " (MEASURED_PEAKS.json)" if peak_src == "measured" else "fallback"}}
        line = {"metric": "realtime factor (audio-s/s, 48 kHz stereo)", "value": value, "unit": "audio-s/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world), "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
                "e2e": e2e}
        if args.seconds != TRACK_SECONDS:
            line["config"]["note"] = f"REDUCED track length {args.seconds} s (not the headline workload)"
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline()
            cb.pop("seconds", None)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
