#!/usr/bin/env python3
"""Benchmark of the multi-band STFT centre-extraction path (BASELINE.json metric: realtime factor,
audio-seconds processed per second, 48 kHz stereo).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg5|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (config.workload):
  cfg2 (default; BASELINE.json configs[1], the configuration the metric is quoted on): a 1-hour 48 kHz stereo
       track, 3 bands (crossovers 0/200/2000 Hz -> STFT sizes 65536, 8192, 1024 by the dynamic-resolution rule),
       Ls/C/Rs out.  One step = the whole track through every band, bands summed in band order.  With N > 1 every
       rank runs its own 1-hour track ("scaling": "weak"; independent shards, no collective).
       --split-track: ONE 1-hour track cut into N halo'd time segments, one per rank ("scaling": "strong").
  cfg5 (configs[4]): 512 synthetic 5-minute tracks (seeds 1000..1511), main.py's default six bands, track i on rank
       i mod N (upmix_b200.sharding.tracks_for_rank), processed in waves of tracks per GPU ("scaling": "strong").
  cfg1 (configs[0] shape): 10 s, default six bands.
  cfg3-stream: the Bela program's bands, one 2048-sample hardware block per call (latency of upmix_stream_block).

`value`: inputs resident in HBM, CUDA events on the launching stream, max over ranks.  `e2e`: the same metric through
the public call `extract_center_left_right_multi_band_in_memory` with pinned HOST tensors (H2D and D2H inside the
timed region; the host pipeline is upmix_process_host_ex in the C ABI); `e2e.numpy` is the call exactly as main.py
makes it (main.py:43-50, 78-80: float64 strided views of one interleaved pageable array in, fresh float32 arrays
out); `e2e.copy_ceiling` is a bare pinned H2D + D2H of the same bytes on two streams, timed in the same job.
`roofline`: this path is FP32-bound (north_star) -- algorithmic flops = 10*T*log2(N) per stereo
sample and band (T = 5 real transforms per frame) over the measured FMA throughput of this GPU, for the kernel with
the longest launch; the whole-path and HBM views are reported beside it.  `cpu_baseline`: the oracle port of the
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
T_REAL_FFTS = 5
BYTES_PER_SAMPLE = 20           # 2 x 4 B in, 3 x 4 B out
FP32_FALLBACK_TFLOPS = 74.4     # 148 SM x 128 lanes x 2 x 1.965 GHz (SURVEY.md 8d planning figure)
HBM_FALLBACK_GBS = 6650.0       # B200_PROFILING.md fallback

WORKLOADS = {
    "cfg2": dict(edges=[0.0, 200.0, 2000.0], seconds=3600, tracks=1, per_rank=True,
                 text="cfg2: 1-hour 48 kHz stereo track per GPU, 3 bands (crossovers 0/200/2000 Hz; STFT 65536/8192/1024, "
                      "75% overlap Blackman-Harris WOLA), Ls/C/Rs out"),
    "cfg5": dict(edges=[0.0, 30.0, 120.0, 480.0, 1920.0, 7680.0], seconds=300, tracks=512, per_rank=False,
                 text="cfg5: batch of 512 synthetic 5-minute 48 kHz stereo tracks, main.py's default 6 bands (STFT "
                      "65536/65536/16384/4096/1024/256), track i on rank i mod N, waves of tracks per GPU, Ls/C/Rs out"),
    "cfg1": dict(edges=[0.0, 30.0, 120.0, 480.0, 1920.0, 7680.0], seconds=10, tracks=1, per_rank=True,
                 text="cfg1 shape: 10 s 48 kHz stereo, main.py's default 6 bands, Ls/C/Rs out"),
    "cfg3-stream": dict(edges=[0.0, 500.0, 2000.0, 8000.0], seconds=60, tracks=1, per_rank=True,
                        text="cfg3 block streaming: the Bela program's 4 bands (bela/upmix.cpp:498-514; STFT 8192/4096/1024/256 at "
                             "hwBlock 2048), L+0.5C / R+0.5C out, one hardware block per call with carried state (upmix_stream_block)"),
}


def quiet_chain(ce, edges, **kw):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return ce.chain_bands(list(edges), 0.75, ce.make_blackman_harris, SR, "raised_cosine", **kw)


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


def cpu_baseline(edges, sample_seconds, seed=1):
    """Oracle port of the reference prototype (frame loop + ThreadPoolExecutor over bands,
    center_extraction.py:449-460, 499-501) on the first `sample_seconds` of the workload's first track."""
    from oracle import upmix_oracle as uo
    n = int(sample_seconds * SR)
    L, R = uo.synth_stereo(n, seed)
    L64, R64 = L.astype(np.float64), R.astype(np.float64)
    bands = uo.chain(list(edges), 0.75, uo.blackman_harris, SR)
    t0 = time.perf_counter()
    uo.upmix_multiband(bands, L64, R64, batched=False, threads=True)
    dt = time.perf_counter() - t0
    return {"value": sample_seconds / dt, "unit": "audio-s/s", "cores": min(len(bands), os.cpu_count() or 1),
            "kind": "port", "sample": f"first {int(sample_seconds)} s of a track of the workload, frame-loop oracle, "
            f"one thread per band ({len(bands)} bands), numpy {np.__version__}, {os.cpu_count()} host cpus",
            "seconds": dt}


def workload_config(name, n_gpus, sizes, sample_note=None, split=False):
    w = WORKLOADS[name]
    cfg = {"workload": w["text"], "sample_rate": SR, "track_seconds": w["seconds"], "bands": len(sizes), "stft_sizes": sizes}
    if name == "cfg5":
        cfg.update(tracks_total=w["tracks"], sharding=f"tracks i mod {n_gpus}, no collective",
                   l2="every wave's inputs and outputs exceed the 126 MB L2")
    elif split:
        cfg.update(tracks_total=1, sharding=f"one track cut into {n_gpus} time segments with window-length halos, no collective",
                   l2="inputs and outputs per step exceed the 126 MB L2")
    else:
        cfg.update(tracks_per_gpu=1, sharding=f"independent tracks x{n_gpus}, no collective",
                   l2="inputs (1.38 GB) and outputs (2.07 GB) per step exceed the 126 MB L2" if name == "cfg2"
                   else "REDUCED size: inputs and outputs fit the L2 (not a headline number)")
    if sample_note:
        cfg["note"] = sample_note
    return cfg


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (oracle port: the Python prototype cannot
    travel to the GPU box) on a bounded sample per step."""
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    sample = min(60.0, float(w["seconds"]))
    vals = []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(w["edges"], sample, seed=1000 if args.workload == "cfg5" else 1)
        if i >= args.warmup:
            vals.append(b)
    dt = float(np.mean([v["seconds"] for v in vals]))
    value = sample / dt
    cb = dict(vals[-1], value=value)
    cb.pop("seconds", None)
    from oracle import upmix_oracle as uo
    sizes = [int(b.n_fft) for b in uo.chain(list(w["edges"]), 0.75, uo.blackman_harris, SR)]
    line = {"impl": "reference", "metric": "realtime factor (audio-s/s, 48 kHz stereo)", "value": value,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong" if (args.workload == "cfg5" or args.split_track) else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, args.gpus, sizes, sample_note=f"each step = {int(sample)} s excerpt on host cores",
                                      split=args.split_track),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_stream(args, local_rank):
    """--workload cfg3-stream: latency of one hardware block of the Bela-equivalent streaming mode.  One step = 100 blocks
    queued back to back on one stream (device timeline, CUDA events); `wall_us_per_block` = submit one block and wait."""
    import contextlib
    import io
    import torch
    from upmix_b200 import _native, bela
    torch.cuda.set_device(local_rank)
    hw = args.hw_block
    up = bela.MultiBandUpmix()
    with contextlib.redirect_stdout(io.StringIO()):
        up.setup(hw, float(SR), 4, WORKLOADS["cfg3-stream"]["edges"] + [24000.0])
    per_step = 100
    total = (args.warmup + args.steps) * per_step + 150
    g = torch.Generator(device="cuda").manual_seed(2)
    x = 0.1 * torch.randn(2, hw * total, device="cuda", generator=g)
    i = 0
    for _ in range(args.warmup * per_step):
        up.process(x[0, i * hw:(i + 1) * hw], x[1, i * hw:(i + 1) * hw])
        i += 1
    torch.cuda.synchronize()
    _native.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps * per_step):
        up.process(x[0, i * hw:(i + 1) * hw], x[1, i * hw:(i + 1) * hw])
        i += 1
    e1.record()
    torch.cuda.synchronize()
    launches = _native.launch_count()
    ms_per_step = e0.elapsed_time(e1) / args.steps
    us_block = ms_per_step * 1e3 / per_step
    wall = []
    for _ in range(100):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        up.process(x[0, i * hw:(i + 1) * hw], x[1, i * hw:(i + 1) * hw])
        torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
        i += 1
    wall.sort()
    sizes = [b.block_size for b in up.bands]
    line = {"metric": "realtime factor (audio-s/s, 48 kHz stereo)", "value": hw / SR / (us_block * 1e-6), "unit": "audio-s/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS["cfg3-stream"]["text"], "hw_block": hw, "blocks_per_step": per_step, "stft_sizes": sizes,
                       "note": "latency workload (not the headline): inputs and state stay in L2"},
            "us_per_block": us_block, "wall_us_per_block": {"median": wall[len(wall) // 2] * 1e6, "min": wall[0] * 1e6,
                                                            "what": "python call + one graph launch + synchronize"},
            "block_ms_of_audio": hw / SR * 1e3, "gpu_launches": int(launches)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--split-track", action="store_true", help="cfg2: one track cut into one halo'd time segment per rank")
    ap.add_argument("--seconds", type=int, default=0, help="override the track length (marks the line as REDUCED)")
    ap.add_argument("--tracks", type=int, default=0, help="cfg5: override the number of tracks (marks the line as REDUCED)")
    ap.add_argument("--wave-tracks", type=int, default=32, help="cfg5: tracks per wave on one GPU")
    ap.add_argument("--hw-block", type=int, default=2048, help="cfg3-stream: samples per hardware block")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "cfg3-stream" and args.impl == "ours":
        if rank == 0:
            run_stream(args, local_rank)
        return
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
        dist.init_process_group("nccl", device_id=dev)

    import upmix_b200.center_extraction as ce
    from upmix_b200 import _native, sharding

    W = WORKLOADS[args.workload]
    seconds = args.seconds or W["seconds"]
    n = seconds * SR
    ext = quiet_chain(ce, W["edges"])
    sizes = [e.block_size for e in ext]
    plan = ce.plan_for(ext)
    reduced = seconds != W["seconds"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- the step of each workload ---------------------------------------------------------------------------------
    e2e_fn = None
    if args.workload == "cfg5":
        n_tracks_total = args.tracks or W["tracks"]
        reduced = reduced or n_tracks_total != W["tracks"]
        mine = list(sharding.tracks_for_rank(n_tracks_total, rank, world))
        Ls = torch.empty((len(mine), n), dtype=torch.float32, device=dev)
        Rs = torch.empty((len(mine), n), dtype=torch.float32, device=dev)
        for i, t in enumerate(mine):                                   # track t: seed 1000 + t
            g = torch.Generator(device=dev).manual_seed(1000 + t)
            Ls[i] = 0.1 * torch.randn(n, device=dev, generator=g)
            Rs[i] = 0.5 * Ls[i] + 0.05 * torch.randn(n, device=dev, generator=g)
        wt = max(1, min(args.wave_tracks, len(mine)))
        out = torch.empty((3, wt, n), dtype=torch.float32, device=dev)   # one wave of outputs, reused wave after wave
        audio_seconds_total = n_tracks_total * seconds

        def step():
            for t0 in range(0, len(mine), wt):
                nt = min(wt, len(mine) - t0)
                plan.process_segment(Ls[t0:t0 + nt], Rs[t0:t0 + nt], 0, n, 0, n, out=out[:, :nt])

        h2d = 8 * n * len(mine)
        d2h = 12 * n * len(mine)
        ewt = max(1, min(8, len(mine)))

        def make_e2e():
            hin = torch.empty((2, ewt, n), dtype=torch.float32, pin_memory=True)
            hout = torch.empty((3, ewt, n), dtype=torch.float32, pin_memory=True)
            hin[0].copy_(Ls[:ewt])
            hin[1].copy_(Rs[:ewt])
            din = [torch.empty((2, ewt, n), dtype=torch.float32, device=dev) for _ in range(2)]
            dout = [torch.empty((3, ewt, n), dtype=torch.float32, device=dev) for _ in range(2)]
            s_up, s_down, main = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)

            def run():
                # every wave of this rank's tracks: host -> device, all bands, device -> host (double-buffered; the host
                # buffers of one wave are reused as source and destination of every wave: the bytes moved are real)
                free_in = [None, None]
                free_out = [None, None]
                for wi, t0 in enumerate(range(0, len(mine), ewt)):
                    nt = min(ewt, len(mine) - t0)
                    k = wi & 1
                    with torch.cuda.stream(s_up):
                        if free_in[k] is not None:
                            s_up.wait_event(free_in[k])
                        din[k][:, :nt].copy_(hin[:, :nt], non_blocking=True)
                        up = torch.cuda.Event()
                        up.record(s_up)
                    main.wait_event(up)
                    if free_out[k] is not None:
                        main.wait_event(free_out[k])
                    plan.process_segment(din[k][0, :nt], din[k][1, :nt], 0, n, 0, n, out=dout[k][:, :nt])
                    done = torch.cuda.Event()
                    done.record(main)
                    free_in[k] = done
                    with torch.cuda.stream(s_down):
                        s_down.wait_event(done)
                        hout[:, :nt].copy_(dout[k][:, :nt], non_blocking=True)
                        fo = torch.cuda.Event()
                        fo.record(s_down)
                    free_out[k] = fo
                main.wait_stream(s_down)
                torch.cuda.synchronize(dev)
                return float(hout[0, 0, n // 2])
            return run
        e2e_fn = make_e2e
        e2e_api = "waves of 8 tracks: pinned H2D, upmix_process (device tensors through extract_...'s plan), pinned D2H, double-buffered"
    else:
        # synthetic track, generated on the device (SURVEY.md 8d formula); cfg2: seed 1 + rank
        split = args.split_track and world > 1
        g = torch.Generator(device=dev).manual_seed(1 if split else 1 + rank)
        L = 0.1 * torch.randn(n, device=dev, generator=g)
        R = 0.5 * L + 0.05 * torch.randn(n, device=dev, generator=g)
        if split:
            a_seg, b_seg = sharding.plan_segments(n, world, sharding.largest_hop(ext))[rank]
            lo, hi = sharding.input_range(a_seg, b_seg, plan.halo, n)
            Lh, Rh = L[lo:hi].contiguous(), R[lo:hi].contiguous()
            out = torch.empty((3, 1, b_seg - a_seg), dtype=torch.float32, device=dev)

            def step():
                plan.process_segment(Lh[None], Rh[None], lo, n, a_seg, b_seg, out=out)
            audio_seconds_total = seconds
            h2d, d2h = 8 * (hi - lo), 12 * (b_seg - a_seg)
        else:
            out = torch.empty((3, 1, n), dtype=torch.float32, device=dev)

            def step():
                plan.process_segment(L[None], R[None], 0, n, 0, n, out=out)
            audio_seconds_total = world * seconds
            h2d, d2h = 8 * n, 12 * n

        def make_e2e():
            if split:
                hl = torch.empty(hi - lo, dtype=torch.float32, pin_memory=True)
                hr = torch.empty(hi - lo, dtype=torch.float32, pin_memory=True)
                hl.copy_(Lh)
                hr.copy_(Rh)
                ho = torch.empty((3, b_seg - a_seg), dtype=torch.float32, pin_memory=True)
                dl, dr = torch.empty_like(Lh), torch.empty_like(Rh)

                def run():
                    dl.copy_(hl, non_blocking=True)
                    dr.copy_(hr, non_blocking=True)
                    plan.process_segment(dl[None], dr[None], lo, n, a_seg, b_seg, out=out)
                    ho.copy_(out[:, 0], non_blocking=True)
                    torch.cuda.synchronize(dev)
                    return float(ho[0, (b_seg - a_seg) // 2])
                return run
            hl = torch.empty(n, dtype=torch.float32, pin_memory=True)
            hr = torch.empty(n, dtype=torch.float32, pin_memory=True)
            hl.copy_(L)
            hr.copy_(R)

            def run():
                res = ce.extract_center_left_right_multi_band_in_memory(hl, hr, SR, ext)
                return float(res[0][n // 2])                                 # touch the result on the host
            return run
        e2e_fn = make_e2e
        e2e_api = ("one halo'd time segment per rank: pinned H2D, upmix_process_segment, pinned D2H" if split else
                   "upmix_b200.center_extraction.extract_center_left_right_multi_band_in_memory(pinned CPU tensors)")

    # ---- timed region ------------------------------------------------------------------------------------------------
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
    ms_per_step = float(t.item()) / args.steps
    value = audio_seconds_total / (ms_per_step * 1e-3)

    # ---- per-band timing (single-band plans over one track) and the longest launch ---------------------------------
    band_ms, kernel_ms = [], {}
    if rank == 0 and args.workload != "cfg5" and not args.split_track:
        for e in ext:
            p1 = ce.plan_for([e])
            for _ in range(2):
                p1.process_segment(L[None], R[None], 0, n, 0, n, out=out)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a.record()
            for _ in range(3):
                p1.process_segment(L[None], R[None], 0, n, 0, n, out=out)
            b.record()
            torch.cuda.synchronize(dev)
            band_ms.append(a.elapsed_time(b) / 3)
            p1.release_workspace()

    # ---- end to end ---------------------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        run = e2e_fn()
        run()                                                           # warm-up (allocations, pinned staging)
        run()
        barrier()
        reps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(reps):
            run()
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": audio_seconds_total / float(tt.item()), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": float(tt.item()) * 1e3, "api": e2e_api}
        # bare copies of the same bytes (pinned, both directions at once on two streams): the ceiling of this box.  The two
        # directions move DIFFERENT amounts (8 B up, 12 B down per stereo sample) and slow each other while both are busy
        # (~47 GB/s each against ~54 alone), so the ceiling copies the exact bytes in the exact ratio: `pieces` copies per
        # direction, sized n_in / pieces and n_o / pieces floats, at most 1 GiB each
        n_in, n_o = h2d // 4, d2h // 4
        pieces = max(1, -(-n_o // (1 << 28)))
        p_in, p_o = -(-n_in // pieces), -(-n_o // pieces)
        hb_in = torch.empty(p_in, dtype=torch.float32, pin_memory=True)
        hb_out = torch.empty(p_o, dtype=torch.float32, pin_memory=True)
        db_in, db_out = torch.empty_like(hb_in, device=dev), torch.empty_like(hb_out, device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def copies():
            for _ in range(pieces):
                with torch.cuda.stream(s1):
                    db_in.copy_(hb_in, non_blocking=True)
                with torch.cuda.stream(s2):
                    hb_out.copy_(db_out, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            copies()
        dtc = (time.perf_counter() - t0) / reps
        moved = pieces * (p_in + p_o) * 4
        dtc *= (h2d + d2h) / moved                                       # (rounding of the piece sizes: < 1e-6)
        tc = torch.tensor([dtc], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        e2e["copy_ceiling"] = {"ms_per_step": float(tc.item()) * 1e3, "frac": float(tc.item()) / float(tt.item()),
                               "what": "bare pinned H2D + D2H of exactly the same bytes (8 B up, 12 B down per stereo sample), two streams, "
                                       "all ranks at once, max over ranks"}
        del hb_in, hb_out, db_in, db_out
        # the call exactly as main.py makes it: float64 strided views of an interleaved pageable array (rank 0 only)
        if rank == 0 and args.workload != "cfg5" and not args.split_track:
            wave = np.empty((n, 2), dtype=np.float64)
            wave[:, 0] = L.cpu().numpy()
            wave[:, 1] = R.cpu().numpy()
            def numpy_leg():
                for _ in range(2):                                      # warm-up: staging rings, allocator steady state
                    res = ce.extract_center_left_right_multi_band_in_memory(wave[:, 0], wave[:, 1], SR, ext)
                    del res
                t0 = time.perf_counter()
                for _ in range(2):
                    res = ce.extract_center_left_right_multi_band_in_memory(wave[:, 0], wave[:, 1], SR, ext)
                    _ = float(res[0][n // 2])
                    del res                                             # results are dropped before the next call
                return (time.perf_counter() - t0) / 2
            dtn = numpy_leg()
            os.environ["UPMIX_NUMPY_PINNED_OUT"] = "1"
            dtp = numpy_leg()
            os.environ.pop("UPMIX_NUMPY_PINNED_OUT")
            e2e["numpy"] = {"value": seconds / dtn, "unit": "audio-s/s", "ms_per_step": dtn * 1e3,
                            "host_bytes_read": 16 * n, "what": "float64 strided views of one interleaved pageable array in "
                            "(main.py:43-50), fresh pageable float32 arrays out (np.empty, filled by the copy-out workers); this rank alone",
                            "ratio_to_pinned": dtn / float(tt.item()),
                            "pinned_outputs": {"ms_per_step": dtp * 1e3, "ratio_to_pinned": dtp / float(tt.item()),
                                               "what": "same input, outputs in cached pinned blocks (UPMIX_NUMPY_PINNED_OUT=1)"}}
            del wave

    line = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        try:
            fp32_tflops, sms = _native.measure_fp32_tflops(local_rank)
            fp32_src = "measured in this run (FMA probe kernel)"
        except Exception as ex:  # pragma: no cover
            fp32_tflops, sms, fp32_src = FP32_FALLBACK_TFLOPS, 148, f"fallback ({ex})"
        samples_per_s = value * SR / world                                # per GPU
        w_all = 10 * T_REAL_FFTS * sum(math.log2(s) for s in sizes)
        roofline = {"bound": "fp32", "peak": fp32_tflops, "unit": "TFLOP/s", "peak_source": fp32_src, "sm_count": sms,
                    "whole_path": {"flops_per_sample": w_all, "achieved_tflops": samples_per_s * w_all / 1e12,
                                   "frac": samples_per_s * w_all / 1e12 / fp32_tflops},
                    "hbm_view": {"bound": "hbm", "achieved": samples_per_s * BYTES_PER_SAMPLE / 1e9,
                                 "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": samples_per_s * BYTES_PER_SAMPLE / 1e9 / peaks.get("hbm_gbs", HBM_FALLBACK_GBS),
                                 "peak_source": peak_src + " (MEASURED_PEAKS.json)" if peak_src == "measured" else "fallback"}}
        if band_ms:
            # dominant kernel = the longest single launch of a step: the fused kernel of the dense top band (one launch
            # per step); the band-limited bands run as three shorter launches each (dec_fwd / dec_inv x2, listed beside it)
            fused = [i for i, (s_, e) in enumerate(zip(sizes, ext)) if s_ <= 8192 and not _is_decimated(e)]
            dom = max(fused, key=lambda i: band_ms[i]) if fused else int(np.argmax(band_ms))
            dom_n = sizes[dom]
            dom_flops = 10 * T_REAL_FFTS * math.log2(dom_n) * n
            dom_tflops = dom_flops / (band_ms[dom] * 1e-3) / 1e12
            traffic, traffic_note = None, "no ncu capture of this kernel in profiles/traffic.json"
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                ent = tr.get(f"band_fused_kernel<{dom_n}>")
                if ent and ent.get("source_sha1") == kernel_source_sha1():
                    traffic = ent["dram_bytes_per_sample"] * n
                    traffic_note = f"dram__bytes_read+write per launch scaled from {ent['capture']}"
                elif ent:
                    traffic_note = "profiles/traffic.json was captured from other kernel sources (sha1 differs): refused"
            except Exception:
                pass
            roofline.update({"kernel": f"band_fused_kernel<{dom_n}> (timed as the single-band plan of the N={dom_n} band: one launch "
                                       "of this kernel per step, writing its hops straight to the outputs)",
                             "achieved": dom_tflops, "frac": dom_tflops / fp32_tflops, "traffic": traffic, "traffic_note": traffic_note,
                             "algorithmic": f"10*T*log2(N) = {10 * T_REAL_FFTS * math.log2(dom_n):.0f} flops per stereo sample for this band "
                                            f"(T={T_REAL_FFTS} real FFTs/frame, 75% overlap), {n} samples per launch; 20 B/sample compulsory HBM bytes",
                             "note": "round 1's dominant launch, band_fused_kernel<8192> (5.64 ms, 0.275), no longer runs: the 8192- and "
                                     "65536-point bands take the decimated kernels (band_tflops_nominal below: they do less than the nominal "
                                     "work); this launch took 4.7 ms (0.25) in round 1.  FP32-pipe utilisation, instruction mix and the "
                                     "shared-memory wavefront budget that bound it: profiles/r02_ncu_summary.md, profiles/r02_tuning.md",
                             "band_ms": dict(zip([str(s) for s in sizes], band_ms)),
                             "band_tflops_nominal": {str(s): 10 * T_REAL_FFTS * math.log2(s) * n / (m * 1e-3) / 1e12 for s, m in zip(sizes, band_ms)}})
        else:
            roofline.update({"kernel": "whole path (no per-band timing in this mode)", "achieved": roofline["whole_path"]["achieved_tflops"],
                             "frac": roofline["whole_path"]["frac"], "traffic": None})
        scaling = "strong" if (args.workload == "cfg5" or args.split_track) else "weak"
        line = {"metric": "realtime factor (audio-s/s, 48 kHz stereo)", "value": value, "unit": "audio-s/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.workload, world, sizes, split=args.split_track), "roofline": roofline, "clocks": clocks,
                "gpu_launches": int(launches), "e2e": e2e}
        if reduced:
            line["config"]["note"] = "REDUCED workload size (not the headline workload)"
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(W["edges"], min(240.0 if args.workload == "cfg2" else 60.0, float(seconds)),
                              seed=1000 if args.workload == "cfg5" else 1)
            cb.pop("seconds", None)
            line["cpu_baseline"] = cb
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def kernel_source_sha1():
    """Fingerprint of the fused kernel's sources; profiles/traffic.json records the one its ncu capture was taken with."""
    import hashlib
    h = hashlib.sha1()
    for f in ("fft_device.cuh", "upmix_fused.cuh", "upmix_kernels.cuh"):
        h.update(open(os.path.join(ROOT, "upmix_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def _is_decimated(e):
    """Does this extractor's band take the decimated kernels (every live bin below 512 and below block_size/16)?"""
    g = e.band_gain()
    nz = np.nonzero(g)[0]
    top = int(nz.max()) if nz.size else 0
    p = 128 if top < 128 else 256 if top < 256 else 512 if top < 512 else 0
    return bool(p) and e.block_size >= 16 * p and 4 * e.hop_size == e.block_size


if __name__ == "__main__":
    main()
