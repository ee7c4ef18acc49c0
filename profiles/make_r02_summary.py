#!/usr/bin/env python3
"""Builds profiles/r02_ncu_summary.md and profiles/traffic.json from the round-2 captures in gpurun_out/
(r02_cfg2.ncu-rep, fused1024_a.ncu-rep, fb512_a.ncu-rep, fb1024_a.ncu-rep, r02_flops.csv)."""
import collections, csv, hashlib, io, json, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
SAMPLES = 28800000


def summary(rep):
    return subprocess.run(["python", f"{ROOT}/profiles/ncu_summary.py", os.path.join(G, rep), str(SAMPLES)], capture_output=True, text=True).stdout


s_cfg2, s_fb1024, s_fb512, s_fused = summary("r02_cfg2.ncu-rep"), summary("fb1024_a.ncu-rep"), summary("fb512_a.ncu-rep"), summary("fused1024_a.ncu-rep")
per_kernel = [float(l.split("|")[9]) for l in s_cfg2.splitlines()[2:] if l.strip()]
times = [float(l.split("|")[2]) for l in s_cfg2.splitlines()[2:] if l.strip()]
rows = list(csv.reader(open(os.path.join(G, "r02_flops.csv"))))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    d.setdefault((r[ii], r[ki].replace("upmix::", "").split("(")[0].replace("void ", "")), {})[r[mi]] = float(r[vi].replace(",", ""))
tab = ("| kernel | FP32 thread instr (G) | of which scalar fadd / fmul / ffma (G) | pipe_fma warp instr (M) | pipe_alu (M) | pipe_lsu (M) | pipe_xu (M) | "
       "all warp instr (M) | nominal GFLOP |\n|---|---|---|---|---|---|---|---|---|\n")
nominal = {"dec_fwd_kernel<512, 128>": "65536-point band: 800 x 28.8 M = 23.0 (whole band: fwd + mask + 2 inv)",
           "dec_fwd_kernel<512, 16>": "8192-point band: 650 x 28.8 M = 18.7 (whole band)", "band_fused_kernel<1024, 0, 1>": "1024-point band: 500 x 28.8 M = 14.4"}
for k, m in d.items():
    g = lambda n: m.get(n, 0.0)
    tab += (f"| {k[1]} | {g('smsp__sass_thread_inst_executed_op_fp32_pred_on.sum') / 1e9:.2f} | {g('smsp__sass_thread_inst_executed_op_fadd_pred_on.sum') / 1e9:.2f} / "
            f"{g('smsp__sass_thread_inst_executed_op_fmul_pred_on.sum') / 1e9:.2f} / {g('smsp__sass_thread_inst_executed_op_ffma_pred_on.sum') / 1e9:.2f} | "
            f"{g('sm__inst_executed_pipe_fma.sum') / 1e6:.1f} | {g('sm__inst_executed_pipe_alu.sum') / 1e6:.1f} | {g('sm__inst_executed_pipe_lsu.sum') / 1e6:.1f} | "
            f"{g('sm__inst_executed_pipe_xu.sum') / 1e6:.1f} | {g('smsp__inst_executed.sum') / 1e6:.1f} | {nominal.get(k[1], '')} |\n")
md = f"""# Round 2: Nsight Compute summary of the final kernels (B200, sm_100a, driver 580, `--clock-control none`)

Commands (each after the same command had exited 0 without ncu in the same gpurun call):

    ncu --set full --clock-control none --import-source on -k regex:"dec_|band_" -s 8 -c 8 -o r02_cfg2 python profiles/profile_driver.py 600 2
    ncu --metrics smsp__sass_thread_inst_executed_op_{{fadd,fmul,ffma,fp32}}_pred_on.sum,sm__inst_executed_pipe_{{fma,fmaheavy,alu,lsu,xu}}.sum,... -> profiles/r02_fp32_counters.csv
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline -> profiles/r02_launches.csv

Workload of the first two: the bench plan (cfg 2: 65536 / 8192 / 1024 points) over a 600 s track (28.8 M stereo samples per
launch), second pass captured.  Under ncu the launches are serialised and cold-cache: compare shares, not absolutes (bench, same
code, 1-hour track: 13.6 ms per step = 2.27 ms per 600 s; sum of the eight launches below: {sum(times) / 1e3:.2f} ms; the dominant launch,
`band_fused_kernel<1024>`, is {100 * times[-1] / sum(times):.0f} % of the step here and 4.62 / 13.6 = 34 % in the bench).

## One step of the bench plan, launch by launch

{s_cfg2}
* `dec_fwd` / `dec_mask` / `dec_inv ... 0` (Ls + i Rs) / `dec_inv ... 1` (centre) are the decimated kernels of the two band-limited
  bands (upmix_dec.cuh); template arguments: P (points per decimated sequence), Q (sequences), centre, accumulating.  The
  65536-point band is the first of the plan and stores; the 8192-point band adds to the outputs: its previous sums come in
  through `cp.async` (LDGSTS) into shared memory, requested one frame ahead (two buffers).
* DRAM bytes per stereo sample of the whole step: {sum(per_kernel):.1f} B against 20 B compulsory (8 in + 12 out); round 1 moved ~243 B.  The
  65536-point band: {sum(per_kernel[0:4]):.1f} B (round 1: 180 B, its four-step scratch is gone), the 8192-point band {sum(per_kernel[4:7]):.1f} B and the 1024-point band
  {per_kernel[7]:.1f} B, 24 B of each being the read-modify-write of the outputs that the band-order sum costs.
* What bounds them: every kernel issues on 37-49 % of cycles with 14-15 resident warps per SM (128 registers per thread).  The
  forward kernels are closest to a hardware limit: the L1 data pipe (LSU wavefronts) runs at 70-79 % (`mio_throttle` leads for the
  128-sequence band: 96 scalar global loads per thread and frame); the inverse kernels wait on dependent loads
  (`long_scoreboard`), fixed-latency dependencies (`wait`) and their four barriers per frame.

## FP32 work actually executed (real against nominal flops)

{tab}
The `op_fadd / fmul / ffma` counters see only the scalar forms; the kernels' arithmetic is packed (`FADD2 / FMUL2 / FFMA2`, two
lanes per instruction: profiles/r02_sass_ops.txt), which `op_fp32` counts once per thread instruction.  With two operations per
packed add / multiply and four per packed FMA the 1024-point kernel executes about 1.5x its nominal 14.4 GFLOP (mask, windows,
twiddle recurrences); the decimated bands execute about 0.9x their nominal figure, because they transform 512 points per
sequence instead of 8192 / 65536 (N log P + O(KQ) instead of N log N) -- which is where their 24-32 nominal TFLOP/s come from.

## Dense-band kernels (single-band plans, 600 s, storing)

One frame per CTA (`band_fused_kernel<1024>`, the bench's dominant launch; `roofline.traffic` in bench.py comes from this capture):

{s_fused}
Frame-batched (`band_fb_kernel`, upmix_fb.cuh; serves 256 and 512 points by default, 16 frames per tile; the 1024-point capture below runs 8-frame tiles):

{s_fb512}
{s_fb1024}
The frame-batched 1024-point kernel (8-frame tiles, two CTAs of 256 threads per SM; the 16-frame tile of one 512-thread CTA
took 1015 us) executes 13 % fewer warp instructions than the one-frame kernel with 21 % fewer shared-memory wavefronts, but
its tiles leave 23 KB of L1 (shared-memory carve-out 233 KB): 59 % of the sectors its table loads ask for (pass twiddles,
windows, gains) come from L2 -- `long_scoreboard` on the multiply after the twiddle `LDG.128` leads its stalls -- where the
one-frame kernel (88 KB of L1) hits 99 %.  4.71 against 4.60 ms per band-hour, so 1024 points keep the one-frame kernel
(UPMIX_FB_MAX_N=1024 switches).  At 512 points (two CTAs per SM) and 256 points (four) the frame-batched kernel wins: 4.11
against 5.01 and 3.71 against 5.65 ms per band-hour.
"""
open(os.path.join(ROOT, "profiles", "r02_ncu_summary.md"), "w").write(md)
hh = hashlib.sha1()
for f in ("fft_device.cuh", "upmix_fused.cuh", "upmix_kernels.cuh"):
    hh.update(open(os.path.join(ROOT, "upmix_b200", "csrc", f), "rb").read())


def dram(rep):
    out = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    a, u, b = rr[0], rr[1], rr[2]
    sc = lambda k: float(b[a.index(k)].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[a.index(k)]]
    return sc("dram__bytes_read.sum"), sc("dram__bytes_write.sum")


rd, wr = dram("fused1024_a.ncu-rep")
tr = {"band_fused_kernel<1024>": {"dram_bytes_per_sample": (rd + wr) / SAMPLES,
                                   "capture": "profiles/r02_ncu_summary.md (600 s track, ncu --set full, single-band plan storing its hops: what bench.py times)",
                                   "dram_read_bytes": rd, "dram_write_bytes": wr, "samples": SAMPLES, "source_sha1": hh.hexdigest(),
                                   "accumulating": {"dram_bytes_per_sample": per_kernel[7], "capture": "profiles/r02_ncu_summary.md, third band of the bench plan"}}}
json.dump(tr, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("written; traffic", tr["band_fused_kernel<1024>"]["dram_bytes_per_sample"])
