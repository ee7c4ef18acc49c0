#!/usr/bin/env python3
"""Markdown summary of an ncu report (one column block per kernel launch): python profiles/ncu_summary.py rep.ncu-rep [samples_per_launch]
Prints the rows this repo's roofline discussion uses: time, registers, occupancy, issue / FMA / L1 data-pipe utilisation, DRAM
bytes, FP32 thread-instruction counts (real flops against the nominal 10*T*log2 N), top stall reasons."""
import csv, io, subprocess, sys
rep = sys.argv[1]
samples = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units, body = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(h)}
def val(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return float("nan")
def unit(k):
    return units[col[k]] if k in col else ""
print(f"| kernel | time us | regs | warps act % | issue % | FMA pipe % | L1 data pipe % | DRAM rd+wr MB | B/sample | warp inst M | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
stalls = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for r in body:
    name = r[col["Kernel Name"]].replace("upmix::", "").split("(")[0].replace("void ", "")
    t = val(r, "gpu__time_duration.sum")
    t_us = t / 1e3 if unit("gpu__time_duration.sum") in ("ns", "nsecond") else t if unit("gpu__time_duration.sum") in ("us", "usecond") else t * 1e3
    def scaled(k):
        v, u = val(r, k), unit(k)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
    fadd = val(r, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum")
    fmul = val(r, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum")
    ffma = val(r, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum")
    st = sorted(((val(r, k), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stalls), reverse=True)[:4]
    bps = f"{(rd + wr) / samples:.1f}" if samples else "-"
    print(f"| {name} | {t_us:.1f} | {val(r, 'launch__registers_per_thread'):.0f} | {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{val(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active'):.1f} | {val(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{val(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f} ({val(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6:.0f} M shared wavefronts, {val(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:.1f} M conflicts) | "
          f"{(rd + wr) / 1e6:.0f} | {bps} | {val(r, 'smsp__inst_executed.sum') / 1e6:.1f} | "
          + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |")
