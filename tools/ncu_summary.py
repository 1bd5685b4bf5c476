#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + opcode mix + stall reasons (reads `ncu -i ... --page raw/source --csv`)."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:75s} {vals[i]:>18s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
op, samp, thr, stalls = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ti = ts = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    if not m:
        continue
    o = m.group(2).split(".")[0]
    if o == "MUFU":
        o = m.group(2)
    try:
        n = int(r[ix["Instructions Executed"]] or 0)
        s = int(r[ix["# Samples"]] or 0)
    except ValueError:        # a repeated header: the report holds several kernels (the totals below cover all of them)
        continue
    op[o] += n; samp[o] += s; ti += n; ts += s
    thr[o] += int(r[ix["Thread Instructions Executed"]] or 0)
    for c in stall_cols:
        if r[ix[c]]:
            stalls[c] += int(r[ix[c]])
print(f"total warp instructions {ti}, samples {ts}")
fp64 = sum(op[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"FP64 share of issued instructions: {fp64 / ti:.3f}")
for o, n in op.most_common(24):
    print(f"  {o:14s} {n / ti:6.3f} inst   samples {samp[o] / max(ts, 1):6.3f}   avg active threads {thr[o] / max(n, 1):5.1f}")
tot = sum(stalls.values())
print("stalls:", {k: round(v / tot, 3) for k, v in stalls.most_common(9)})
