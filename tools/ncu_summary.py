#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed): headline metrics, SASS opcode mix
and the top stall lines. Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"== kernel: {name}")
    for h, u, v in zip(hdr, units, r):
        if h in WANT:
            print(f"  {h} [{u}] = {v}")
    st = [(float(v), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h, v in zip(hdr, r)
          if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
    print("  stall reasons (warps stalled per issue): " + ", ".join(f"{n}={x:.2f}" for x, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = totsamp = 0
byop, samp, data = collections.Counter(), collections.Counter(), []
for r in rows[hi + 1:]:
    if len(r) <= ie:
        continue
    try:
        n, s = int(r[ie]), int(r[isamp])
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    op = m.group(2).split(".")[0] if m else "?"
    byop[op] += n
    samp[op] += s
    tot += n
    totsamp += s
    data.append((n, s, r[ia]))
print(f"== SASS mix (first kernel in the report): {tot} warp-instructions, {totsamp} samples")
for op, n in byop.most_common(24):
    print(f"  {op:10s} {n / tot * 100:6.2f}% of instructions   {samp[op] / max(totsamp, 1) * 100:6.2f}% of stall samples")
fp64 = sum(byop[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"  FP64-pipe instructions: {fp64 / tot * 100:.2f}% of all issued")
print("== top stall lines")
for n, s, line in sorted(data, key=lambda x: -x[1])[:25]:
    print(f"  {s / max(totsamp, 1) * 100:5.2f}% samples  {n / tot * 100:5.2f}% exec   {line.strip()[:90]}")
