#!/usr/bin/env python
"""Per-handler view of an .ncu-rep of the sweep kernel (read here with `ncu -i`, no GPU needed): the SASS is cut
into regions at every indirect branch (BRX ends a handler's dispatch), unconditional branch and EXIT, and each
region is listed with its share of stall samples and executed instructions, its length, how often its busiest
instruction ran, and its FP64 share. Usage: tools/ncu_regions.py gpurun_out/prof.ncu-rep [n_regions]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[0], r[ia], int(r[ie] or 0), int(r[isamp] or 0)) for r in rows[hi + 1:] if len(r) > isamp]
tot_e, tot_s = sum(d[2] for d in data), sum(d[3] for d in data)


def opc(s):
    t = s.strip().split()
    return t[1] if t[0].startswith("@") else t[0]


regions, cur = [], []
for d in data:
    cur.append(d)
    o = opc(d[1])
    if o.startswith("BRX") or o.startswith("EXIT") or o.startswith("RET") or (o == "BRA" and not d[1].strip().startswith("@")):
        regions.append(cur)
        cur = []
if cur:
    regions.append(cur)
out = []
for reg in regions:
    e, s = sum(d[2] for d in reg), sum(d[3] for d in reg)
    if e == 0:
        continue
    fp = sum(d[2] for d in reg if opc(d[1])[:4] in ("DFMA", "DMUL", "DADD"))
    out.append((s, e, max(d[2] for d in reg), len(reg), reg[0][0], " ".join(opc(d[1]) for d in reg[:7]), fp))
out.sort(reverse=True)
print(f"# {rep}: {tot_e:.3e} warp instructions, {tot_s} stall samples; regions end at BRX / BRA / EXIT")
print("# samples%  exec%  n_sass  busiest_instr_execs  fp64%  first_address  first opcodes")
for s, e, mx, n, addr, ops, fp in out[:top]:
    print(f"{100 * s / tot_s:6.1f} {100 * e / tot_e:6.1f} {n:6d} {mx:12.3e} {100 * fp / max(e, 1):5.0f}  @{addr[-5:]}  {ops}")
