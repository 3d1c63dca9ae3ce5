#!/usr/bin/env python
"""What the planner makes of the headline neighbourhood (host only, no GPU): dispatches per tile with and
without the super-instruction pass, the composition of the fused stream, reductions, data slots.
Usage: python tools/plan_stats.py [gram|eval] > profiles/r1_plan_composition.txt"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rils_rols_b200 import workloads as W  # noqa: E402
from tests import isa_emu as I  # noqa: E402

kind = {"gram": I.KIND_GRAM, "eval": I.KIND_EVAL}[sys.argv[1] if len(sys.argv) > 1 else "gram"]
batch = W.cfg5_neighbourhood()
names = {}
for k, v in vars(I).items():
    if k.startswith("RI_") and isinstance(v, int) and k not in ("RI_FIRST_M", "RI_OPCOUNT"):
        names.setdefault(v, k[3:])
FAMILIES = ("PIN0", "LDP0", "USEP0", "MULP0", "DIVP0", "RDIVP0", "CMULP0", "CDIVP0", "LDPMUL_M0", "LDPDIV_M0", "LDMDIVP0")


def cls(w0):
    o = int(w0) & 0xFF
    md = "+MDOT" if (int(w0) & I.RR_THEN_MDOT and I.md_fusable(o)) else ""
    for base in FAMILIES:
        v = getattr(I, "RI_" + base)
        if v <= o < v + I.RR_NREG:
            return base[:-1] + " j" + md
    return names.get(o, str(o)) + md


def plan(fuse):
    if fuse:
        os.environ.pop("RR_B200_DEBUG_NO_FUSE", None)
    else:
        os.environ["RR_B200_DEBUG_NO_FUSE"] = "1"
    p = I.Plan(batch, 20, kind, tile_cols=23)
    os.environ.pop("RR_B200_DEBUG_NO_FUSE", None)
    return p


plain, fused = plan(False), plan(True)
for tag, p in (("without super-instructions", plain), ("with super-instructions   ", fused)):
    ops = p.ins["w0"] & 0xFF
    slots = int(((p.ins["w0"] & 0xFFFF) == (I.RI_NOP | I.RR_MDOT_ROWS << 8)).sum())
    nops = int((ops == I.RI_NOP).sum()) - slots
    comb = int((ops == I.RI_COMBINE).sum())
    print(f"{tag}: {len(ops) - slots - nops:6d} dispatches per tile ({comb} of them RI_COMBINE), {slots} data slots, "
          f"{nops} padding NOPs, {p.n_dots} reductions, {p.max_tile_cols} tile columns, w_issued {p.w_issued:.0f}")
print(f"batch: {batch.n_cand} candidates, {batch.n_terms} term instances, {fused.n_terms_distinct} distinct terms, "
      f"w_contract {fused.w_contract:.0f}")
print("\ncomposition of the fused stream (data slots and padding left out):")
c = collections.Counter(cls(w) for w in fused.ins["w0"] if (int(w) & 0xFF) != I.RI_NOP)
for k, v in c.most_common():
    print(f"  {k:18s}{v:6d}")
