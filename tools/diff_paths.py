#!/usr/bin/env python
"""Differential check of the two interpreter paths on the GPU: the PTX core (4 samples per thread, full
tiles) against the generic C++ interpreter (RR_B200_S=1), same batches, same data. Batches: every recorded
neighbourhood / perturbation set of the golden fixtures, data rows replicated with jitter to `n` rows."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rils_rols_b200 import batch as B  # noqa: E402
from rils_rols_b200.engine import Engine  # noqa: E402


def decode(batch, t):
    c0, c1 = batch.term_code_begin[t], batch.term_code_begin[t + 1]
    st = []
    for w in batch.code[c0:c1]:
        op, arg = int(w) & 0xFF, int(w) >> 8
        if op == B.OP_CONST: st.append(repr(float(batch.consts[arg])))
        elif op == B.OP_VAR: st.append(f"x{arg}")
        elif B.ARITY[op] == 1: st.append(f"{B.OP_NAMES[op]}({st.pop()})")
        else:
            b = st.pop(); a = st.pop(); st.append(f"({a} {B.OP_NAMES[op]} {b})")
    return st[-1]


def run(X, y, batch, s, flags):
    os.environ["RR_B200_S"] = str(s)
    with Engine(X, y, flags=flags) as e:
        r = e.score(batch)
        return np.array(r.ssr, copy=True), np.array(r.coef, copy=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
    bad_total = 0
    for name in ("cfg1_toy", "cfg2_diabetes", "cfg3_breast_cancer"):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        X0, y0 = z["X"], z["y"]
        rng = np.random.default_rng(7)
        idx = rng.integers(0, X0.shape[0], n)
        X = X0[idx] * (1.0 + 1e-3 * rng.standard_normal((n, X0.shape[1])))
        y = y0[idx] + 1e-3 * rng.standard_normal(n)
        prefixes = sorted({k[: k.index("_mode")] for k in z.keys() if k.endswith("_mode")})
        for pf in prefixes:
            batch = B.Batch.load_fields(z, pf + "_")
            flags = B.FLAG_FORCE_GRAM
            s4, c4 = run(X, y, batch, 4, flags)
            s1, c1 = run(X, y, batch, 1, flags)
            fin = np.isfinite(s4) & np.isfinite(s1)
            mism_fin = np.isfinite(s4) != np.isfinite(s1)
            rel = np.zeros_like(s4)
            rel[fin] = np.abs(s4[fin] - s1[fin]) / np.maximum(np.abs(s1[fin]), 1e-300)
            bad = np.where((rel > 1e-7) | mism_fin)[0]
            print(f"{name}/{pf}: mode {batch.mode} cands {batch.n_cand} finite {int(fin.sum())} max rel {rel.max():.2e} bad {len(bad)}")
            bad_total += len(bad)
            for c in bad[:6]:
                t0, t1 = batch.cand_term_begin[c], batch.cand_term_begin[c + 1]
                print(f"   cand {c}: ssr S4 {s4[c]!r} S1 {s1[c]!r}  terms: " + " | ".join(decode(batch, t) for t in range(t0, t1)))
    print("BAD", bad_total)


if __name__ == "__main__":
    main()
