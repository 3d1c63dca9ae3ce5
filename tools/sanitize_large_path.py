#!/usr/bin/env python
"""One pass of the large-n path (Gram sweeps on the G8 kernel - or the row machine with RR_B200_R8=1
RR_B200_R8_MIN_FILL=0 - on full tiles and the partial last tile, classic piece of the wide candidates, double-double
and residual passes) on a problem small enough for compute-sanitizer (the large-n plan shapes need more 512-row
tiles than resident blocks: n >= 151 552):

    compute-sanitizer --tool racecheck python tools/sanitize_large_path.py
    compute-sanitizer --tool memcheck  python tools/sanitize_large_path.py
    compute-sanitizer --tool synccheck python tools/sanitize_large_path.py

Prints the engine statistics; exits non-zero when the result disagrees with a second, identical call
(the reductions are bit-deterministic)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rils_rols_b200 import batch as B  # noqa: E402
from rils_rols_b200 import workloads as W  # noqa: E402
from rils_rols_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 18) + 300
n_cand = int(sys.argv[2]) if len(sys.argv) > 2 else 384
X, y = W.cfg5_data(n)
nb = W.cfg5_neighbourhood()
sub = nb.subset(list(range(min(n_cand, nb.n_cand))))
with Engine(X, y, device=0, flags=B.FLAG_FORCE_GRAM) as eng:
    r1 = eng.score(sub)
    r2 = eng.score(sub)
    st = eng.stats()
same = np.array_equal(np.asarray(r1.ssr), np.asarray(r2.ssr), equal_nan=True) and \
    np.array_equal(np.asarray(r1.coef), np.asarray(r2.coef), equal_nan=True)
print({k: st[k] for k in ("sweep_launches", "kernel_launches", "refined", "dd", "nonfinite", "row_groups")}, "deterministic:", same)
sys.exit(0 if same else 1)
