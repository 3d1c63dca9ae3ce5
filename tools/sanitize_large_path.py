#!/usr/bin/env python
"""One pass of the large-n path (4-samples-per-thread PTX core on full tiles, C++ interpreter on the partial
tile, double-double and residual passes) on a problem small enough for compute-sanitizer:

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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536 + 300
n_cand = int(sys.argv[2]) if len(sys.argv) > 2 else 768
X, y = W.cfg5_data(n)
nb = W.cfg5_neighbourhood()
sub = nb.subset(list(range(min(n_cand, nb.n_cand))))
with Engine(X, y, device=0, flags=B.FLAG_FORCE_GRAM) as eng:
    r1 = eng.score(sub)
    r2 = eng.score(sub)
    st = eng.stats()
same = np.array_equal(np.asarray(r1.ssr), np.asarray(r2.ssr), equal_nan=True) and \
    np.array_equal(np.asarray(r1.coef), np.asarray(r2.coef), equal_nan=True)
print({k: st[k] for k in ("sweep_launches", "kernel_launches", "refined", "dd", "nonfinite")}, "deterministic:", same)
sys.exit(0 if same else 1)
