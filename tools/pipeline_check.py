#!/usr/bin/env python
"""Smallest check of run_gram's two-half planning (second half on a helper thread): one large-n, >= 1024
candidate batch scored with and without it must agree bit for bit."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rils_rols_b200 import batch as B  # noqa: E402
from rils_rols_b200 import workloads as W  # noqa: E402
from rils_rols_b200.engine import Engine  # noqa: E402

t0 = time.perf_counter()
X, y = W.cfg5_data(40000)
sub = W.cfg5_neighbourhood().subset(range(0, 1200))
with Engine(X, y, device=0, flags=B.FLAG_FORCE_GRAM) as eng:
    a = eng.score(sub)
    os.environ["RR_B200_PIPELINE"] = "0"
    b = eng.score(sub)
    n_sweeps = eng.stats()["sweep_launches"]
same = np.array_equal(a.ssr, b.ssr, equal_nan=True) and np.array_equal(a.coef, b.coef, equal_nan=True)
print("PIPELINE_OK" if same else "PIPELINE_MISMATCH", "sweeps", n_sweeps, "%.1f s" % (time.perf_counter() - t0))
sys.exit(0 if same else 1)
