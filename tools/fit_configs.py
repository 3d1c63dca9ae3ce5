#!/usr/bin/env python
"""fit() wall time of BASELINE configs 1-4 through the drop-in front end (GPU), next to the unmodified
reference on one host core where oracle/_ref is available. Prints one JSON line per config."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rils_rols_b200 import workloads  # noqa: E402
import rils_rols_b200  # noqa: E402

M = rils_rols_b200.driver_module()

try:
    from oracle import pyoracle as O

    R = O.load_ref()
except Exception:
    R = None


def run(mod, cls, X, y, calls, mc, seconds=100000):
    rr = mod.rils_rols(cls, calls, seconds, 0.001, mc, 1.0, False, 12345)
    t = time.perf_counter()
    rr.fit(X.reshape(-1, 1), y, X.shape[0], X.shape[1])
    wall = time.perf_counter() - t
    yp = rr.predict(X.reshape(-1, 1), X.shape[0], X.shape[1])
    if cls:
        score = float(np.mean((yp >= 0.5) == (y >= 0.5)))
    else:
        score = float(1 - ((y - yp) ** 2).sum() / ((y - y.mean()) ** 2).sum())
    return dict(wall_s=wall, total_time=rr.get_total_time(), fit_calls=rr.get_fit_calls(), train_score=score,
                model=rr.get_model_string())


ref_too = "--ref" in sys.argv
# one-time process costs (CUDA context, loading the kernels of librr_b200.so) are not part of fit(): a small fit
# first, reported on its own line
_X, _y = workloads.config_data("cfg1_toy")
_t = time.perf_counter()
run(M, False, _X, _y, 300, 50)
print(json.dumps({"warm_up": "first fit() of the process, 300 fitness calls on cfg1 (CUDA context + module load)",
                  "wall_s": time.perf_counter() - _t}), flush=True)
for name, cls, mc, calls in (("cfg1_toy", False, 50, 100000), ("cfg2_diabetes", False, 20, 100000),
                             ("cfg3_breast_cancer", True, 20, 100000)):
    X, y = workloads.config_data(name)
    out = {"config": name, "n": int(X.shape[0]), "d": int(X.shape[1]), "b200": run(M, cls, X, y, calls, mc)}
    if ref_too and R is not None:
        out["reference_1core"] = run(R, cls, X, y, calls, mc)
    print(json.dumps(out), flush=True)
if "--skip4" in sys.argv:
    sys.exit(0)
X, y = workloads.cfg4_data(1_000_000, 10)
out = {"config": "cfg4_1Mx10", "n": 1000000, "d": 10, "b200": run(M, False, X, y, 100000, 50)}
if ref_too and R is not None:
    r = run(R, False, X, y, 300, 50)
    out["reference_1core_300_calls"] = r
print(json.dumps(out), flush=True)
