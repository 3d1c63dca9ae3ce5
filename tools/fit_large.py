#!/usr/bin/env python
"""fit() on the 2^24 x 20 synthetic set through the pybind11 boundary: the driver shards the rows over all visible GPUs
inside ONE engine object (rr_engine_create_sharded, ncclCommInitAll) - no launcher, no torch. Prints wall time,
the engine's statistics and per-GPU utilisation sampled with NVML while fit() runs.
  python tools/fit_large.py [n_rows] [max_fit_calls]        RR_B200_GPUS=1 forces one GPU"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rils_rols_b200  # noqa: E402
from rils_rols_b200 import workloads  # noqa: E402

M = rils_rols_b200.driver_module()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
X, y = workloads.cfg5_data(n)
util, stop = [], False


def sample():
    try:
        import pynvml as nv

        nv.nvmlInit()
        hs = [nv.nvmlDeviceGetHandleByIndex(i) for i in range(nv.nvmlDeviceGetCount())]
        while not stop:
            util.append([nv.nvmlDeviceGetUtilizationRates(h).gpu for h in hs])
            time.sleep(0.05)
    except Exception as ex:  # pragma: no cover
        util.append(str(ex))


th = threading.Thread(target=sample, daemon=True)
th.start()
rr = M.rils_rols(False, calls, 100000, 0.001, 50, 1.0, False, 12345)
t = time.perf_counter()
rr.fit(X.reshape(-1, 1), y, X.shape[0], X.shape[1])
wall = time.perf_counter() - t
stop = True
th.join(timeout=1)
u = np.array([r for r in util if isinstance(r, list)])
busy = (u > 50).mean(axis=0).round(3).tolist() if u.size else None
print(json.dumps({"n": n, "d": int(X.shape[1]), "fit_calls": rr.get_fit_calls(), "fit_wall_s": wall, "model": rr.get_model_string(),
                  "engine": {k: (float(v) if isinstance(v, float) else int(v)) for k, v in rr.get_engine_stats().items()},
                  "gpus_seen": int(u.shape[1]) if u.size else 0, "fraction_of_samples_gpu_busy": busy,
                  "RR_B200_GPUS": os.environ.get("RR_B200_GPUS")}))
