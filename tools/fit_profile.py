#!/usr/bin/env python
"""Where does fit() spend its time on a small config? Runs BASELINE config 2 with RR_B200_VERBOSE=1 and
sums the engine's per-pass timing lines (stderr) by label."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time; sys.path.insert(0, %r)
import numpy as np
from rils_rols_b200 import workloads
import rils_rols_b200

M = rils_rols_b200.driver_module()
name = sys.argv[1]
X, y = workloads.config_data(name)
rr = M.rils_rols(name == "cfg3_breast_cancer", 100000, 100000, 0.001, 20, 1.0, False, 12345)
t = time.perf_counter(); rr.fit(X.reshape(-1, 1), y, X.shape[0], X.shape[1]); print("wall", time.perf_counter() - t, rr.get_engine_stats())
''' % ROOT
env = dict(os.environ, RR_B200_VERBOSE="1")
r = subprocess.run([sys.executable, "-c", code, sys.argv[1] if len(sys.argv) > 1 else "cfg2_diabetes"], capture_output=True, text=True, env=env)
tot = collections.Counter(); cnt = collections.Counter()
for line in r.stderr.splitlines():
    m = re.match(r"\[rr_b200\] (.+?)\s+([0-9.]+) ms", line)
    if m:
        tot[m.group(1)] += float(m.group(2)); cnt[m.group(1)] += 1
print(r.stdout.strip()[-600:])
for k, v in tot.most_common():
    print(f"{k:32s} {v:10.1f} ms in {cnt[k]:6d} calls ({v / cnt[k]:.3f} ms each)")
