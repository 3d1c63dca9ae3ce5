import sys, os, time, numpy as np
sys.path.insert(0, "/root/repo")
os.environ["RR_B200_VERBOSE"] = "1"
from rils_rols_b200 import workloads as W
from rils_rols_b200.engine import Engine
n = 1 << 24
X, y = W.cfg5_data(n)
idx = np.random.default_rng(0).permutation(n).astype(np.int32)
for label, kw in (("plain rowmajor", None), ("identity index", np.arange(n, dtype=np.int32)), ("shuffled index", idx)):
    t = time.perf_counter()
    eng = Engine(X, y) if kw is None else Engine.sharded(X, y, n_gpus=1, row_index=kw)
    print(label, "%.3f s" % (time.perf_counter() - t), "ingest_ms", eng.stats()["ingest_ms"], flush=True)
    eng.close()
