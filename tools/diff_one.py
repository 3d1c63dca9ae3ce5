import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rils_rols_b200 import batch as B
from rils_rols_b200.engine import Engine
sys.path.insert(0, os.path.join(ROOT, "tools"))
from diff_paths import decode  # noqa

def run(X, y, batch, s, flags):
    os.environ["RR_B200_S"] = str(s)
    with Engine(X, y, flags=flags) as e:
        r = e.score(batch)
        return np.array(r.ssr, copy=True), np.array(r.coef, copy=True), np.array(r.flags, copy=True), e.stats()

name, pf = "cfg1_toy", "ls3"
z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
X0, y0 = z["X"], z["y"]
n = 40960
rng = np.random.default_rng(7)
idx = rng.integers(0, X0.shape[0], n)
X = X0[idx] * (1.0 + 1e-3 * rng.standard_normal((n, X0.shape[1])))
y = y0[idx] + 1e-3 * rng.standard_normal(n)
full = B.Batch.load_fields(z, pf + "_")
for sel in ([48], [62], [123], [0, 48], list(range(40, 50)), list(range(0, 130))):
    b = full.subset(sel)
    for flags in (B.FLAG_FORCE_GRAM,):
        r4 = run(X, y, b, 4, flags)
        r1 = run(X, y, b, 1, flags)
        for i, c in enumerate(sel):
            if c not in (48, 62, 123): continue
            sl = b.coef_slice(i)
            print(f"sel {sel[:3]}..({len(sel)}) cand {c}: S4 ssr {r4[0][i]:.6e} coef {r4[1][sl]} flags {r4[2][i]:#x} | S1 ssr {r1[0][i]:.6e} coef {r1[1][sl]} flags {r1[2][i]:#x}")
    print("  stats S4", {k: r4[3][k] for k in ('sweep_launches','refined','dd')}, " S1", {k: r1[3][k] for k in ('sweep_launches','refined','dd')})
