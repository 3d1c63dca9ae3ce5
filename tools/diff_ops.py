#!/usr/bin/env python
"""Element-wise differential check on the GPU: EVAL_ONLY programs (no OLS), PTX core (S=4) against the
generic interpreter (S=1). Prints one line per program."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rils_rols_b200 import batch as B  # noqa: E402
from rils_rols_b200.engine import Engine  # noqa: E402

n = 40960
rng = np.random.default_rng(3)
X = np.column_stack([rng.integers(1, 100, n).astype(float), rng.integers(1, 100, n).astype(float),
                     rng.uniform(-3, 3, n), rng.uniform(0.1, 3, n)])
y = rng.standard_normal(n)
v = B.Expr.var
c = B.Expr.const
progs = {
    "x0/x1": v(0) / v(1),
    "c/x1": -163.34342422886692 / v(1),
    "x1/c": v(1) / -163.34342422886692,
    "-1/(c/x1)": -1.0 / (-163.34342422886692 / v(1)),
    "x2/x3": v(2) / v(3),
    "x3/x2": v(3) / v(2),
    "x0/(x1-x1)": v(0) / (v(1) - v(1)),
    "0/x1": (v(1) - v(1)) / v(1),
    "1/(x2<0.5)": 0.928 / (-0.894 * (v(2) < 0.5)),
    "sqrt(x0)": B.sqrt(v(0)),
    "sqrt(x2)": B.sqrt(v(2)),
    "sqrt(x3*1e-300)": B.sqrt(v(3) * 1e-300),
    "sqrt(0)": B.sqrt(v(1) - v(1)),
    "x0*x1+x2": v(0) * v(1) + v(2),
    "(x0/x1)/(c*(x0/x1))": (v(0) / v(1)) / (-1.85 * (v(0) / v(1))),
    "exp(x2)*sin(x3)": B.exp(v(2)) * B.sin(v(3)),
    "ln(x3)/cos(x2)": B.ln(v(3)) / B.cos(v(2)),
    "x2/1e-310": v(2) / 1e-310,
    "1e-310/x2": 1e-310 / v(2),
    "x2/1e305/1e10": (v(2) / 1e305) / 1e10,
    "sin(x2)": B.sin(v(2)), "cos(x2)": B.cos(v(2)), "sin(x0*700)": B.sin(v(0) * 700.0), "cos(x0*x1*9)": B.cos(v(0) * v(1) * 9.0),
    "sin(x2*1e-9)": B.sin(v(2) * 1e-9), "cos(x2*1e-9)": B.cos(v(2) * 1e-9), "sin(0)": B.sin(v(1) - v(1)),
    "sin(x0*1e5)": B.sin(v(0) * 1e5), "cos(x2/0)": B.cos(v(2) / (v(1) - v(1))),
    "exp(x2)": B.exp(v(2)), "exp(x0*7)": B.exp(v(0) * 7.0), "exp(-x0*7.3)": B.exp(v(0) * -7.3), "exp(x0*8)": B.exp(v(0) * 8.0),
    "exp(x2*1e-20)": B.exp(v(2) * 1e-20), "exp(ln(x3))": B.exp(B.ln(v(3))),
    "ln(x3)": B.ln(v(3)), "ln(x0)": B.ln(v(0)), "ln(x2)": B.ln(v(2)), "ln(x3*1e-310)": B.ln(v(3) * 1e-310), "ln(0)": B.ln(v(1) - v(1)),
    "ln(x3*1e300)": B.ln(v(3) * 1e300), "ln(exp(x2))": B.ln(B.exp(v(2))), "ln(1+x2*1e-12)": B.ln(1.0 + v(2) * 1e-12),
}
batch = B.Batch.from_exprs(B.MODE_EVAL_ONLY, [[e] for e in progs.values()])
out = {}
for s in (4, 1):
    os.environ["RR_B200_S"] = str(s)
    with Engine(X, y, flags=B.FLAG_FORCE_GRAM) as e:
        out[s] = np.array(e.score(batch).ssr, copy=True)
bad = 0
for i, name in enumerate(progs):
    a, b = out[4][i], out[1][i]
    same = (a == b) or (np.isnan(a) and np.isnan(b)) or (np.isfinite(a) and np.isfinite(b) and abs(a - b) <= 1e-11 * abs(b))
    bad += not same
    print(f"{'ok ' if same else 'BAD'} {name:24s} S4 {a!r:28} S1 {b!r}")
print("BAD", bad)
