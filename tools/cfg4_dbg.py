import sys, os, numpy as np
sys.path.insert(0, "/root/repo")
from oracle import pyoracle as O
from rils_rols_b200 import batch as B, workloads as W
from rils_rols_b200.engine import Engine
z = np.load("tests/golden/cfg4_large.npz")
b = B.Batch.load_fields(z, "ls0_")
X, y = W.cfg4_data(1_000_000, 10)
sub = b.subset([28, 15, 3])
v = B.Expr.var
alone = B.Batch.from_exprs(B.MODE_OLS_FIT, [[v(7), v(9)]])
for n in (1_000_000, 999_936, 2**19):
    with Engine(X[:n], y[:n]) as eng:
        r = eng.score(b); ra = eng.score(alone); rs = eng.score(sub)
        st = eng.stats()
        ores, *_ = O.score_batch(O.feature_major(X[:n]), y[:n], alone)
        print(os.environ.get("RR_B200_G8", "1"), n, "in batch", r.coef[b.coef_slice(28)], hex(r.flags[28]), "alone", ra.coef[:3], hex(ra.flags[0]), "sub", rs.coef[:3], "oracle", ores.coef[:3])
