"""CPU suite, part 4: the sample-sharded (multi-GPU) path on world_size-2 gloo.

The engine shards rows across ranks and exchanges exactly three things through its all-reduce
hook (rr_engine.cu): [sum y, n] -> global mean, [sum yc, yc.yc] -> sst, and per sweep the vector of
per-candidate partial reductions. This test runs that protocol on two CPU processes: each rank
executes the SAME planner output on its row shard (numpy ISA emulator), the partials are summed
with torch.distributed (gloo), and the solved coefficients / residuals must equal the oracle's on
the unsharded data."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import pyoracle as O
    from rils_rols_b200 import batch as B
    from tests import isa_emu as EMU

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", "cfg2_diabetes.npz"))
    X, y = z["X"], z["y"]
    n, d = X.shape
    lo, hi = rank * n // world, (rank + 1) * n // world
    Xs, ys = X[lo:hi], y[lo:hi]
    full = B.Batch.load_fields(z, "ls1_")
    batch = full.subset(range(0, full.n_cand, 3))

    # engine-create protocol: global mean, then centred sums
    t = torch.tensor([ys.sum(), float(len(ys))], dtype=torch.float64)
    dist.all_reduce(t)
    mean, n_total = float(t[0] / t[1]), int(t[1])
    yc = ys - mean
    t2 = torch.tensor([yc.sum(), float(yc @ yc)], dtype=torch.float64)
    dist.all_reduce(t2)
    sum_yc, sst = float(t2[0]), float(t2[1])

    # one Gram sweep on the shard, then the all-reduce of the partials
    plan = EMU.Plan(batch, d, EMU.KIND_GRAM)
    cols = np.vstack([np.ascontiguousarray(Xs.T), ys[None, :], yc[None, :]])
    dots, _ = EMU.run(plan, cols)
    td = torch.from_numpy(np.where(np.isfinite(dots), dots, 0.0).copy())
    bad = torch.from_numpy((~np.isfinite(dots)).astype(np.float64))
    dist.all_reduce(td)
    dist.all_reduce(bad)
    dots = td.numpy().copy()
    dots[bad.numpy() > 0] = np.nan

    if rank == 0:
        assert n_total == n and abs(mean - y.mean()) < 1e-12 * abs(y.mean())
        assert abs(sst - ((y - y.mean()) ** 2).sum()) < 1e-10 * sst and abs(sum_yc) < 1e-9 * n
        ores, f0, f1, fs = O.score_batch(O.feature_major(X), y, batch)
        n_ok = 0
        for c in range(batch.n_cand):
            m = int(batch.cand_term_begin[c + 1] - batch.cand_term_begin[c])
            idx = plan.tab[plan.tab_begin[c]:plan.tab_begin[c + 1]]
            G = np.zeros((m + 1, m + 1))
            p = 0
            for i in range(m):
                for j in range(i, m):
                    G[i, j] = G[j, i] = dots[idx[p]]
                    p += 1
            b = np.array([dots[idx[p + i]] for i in range(m)] + [sum_yc])
            p += m
            for i in range(m):
                G[i, m] = G[m, i] = dots[idx[p + i]]
            G[m, m] = n_total
            cr = ores.coef[batch.coef_slice(c)]
            if not np.all(np.isfinite(G)) or ores.nonzero_pivots[c] < m + 1 or not np.all(np.isfinite(cr)):
                continue
            if np.linalg.cond(G) > 1e10:
                continue
            rhs = b + mean * G[:, m]  # uncentred right-hand side, as rr_gram_solve forms it
            coef = np.linalg.solve(G, rhs)
            assert np.allclose(coef, cr, rtol=1e-7, atol=1e-7 * np.max(np.abs(cr))), (c, coef, cr)
            cz = coef.copy()
            cz[m] -= mean
            ssr = sst - 2 * cz @ b + cz @ G @ cz
            if abs(ores.ssr[c]) > 1e-6 * sst:
                assert abs(ssr - ores.ssr[c]) <= 1e-7 * ores.ssr[c], (c, ssr, ores.ssr[c])
            n_ok += 1
        with open(out_path, "w") as f:
            f.write(str(n_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_gram_matches_unsharded_oracle(tmp_path):
    import torch.multiprocessing as mp

    out = tmp_path / "ok.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert int(out.read_text()) >= 60
