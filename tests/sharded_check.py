"""Run under torchrun (one rank per GPU): sample-sharded engine vs the unsharded engine on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/sharded_check.py [hook]
Default: NCCL inside the engine (rr_engine_comm_init); `hook`: the Python all-reduce callback.
Prints SHARDED_OK on rank 0 when every candidate agrees to 1e-9."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from rils_rols_b200 import batch as B
    from rils_rols_b200 import workloads
    from rils_rols_b200.engine import Engine

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n = 1 << 19
    X, y = workloads.cfg5_data(n)
    batch = workloads.cfg5_neighbourhood().subset(range(0, 4096, 8))
    lo, hi = rank * n // world, (rank + 1) * n // world
    eng = Engine(np.ascontiguousarray(X[lo:hi]), np.ascontiguousarray(y[lo:hi]), device=local)
    if len(sys.argv) > 1 and sys.argv[1] == "hook":
        eng.set_allreduce_torch()
    else:
        eng.comm_init_torch()
    info = eng.info()
    assert info.n == hi - lo and info.n_total == n and info.world == world
    res = eng.score(batch)
    ev = B.Batch.from_exprs(B.MODE_EVAL_ONLY, [[B.sin(B.Expr.var(0)) * B.Expr.var(3)], [B.Expr.const(0.0)]])
    res_ev = eng.score(ev)
    ok = True
    if rank == 0:
        with Engine(X, y, device=local) as full:
            finfo = full.info()
            ref = full.score(batch)
            ref_ev = full.score(ev)
        assert abs(info.y_mean - finfo.y_mean) <= 1e-13 * abs(finfo.y_mean) and abs(info.sst - finfo.sst) <= 1e-12 * finfo.sst
        bad = 0
        for c in range(batch.n_cand):
            a, b = res.ssr[c], ref.ssr[c]
            if np.isnan(a) and np.isnan(b):
                continue
            if ref.flags[c] & (B.RES_RANKDEF | B.RES_DD):
                continue  # numerically arbitrary designs: sharding changes the rounding noise
            ca, cb = res.coef[batch.coef_slice(c)], ref.coef[batch.coef_slice(c)]
            if not (abs(a - b) <= 1e-9 * abs(b) and np.allclose(ca, cb, rtol=1e-8, atol=1e-9 * np.max(np.abs(cb)))):
                bad += 1
                print("MISMATCH", c, a, b, ca, cb)
        assert np.allclose(res_ev.ssr[:2], ref_ev.ssr[:2], rtol=1e-12)
        ok = bad == 0
        print("SHARDED_OK" if ok else f"SHARDED_FAIL {bad}", f"world={world} n_total={info.n_total} rows/rank={info.n}", flush=True)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
