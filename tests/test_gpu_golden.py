"""GPU parity against the committed golden fixtures (outputs of the unmodified reference,
tests/golden/make_golden.py) and against the C oracle run live, through the C ABI."""
import numpy as np
import pytest

from oracle import pyoracle as O
from rils_rols_b200 import batch as B
from rils_rols_b200.engine import Engine
from tests import parity

pytestmark = pytest.mark.gpu

CONFIGS = ["cfg1_toy", "cfg2_diabetes", "cfg3_breast_cancer"]
# observed per-class counts (all recorded neighbourhoods of a config, pert0 included; identical for the three engine
# configurations unless noted): see profiles/r2_parity_classes.jsonl
FLOORS = {
    # of 1013: 742 well-posed, 139 rank-deficient drops (all reproduced on the exact path: same nonzero_pivots, same zero
    # pattern, fitness to 1e-15), 91 ill-conditioned (Gram path: 4 of them at garbage level), 41 sentinels
    "cfg1_toy": dict(well_posed=742, arbitrary=4, sentinel_unconfirmed=0, rank_flip=0),
    # of 2248: 2015 / 158 / 42 / 33
    "cfg2_diabetes": dict(well_posed=2015, arbitrary=0, sentinel_unconfirmed=0, rank_flip=0),
    # of 6241: 5285 / 167 / 501 / 288
    "cfg3_breast_cancer": dict(well_posed=5285, arbitrary=0, sentinel_unconfirmed=0, rank_flip=0),
}


def neighbourhoods(z):
    yield "pert0", B.Batch.load_fields(z, "pert0_"), {k: z["pert0_" + k] for k in ("ref_coef", "ref_nonzero_pivots", "ref_f0", "ref_f1", "ref_size")}
    for i in range(int(z["n_ls"])):
        p = f"ls{i}_"
        yield f"ls{i}", B.Batch.load_fields(z, p), {k: z[p + k] for k in ("ref_coef", "ref_nonzero_pivots", "ref_f0", "ref_f1", "ref_size")}


@pytest.mark.parametrize("flags,name", [(0, "default"), (B.FLAG_FORCE_GRAM, "gram"), (B.FLAG_FORCE_EXACT | B.FLAG_NO_CSE, "exact-nocse")])
@pytest.mark.parametrize("cfg", CONFIGS)
def test_golden_neighbourhoods(golden, cfg, flags, name):
    z = golden(cfg)
    X, y = z["X"], z["y"]
    Xfm = O.feature_major(X)
    reports = []
    with Engine(X, y, flags=flags) as eng:
        info = eng.info()
        assert info.n == X.shape[0] and info.d == X.shape[1]
        # engine-level constants of R2(), rils_rols_cpp.cpp:41-43
        assert abs(info.y_mean - y.mean()) <= 1e-12 * max(1.0, abs(y.mean()))
        sst = float(((y - y.mean()) ** 2).sum())
        assert abs(info.sst - sst) <= 1e-12 * sst
        for label, batch, ref in neighbourhoods(z):
            res = eng.score(batch)
            # nonzero_pivots is only comparable where the engine ran the reference's own QR
            rep = parity.compare(batch, res, ref, Xfm, y, info.sst, O.evaluate, f"{cfg}/{name}/{label}",
                                 check_nzp=bool(flags != B.FLAG_FORCE_GRAM))
            reports.append(rep)
    tot = parity.summary(reports)
    print(f"\n{cfg}/{name}: {tot}")
    parity.record(f"golden/{cfg}/{name}", tot)
    # floors = observed on B200 (profiles/r2_parity_classes.jsonl) minus 1 % of the candidates
    floor = FLOORS[cfg]
    assert tot["well_posed"] >= floor["well_posed"] - 0.01 * tot["n_cand"], tot
    assert tot["arbitrary"] <= floor["arbitrary"] + 0.01 * tot["n_cand"], tot
    assert tot["sentinel_unconfirmed"] <= floor["sentinel_unconfirmed"] + 0.01 * tot["n_cand"], tot
    if flags != B.FLAG_FORCE_GRAM:
        assert tot["rank_flip"] <= floor["rank_flip"] + 0.01 * tot["n_cand"], tot


@pytest.mark.parametrize("cfg", CONFIGS)
def test_exact_path_matches_oracle_bitwise_on_arithmetic_terms(golden, cfg):
    """Candidates whose terms use only + - * / sqrt and comparisons: the exact path runs the same
    operations in the same order as the oracle, so coefficients must agree to the last bit."""
    z = golden(cfg)
    X, y = z["X"], z["y"]
    Xfm = O.feature_major(X)
    transcend = {B.OP_SIN, B.OP_COS, B.OP_LN, B.OP_EXP, B.OP_POW}
    with Engine(X, y, flags=B.FLAG_FORCE_EXACT) as eng:
        for label, batch, ref in neighbourhoods(z):
            if batch.mode != B.MODE_OLS_FIT:
                continue
            res = eng.score(batch)
            ores, _, _, _ = O.score_batch(Xfm, y, batch)
            n_checked = 0
            for c in range(batch.n_cand):
                t0, t1 = batch.cand_term_begin[c], batch.cand_term_begin[c + 1]
                ops = set((batch.code[batch.term_code_begin[t0]:batch.term_code_begin[t1]] & 0xFF).tolist())
                if ops & transcend:
                    continue
                sl = batch.coef_slice(c)
                a, b = res.coef[sl], ores.coef[sl]
                assert np.array_equal(a.view(np.uint64), b.view(np.uint64)) or (np.isnan(a).all() and np.isnan(b).all()), \
                    f"{cfg}/{label} cand {c}: {a} vs {b}"
                assert res.nonzero_pivots[c] == ores.nonzero_pivots[c]
                n_checked += 1
            assert n_checked > 0
