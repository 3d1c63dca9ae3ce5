"""CPU suite, part 1: the oracle itself.

* the plain-C restatement (oracle/rr_oracle.c) against the committed golden fixtures, which are
  outputs of the UNMODIFIED reference (tests/golden/make_golden.py): bit-for-bit;
* against the reference run live, where oracle/_ref could be built (i.e. where /root/reference
  exists: the build container; skipped on the GPU box);
* the edge cases of node::evaluate_inner / fitness() the kernels must reproduce.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from rils_rols_b200 import batch as B
from rils_rols_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = ["cfg1_toy", "cfg2_diabetes", "cfg3_breast_cancer"]
REF_KEYS = ("ref_coef", "ref_nonzero_pivots", "ref_f0", "ref_f1", "ref_size")


def same_bits(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def prefixes(z):
    out = ["pert0_"]
    out += [f"ls{i}_" for i in range(int(z["n_ls"]))]
    return out


@pytest.mark.parametrize("cfg", CONFIGS)
def test_c_oracle_reproduces_reference_goldens_bitwise(golden, cfg):
    z = golden(cfg)
    Xfm = O.feature_major(z["X"])
    n_checked = 0
    for p in prefixes(z):
        batch = B.Batch.load_fields(z, p)
        res, f0, f1, fs = O.score_batch(Xfm, z["y"], batch)
        assert same_bits(f0, z[p + "ref_f0"]), p
        assert same_bits(f1, z[p + "ref_f1"]), p
        assert np.array_equal(fs, z[p + "ref_size"]), p
        if batch.mode == B.MODE_OLS_FIT:
            assert same_bits(res.coef[: batch.n_coef], z[p + "ref_coef"]), p
            assert np.array_equal(res.nonzero_pivots[: batch.n_cand], z[p + "ref_nonzero_pivots"]), p
        n_checked += batch.n_cand
    assert n_checked > 500


def test_c_oracle_reproduces_cfg5_golden_bitwise(golden):
    z = golden("cfg5_neighbourhood")
    n = int(z["n_score"])
    X, y = workloads.cfg5_data(n)
    batch = B.Batch.load_fields(z)
    assert batch.n_cand == 4096
    sub = batch.subset(range(0, 4096, 16))  # 256 candidates keep the CPU suite short
    idx = np.arange(0, 4096, 16)
    res, f0, f1, fs = O.score_batch(O.feature_major(X), y, sub)
    assert same_bits(f0, z["ref_f0"][idx]) and same_bits(f1, z["ref_f1"][idx])
    assert np.array_equal(fs, z["ref_size"][idx])
    assert np.array_equal(res.nonzero_pivots[: sub.n_cand], z["ref_nonzero_pivots"][idx])
    ref_coef = np.concatenate([z["ref_coef"][batch.coef_slice(int(c))] for c in idx])
    assert same_bits(res.coef[: sub.n_coef], ref_coef)
    # the workload statistics SURVEY.md 8(d) quotes for this neighbourhood
    k = np.diff(batch.cand_term_begin) + 1
    assert 6.0 < k.mean() < 7.0 and batch.code.size / 4096 <= 50
    assert 150 < batch.contract_work().mean() < 200


def test_c_oracle_matches_live_reference():
    R = O.load_ref()
    if R is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    X, y = workloads.config_data("cfg2_diabetes")
    h = R.RefHarness(False, 0.001, 20, 12345)
    h.set_data(X, y)
    v = B.Expr.var
    base = (v(2) * v(3) + B.sin(v(8)) + v(6) / v(1) + B.exp(v(2) - v(1))).program()
    tuned = h.tune(base[0], base[1], True)
    # the factor list of tune_constants(), evaluated by the oracle, must be the matrix the QR saw
    Xfm = O.feature_major(X)
    for j, (code, consts) in enumerate(zip(tuned["term_code"], tuned["term_consts"])):
        assert same_bits(O.evaluate(Xfm, code, consts), tuned["A_cols"][j])
    assert same_bits(tuned["A_cols"][-1], np.ones(X.shape[0]))
    cands = h.all_candidates(tuned["tuned_code"], tuned["tuned_consts"], True)[:400]
    g = h.score_list([(c[0], c[1]) for c in cands], True)
    batch = B.Batch(g["mode"], g["cand_term_begin"], g["term_code_begin"], g["code"], g["consts"])
    res, f0, f1, fs = O.score_batch(Xfm, y, batch)
    assert same_bits(res.coef[: batch.n_coef], g["ref_coef"])
    assert np.array_equal(res.nonzero_pivots[: batch.n_cand], g["ref_nonzero_pivots"])
    assert same_bits(f0, g["ref_f0"]) and same_bits(f1, g["ref_f1"]) and np.array_equal(fs, g["ref_size"])
    # perturbation-style scoring (fitness() without tune_constants), rils_rols_cpp.cpp:828
    g2 = h.score_list([(c[0], c[1]) for c in cands[:200]], False)
    b2 = B.Batch(g2["mode"], g2["cand_term_begin"], g2["term_code_begin"], g2["code"], g2["consts"])
    _, e0, e1, es = O.score_batch(Xfm, y, b2)
    assert same_bits(e0, g2["ref_f0"]) and same_bits(e1, g2["ref_f1"]) and np.array_equal(es, g2["ref_size"])


def test_all_opcodes_follow_evaluate_inner():
    """node.cpp:23-95 semantics incl. the asymmetric NaN handling of MIN/MAX and exact compares."""
    rng = np.random.default_rng(7)
    X = rng.normal(size=(64, 3))
    X[0, 0] = np.nan
    X[1, 1] = np.nan
    X[2, 0] = X[2, 1] = 0.5
    Xfm = O.feature_major(X)
    a, b = Xfm[0], Xfm[1]
    v = B.Expr.var
    with np.errstate(all="ignore"):
        cases = [
            (v(0) + v(1), a + b), (v(0) - v(1), a - b), (v(0) * v(1), a * b), (v(0) / v(1), a / b),
            (B.sin(v(0)), np.sin(a)), (B.cos(v(0)), np.cos(a)), (B.ln(v(0)), np.log(a)), (B.exp(v(0)), np.exp(a)),
            (B.sqrt(v(0)), np.sqrt(a)), (B.sqr(v(0)), a * a), (B.pow_(v(0), v(1)), np.power(a, b)),
            (v(0) < v(1), (a < b).astype(float)), (v(0) > v(1), (a > b).astype(float)),
            (B.eq(v(0), v(1)), (a == b).astype(float)), (B.ne(v(0), v(1)), (a != b).astype(float)),
            (B.min_(v(0), v(1)), np.where(a < b, a, b)), (B.max_(v(0), v(1)), np.where(a > b, a, b)),
            (B.Expr.const(2.5), np.full(64, 2.5)),
        ]
    for e, want in cases:
        code, consts = e.program()
        got = O.evaluate(Xfm, code, consts)
        ok = np.isclose(got, want, rtol=1e-15, atol=0, equal_nan=True) | (np.isinf(got) & (got == want))
        assert ok.all(), B.OP_NAMES[e.op]
    # a<b ? a : b with NaN in a yields b; NaN in b yields b as well (node.cpp:82)
    code, consts = B.min_(v(0), v(1)).program()
    got = O.evaluate(Xfm, code, consts)
    assert got[0] == b[0] and np.isnan(got[1])


def test_fitness_sentinel_and_infinity():
    """NaN anywhere -> (1000, 1000, 1000); +-inf is not mapped (rils_rols_cpp.cpp:529-530)."""
    n = 50
    X = np.linspace(-1.0, 1.0, n).reshape(n, 1)
    y = np.linspace(0.0, 3.0, n)
    Xfm = O.feature_major(X)
    v = B.Expr.var
    batch = B.Batch.from_exprs(B.MODE_EVAL_ONLY, [[B.ln(v(0))], [B.exp(v(0) * 1000.0)], [v(0) * 2.0]])
    _, f0, f1, fs = O.score_batch(Xfm, y, batch)
    assert (f0[0], f1[0], fs[0]) == (1000.0, 1000.0, 1000)
    assert np.isinf(f0[1]) and np.isinf(f1[1]) and fs[1] == 4
    assert np.isfinite(f0[2]) and fs[2] == 3
    # host-side fitness tuple helper == the oracle's
    sst = float(((y - y.mean()) ** 2).sum())
    ssr = float(((y - X[:, 0] * 2.0) ** 2).sum())
    t = B.fitness_tuple(ssr, sst, n, 3)
    assert abs(t[0] - f0[2]) < 1e-14 * f0[2] and abs(t[1] - f1[2]) < 1e-14 * f1[2]  # numpy sums pairwise


def test_ols_snapping_and_size_rules():
    """rils_rols_cpp.cpp:488-517: |c| < 1e-12 drops the term, |c-1| < 1e-12 drops the multiplier."""
    rng = np.random.default_rng(3)
    n = 200
    X = rng.uniform(1, 2, size=(n, 3))
    y = X[:, 0] + 2.5 * np.sin(X[:, 1])  # exact: coefficient 1, coefficient 2.5, no x2, free term 0
    v = B.Expr.var
    batch = B.Batch.from_exprs(B.MODE_OLS_FIT, [[v(0), B.sin(v(1)), v(2)]])
    res, f0, f1, fs = O.score_batch(O.feature_major(X), y, batch)
    c = res.coef[:4]
    assert abs(c[0] - 1) < 1e-12 and abs(c[1] - 2.5) < 1e-9 and abs(c[2]) < 1e-12 and abs(c[3]) < 1e-12
    # x0 (1 node) + 2.5*sin(x1) (2 + 2 nodes) + one PLUS
    assert fs[0] == 1 + 4 + 1
    assert f1[0] < 1e-12
    # empty candidate: only the free term -> constant mean(y), size 1
    b2 = B.Batch(B.MODE_OLS_FIT, [0, 0], [0], np.zeros(0, dtype=np.uint32), np.zeros(0))
    r2, g0, g1, gs = O.score_batch(O.feature_major(X), y, b2)
    assert abs(r2.coef[0] - y.mean()) < 1e-12 and gs[0] == 1 and abs(g0[0] - 1.0) < 1e-12


def test_fast_transcendentals_within_one_ulp(tmp_path):
    """oracle/rr_fastmath_check.c restates the PTX fast paths of sin/cos/exp/log operation by operation and
    measures them against glibc: every result within 1 ulp of the correctly rounded value."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = os.path.join(ROOT, "oracle", "rr_fastmath_check.c")
    exe = str(tmp_path / "fm")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
