#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/ from the UNMODIFIED reference
(oracle/_ref, built by `make -C oracle ref`; needs /root/reference, so this runs in the
build container only — the fixtures are committed, this script documents how).

The reference ships no golden vectors for the scoring path, so the
pins are outputs of the reference itself: for recorded neighbourhoods of the BASELINE.json
configs, the rr_batch arrays (term bytecode as the reference's own factor selection
produced it) plus what the reference computed per candidate — QR coefficients,
nonzero_pivots, and the fitness tuple (1-R2, RMSE, size) of the tuned tree.

    python tests/golden/make_golden.py            # configs 1-3, the d=20 neighbourhood, config 4 (10^6 rows: minutes)
    python tests/golden/make_golden.py --cfg5-n 16777216   # tune the cfg-5 base on the full 2^24 rows
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402
from rils_rols_b200 import batch as B  # noqa: E402
from rils_rols_b200.workloads import cfg4_data, cfg4_truth_expr, cfg5_data, cfg5_base_expr, config_data  # noqa: E402


def to_batch(g) -> B.Batch:
    return B.Batch(g["mode"], g["cand_term_begin"], g["term_code_begin"], g["code"], g["consts"])


def record(h, trees, ols_fit):
    g = h.score_list([(c[0], c[1]) for c in trees], ols_fit)
    out = to_batch(g).save_fields()
    for k in ("ref_coef", "ref_nonzero_pivots", "ref_f0", "ref_f1", "ref_size"):
        out[k] = g[k]
    return out, g


def small_config(R, name, X, y, classification, max_complexity, seed, n_ls=3, max_cands=1500):
    """Replays the first moves of fit_inner (rils_rols_cpp.cpp:799-842) and records the
    perturbation neighbourhood of the start solution and a few LS neighbourhoods."""
    h = R.RefHarness(classification, 0.001, max_complexity, seed)
    h.set_data(X, y)
    zero = B.Expr.const(0.0).program()
    fx = {"X": X, "y": y, "classification": np.int32(classification)}
    perts = h.all_candidates(zero[0], zero[1], False)
    rec, g = record(h, perts, False)
    fx.update({f"pert0_{k}": v for k, v in rec.items()})
    order = np.argsort(g["ref_f0"], kind="stable")
    n_nb = 0
    cur = None
    for rank in range(n_ls):
        p = perts[int(order[rank])]
        cur = h.tune(p[0], p[1], False)
        ls = h.all_candidates(cur["tuned_code"], cur["tuned_consts"], True)[:max_cands]
        rec, g = record(h, ls, True)
        fx.update({f"ls{n_nb}_{k}": v for k, v in rec.items()})
        n_nb += 1
    # two greedy LS steps from the best perturbation to reach a deeper tree
    p = perts[int(order[0])]
    cur = h.tune(p[0], p[1], False)
    for step in range(2):
        ls = h.all_candidates(cur["tuned_code"], cur["tuned_consts"], True)
        g = h.score_list([(c[0], c[1]) for c in ls], True)
        fit = (1 + g["ref_f0"]) * (1 + g["ref_f1"]) * (1 + g["ref_size"] * 0.001)
        best = int(np.argmin(np.where(g["ref_size"] <= max_complexity, fit, np.inf)))
        cur = h.tune(ls[best][0], ls[best][1], False)
    ls = h.all_candidates(cur["tuned_code"], cur["tuned_consts"], True)[:max_cands]
    rec, g = record(h, ls, True)
    fx.update({f"ls{n_nb}_{k}": v for k, v in rec.items()})
    n_nb += 1
    fx["n_ls"] = np.int32(n_nb)
    fx["deep_base_str"] = np.array(cur["tuned_str"])
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **fx)
    print(f"{name}: n={X.shape[0]} d={X.shape[1]} perts={len(perts)} ls_neighbourhoods={n_nb} "
          f"deep base={cur['tuned_str']} -> {os.path.getsize(path) / 1e3:.0f} kB")


def cfg5(R, n_tune, n_score=4096, n_cand=4096, max_term_nodes=50):
    """first 4096 trees (<= 50 term nodes) of
    all_candidates(tuned_base, local_search=true) on the d=20 synthetic set."""
    t0 = time.time()
    X, y = cfg5_data(n_tune)
    h = R.RefHarness(False, 0.001, 50, 12345)
    h.set_data(X, y)
    base = cfg5_base_expr().program()
    tuned = h.tune(base[0], base[1], False)
    print(f"cfg5 base tuned on n={n_tune}: {tuned['tuned_str']} fitness={tuned['fitness']} ({time.time() - t0:.1f}s)")
    cands = h.all_candidates(tuned["tuned_code"], tuned["tuned_consts"], True)
    print(f"cfg5 full neighbourhood: {len(cands)} trees")
    # term bytecode needs tune(); do it on a small prefix of the data (term lists do not depend on n)
    hs = R.RefHarness(False, 0.001, 50, 12345)
    hs.set_data(X[:n_score], y[:n_score])
    keep, skipped = [], 0
    for c in cands:
        if len(keep) == n_cand:
            break
        g = hs.score_list([(c[0], c[1])], True)
        if int(g["term_code_begin"][-1]) > max_term_nodes:
            skipped += 1
            continue
        keep.append(c)
    rec, g = record(hs, keep, True)
    b = to_batch(g)
    k = np.diff(b.cand_term_begin) + 1
    print(f"cfg5 kept {len(keep)} (skipped {skipped} with > {max_term_nodes} term nodes); mean k={k.mean():.2f} "
          f"mean term nodes={b.code.size / len(keep):.1f} mean W={b.contract_work().mean():.1f} "
          f"sentinels={(g['ref_size'] == 1000).sum()} rankdef={(g['ref_nonzero_pivots'] < k).sum()}")
    fx = dict(rec)
    fx.update(n_tune=np.int64(n_tune), n_score=np.int32(n_score), skipped=np.int32(skipped),
              full_neighbourhood=np.int32(len(cands)), tuned_base_str=np.array(tuned["tuned_str"]),
              tuned_code=tuned["tuned_code"], tuned_consts=tuned["tuned_consts"])
    path = os.path.join(HERE, "cfg5_neighbourhood.npz")
    np.savez_compressed(path, **fx)
    print(f"cfg5 -> {os.path.getsize(path) / 1e3:.0f} kB")


def cfg4(R, n=1_000_000, d=10, max_cands_first=100000, max_cands_truth=320):
    """BASELINE config 4 (test_large.py style): 10^6 x 10 synthetic, y = sum_{i<5} sin(1/x_i),
    NOISE-FREE - the near-perfect-fit regime of SURVEY.md 7.2-3. Two neighbourhoods scored by the unmodified
    reference on all 10^6 rows: (ls0) the first local-search neighbourhood fit() reaches (fit_inner :799-842:
    perturbations of const 0, best by 1-R2, tune_constants, all_candidates(local_search=true)), and (ls1) a prefix of
    the neighbourhood of the ground truth itself, where most candidates fit to rounding level."""
    t0 = time.time()
    X, y = cfg4_data(n, d)
    h = R.RefHarness(False, 0.001, 50, 12345)
    h.set_data(X, y)
    zero = B.Expr.const(0.0).program()
    perts = h.all_candidates(zero[0], zero[1], False)
    rec, g = record(h, perts, False)
    fx = {"n": np.int64(n), "d": np.int32(d), "x_checksum": np.float64(X.sum()), "y_checksum": np.float64(y.sum())}
    fx.update({f"pert0_{k}": v for k, v in rec.items()})
    print(f"cfg4: {len(perts)} perturbations scored ({time.time() - t0:.1f}s)")
    order = np.argsort(g["ref_f0"], kind="stable")
    p = perts[int(order[0])]
    cur = h.tune(p[0], p[1], False)
    ls = h.all_candidates(cur["tuned_code"], cur["tuned_consts"], True)[:max_cands_first]
    rec, g = record(h, ls, True)
    fx.update({f"ls0_{k}": v for k, v in rec.items()})
    fx["ls0_base_str"] = np.array(cur["tuned_str"])
    print(f"cfg4 ls0: base {cur['tuned_str']}, {len(ls)} candidates ({time.time() - t0:.1f}s)")
    truth = cfg4_truth_expr().program()
    cur = h.tune(truth[0], truth[1], False)
    print(f"cfg4 truth tuned: {cur['tuned_str']} fitness={cur['fitness']}")
    ls = h.all_candidates(cur["tuned_code"], cur["tuned_consts"], True)
    # every 7th tree: a spread over all generator kinds instead of the first node's candidates only
    ls = ls[::max(1, len(ls) // max_cands_truth)][:max_cands_truth]
    rec, g = record(h, ls, True)
    fx.update({f"ls1_{k}": v for k, v in rec.items()})
    fx["ls1_base_str"] = np.array(cur["tuned_str"])
    fx["n_ls"] = np.int32(2)
    k = np.diff(g["cand_term_begin"]) + 1
    print(f"cfg4 ls1: {len(ls)} candidates, near-perfect (f0 < 1e-20): {(g['ref_f0'] < 1e-20).sum()}, "
          f"rankdef {(g['ref_nonzero_pivots'] < k).sum()}, sentinels {(g['ref_size'] == 1000).sum()} ({time.time() - t0:.1f}s)")
    path = os.path.join(HERE, "cfg4_large.npz")
    np.savez_compressed(path, **fx)
    print(f"cfg4 -> {os.path.getsize(path) / 1e3:.0f} kB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg5-n", type=int, default=1 << 24)
    ap.add_argument("--skip-cfg5", action="store_true")
    ap.add_argument("--skip-small", action="store_true")
    ap.add_argument("--skip-cfg4", action="store_true")
    args = ap.parse_args()
    R = O.load_ref()
    if R is None:
        raise SystemExit("oracle/_ref is not built: make -C oracle ref")
    if not args.skip_small:
        for name, cls, mc in (("cfg1_toy", False, 50), ("cfg2_diabetes", False, 20), ("cfg3_breast_cancer", True, 20)):
            X, y = config_data(name)
            small_config(R, name, X, y, cls, mc, 12345)
    if not args.skip_cfg5:
        cfg5(R, args.cfg5_n)
    if not args.skip_cfg4:
        cfg4(R)


if __name__ == "__main__":
    main()
