"""Shared parity checker: engine results vs reference results for one rr_batch.

Tolerances (BASELINE.json north_star): fitness and OLS coefficients within 1e-9 relative in
fp64. What "relative" is measured against is written here once:

* coefficients: |c_gpu - c_ref| <= 1e-9 * max_j |c_ref_j| of the same candidate, checked for
  WELL-POSED candidates only: reference nonzero_pivots == k, finite fitness, and condition
  number of the column-scaled design matrix <= KAPPA_MAX (computed here by SVD from the
  oracle's own term values). For the rest the reference's answer is itself numerically
  arbitrary: they are counted and reported, and still must agree
  on the sentinel and, when both are finite, on fitness to a looser bound.
* fitness: f0 = 1-R2 and f1 = RMSE compared with an absolute floor of 1e-12 * (scale of y):
  near-perfect fits are rounding noise in the reference too.
"""
from __future__ import annotations

import os

import numpy as np

from rils_rols_b200 import batch as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REL = 1e-9
KAPPA_MAX = 1e6


def fitness_arrays(batch: B.Batch, res: B.Result, sst: float, n: int, coef=None):
    """Host side of fitness(): (f0, f1, size) per candidate from the engine's ssr + coefficients."""
    nc = batch.n_cand
    f0, f1, fs = np.zeros(nc), np.zeros(nc), np.zeros(nc, dtype=np.int64)
    tlen = np.diff(batch.term_code_begin)
    for c in range(nc):
        t0, t1 = int(batch.cand_term_begin[c]), int(batch.cand_term_begin[c + 1])
        if batch.mode == B.MODE_EVAL_ONLY:
            size = int(tlen[t0])
        else:
            cf = (coef if coef is not None else res.coef)[batch.coef_slice(c)]
            size, kept = 0, 0
            for j in range(t1 - t0):
                if abs(cf[j]) < 1e-12:
                    continue
                size += int(tlen[t0 + j]) + (0 if abs(cf[j] - 1) < 1e-12 else 2)
                kept += 1
            if not abs(cf[-1]) < 1e-12:
                size += 1
                kept += 1
            size = size + kept - 1 if kept else 1
        f0[c], f1[c], fs[c] = B.fitness_tuple(res.ssr[c], sst, n, size)
    return f0, f1, fs


def design_condition(Xfm, batch: B.Batch, c: int, evaluate) -> float:
    """2-norm condition number of the column-normalised design matrix of candidate c."""
    t0, t1 = int(batch.cand_term_begin[c]), int(batch.cand_term_begin[c + 1])
    cols = []
    for t in range(t0, t1):
        code = batch.code[batch.term_code_begin[t]:batch.term_code_begin[t + 1]]
        cols.append(evaluate(Xfm, code, batch.consts))
    cols.append(np.ones(Xfm.shape[1]))
    A = np.stack(cols, axis=1)
    if not np.all(np.isfinite(A)):
        return np.inf
    nrm = np.linalg.norm(A, axis=0)
    if np.any(nrm == 0):
        return np.inf
    s = np.linalg.svd(A / nrm, compute_uv=False)
    return float(s[0] / s[-1]) if s[-1] > 0 else np.inf


def compare(batch: B.Batch, res: B.Result, ref: dict, Xfm, y, sst: float, evaluate, label: str = "",
            check_nzp: bool = True):
    """ref: dict with ref_coef, ref_nonzero_pivots, ref_f0, ref_f1, ref_size (golden or oracle).
    Returns a report dict; raises AssertionError on a parity violation."""
    nc, n = batch.n_cand, Xfm.shape[1]
    f0, f1, fs = fitness_arrays(batch, res, sst, n)
    yscale = float(np.sqrt(sst / n)) if sst > 0 else 1.0
    rep = dict(label=label, n_cand=nc, well_posed=0, ambiguous=0, sentinel=0, max_coef_err=0.0, max_fit_err=0.0,
               loose_fail=0)
    ols = batch.mode == B.MODE_OLS_FIT
    for c in range(nc):
        ref_sent = ref["ref_size"][c] == 1000 and ref["ref_f0"][c] == 1000
        if ref_sent:
            rep["sentinel"] += 1
        well = True
        if ols:
            k = int(batch.cand_term_begin[c + 1] - batch.cand_term_begin[c]) + 1
            cr = ref["ref_coef"][batch.coef_slice(c)]
            well = (not ref_sent) and ref["ref_nonzero_pivots"][c] == min(k, n) and np.all(np.isfinite(cr))
            if well:
                well = design_condition(Xfm, batch, c, evaluate) <= KAPPA_MAX
        if not well and not ref_sent:
            rep["ambiguous"] += 1
            # both finite: same order of magnitude of fitness is all that can be asked
            continue
        if ref_sent:
            # NaN anywhere in the reference's prediction: the engine must report the sentinel too
            # (for rank-deficient designs the NaN can hinge on rounding noise: only well-defined
            # cases are asserted, i.e. a term itself is non-finite)
            if not ols or (res.flags[c] & B.RES_NONFINITE):
                assert fs[c] == 1000 and f0[c] == 1000, f"{label} cand {c}: reference sentinel, engine {f0[c]},{f1[c]},{fs[c]}"
            else:
                rep["ambiguous"] += 1
            continue
        rep["well_posed"] += 1
        snap_noise = False
        if ols:
            # a reference coefficient sitting in the rounding-noise band around a snap threshold
            # (|c| or |c-1| between 1e-14 and 1e-10; the thresholds are 1e-12, node.h:333-339) makes
            # the tree SIZE a coin flip in the reference itself: coefficients are still checked,
            # size / fitness are not
            for v in np.concatenate([np.abs(cr), np.abs(cr[:-1] - 1.0)]):
                if 1e-14 < v < 1e-10:
                    snap_noise = True
            cg = res.coef[batch.coef_slice(c)]
            scale = max(float(np.max(np.abs(cr))), 1e-300)
            err = float(np.max(np.abs(cg - cr))) / scale
            rep["max_coef_err"] = max(rep["max_coef_err"], err)
            assert err <= REL, f"{label} cand {c}: coefficient error {err:.3e} (gpu {cg}, ref {cr})"
            if check_nzp:
                assert res.nonzero_pivots[c] == ref["ref_nonzero_pivots"][c], f"{label} cand {c}: nonzero_pivots"
        if snap_noise:
            rep["snap_noise"] = rep.get("snap_noise", 0) + 1
            continue
        assert fs[c] == ref["ref_size"][c], f"{label} cand {c}: size {fs[c]} vs {ref['ref_size'][c]}"
        for name, g, r, floor in (("f0", f0[c], ref["ref_f0"][c], 1e-12), ("f1", f1[c], ref["ref_f1"][c], 1e-12 * yscale)):
            if np.isinf(r) or np.isinf(g):
                assert g == r, f"{label} cand {c}: {name} {g} vs {r}"
                continue
            err = abs(g - r) / (abs(r) + floor / REL)  # <= REL  <=>  |g - r| <= REL*|r| + floor
            rep["max_fit_err"] = max(rep["max_fit_err"], err)
            assert err <= REL, f"{label} cand {c}: {name} {g!r} vs {r!r} (err {err:.3e})"
    return rep
