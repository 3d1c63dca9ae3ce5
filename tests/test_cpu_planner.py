"""CPU suite, part 2: the host half of the engine (no GPU, no compute through the C ABI).

* librr_b200.so loads and exports every symbol include/rr_b200.h declares;
* creating an engine without a B200 fails loudly (there is no CPU fallback);
* the planner's instruction streams, executed by the numpy emulator in tests/isa_emu.py,
  reproduce the oracle: Gram / A^T y entries, EVAL_ONLY residual sums, materialised terms and
  residual passes, with and without cross-candidate sharing, chunked and tile-constrained.
"""
import os
import re

import numpy as np
import pytest

from oracle import pyoracle as O
from rils_rols_b200 import batch as B
from rils_rols_b200 import engine as E
from tests import isa_emu as EMU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rr_b200.h")).read()
    names = re.findall(r"RR_API\s+[\w\s\*]+?\b(rr_\w+)\s*\(", hdr)
    assert len(names) >= 15 and set(E.EXPORTS) <= set(names)
    L = E.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rr_b200.h but not exported"
    assert L.rr_abi_version() == E.ABI_VERSION == 3 and set(names) == set(E.EXPORTS)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(E.EngineError) as ei:
        E.Engine(np.ones((8, 2)), np.ones(8))
    assert "no CUDA device" in str(ei.value) or "CPU fallback" in str(ei.value)
    with pytest.raises(E.EngineError) as ei:
        E.Engine.sharded(np.ones((4096, 2)), np.ones(4096), n_gpus=2)
    assert "no CUDA device" in str(ei.value) or "CPU fallback" in str(ei.value)


def test_malformed_batches_are_rejected_on_the_host():
    """negative / decreasing offsets and offsets beyond the code length must come back as an error string
    from the planner (the C ABI turns it into RR_ERR_INVALID), never as an exception or an out-of-bounds read."""
    def plan_error(ctb, tcb, code, consts):
        b = B.Batch(B.MODE_OLS_FIT, np.asarray(ctb), np.asarray(tcb), np.asarray(code, dtype=np.uint32), np.asarray(consts, dtype=np.float64))
        try:
            EMU.Plan(b, 2, EMU.KIND_GRAM, tile_cols=20)
            return ""
        except ValueError as ex:
            return str(ex) or "error"

    assert plan_error([0, 1], [0, 1], [B.ins(B.OP_VAR, 0)], []) == ""
    for msg in (plan_error([0, -1], [0, 1], [B.ins(B.OP_VAR, 0)], []),          # negative term count
                plan_error([0, 2, 1], [0, 1, 2], [B.ins(B.OP_VAR, 0)] * 2, []),   # decreasing candidate offsets
                plan_error([0, 1], [0, 5], [B.ins(B.OP_VAR, 0)], []),             # term beyond the code array (n_code = 1)
                plan_error([0, 1], [0, 1], [B.ins(B.OP_CONST, 3)], [1.0])):       # constant index out of range
        assert msg, msg


def gram_from_dots(plan, dots, batch, c, n):
    m = int(batch.cand_term_begin[c + 1] - batch.cand_term_begin[c])
    idx = plan.tab[plan.tab_begin[c]:plan.tab_begin[c + 1]]
    assert idx.size == m * (m + 1) // 2 + 2 * m
    G = np.zeros((m + 1, m + 1))
    p = 0
    for i in range(m):
        for j in range(i, m):
            G[i, j] = G[j, i] = dots[idx[p]]
            p += 1
    byc = np.array([dots[idx[p + i]] for i in range(m)])
    p += m
    for i in range(m):
        G[i, m] = G[m, i] = dots[idx[p + i]]
    G[m, m] = n
    return G, byc


def design(Xfm, batch, c):
    cols = []
    for t in range(int(batch.cand_term_begin[c]), int(batch.cand_term_begin[c + 1])):
        cols.append(O.evaluate(Xfm, batch.code[batch.term_code_begin[t]:batch.term_code_begin[t + 1]], batch.consts))
    cols.append(np.ones(Xfm.shape[1]))
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("no_cse,target_chunks,tile_cols", [(False, 1, 56), (True, 1, 56), (False, 7, 56), (False, 1, 12)])
@pytest.mark.parametrize("cfg,prefix", [("cfg1_toy", "ls3_"), ("cfg3_breast_cancer", "ls0_")])
def test_gram_plan_reproduces_design_matrix_products(golden, cfg, prefix, no_cse, target_chunks, tile_cols):
    z = golden(cfg)
    X, y = z["X"], z["y"]
    batch = B.Batch.load_fields(z, prefix).subset(range(0, 300))
    if tile_cols == 12 and X.shape[1] > 4:
        tile_cols = 20  # each chunk stages only the columns it touches; 30 features do not fit at once
    plan = EMU.Plan(batch, X.shape[1], EMU.KIND_GRAM, tile_cols=tile_cols, target_chunks=target_chunks, no_cse=no_cse)
    assert plan.max_tile_cols <= tile_cols
    if target_chunks > 1:
        assert len(plan.chunks) > 1
    G_cols = EMU.engine_columns(X, y)
    dots, _ = EMU.run(plan, G_cols)
    Xfm = O.feature_major(X)
    yc = y - y.mean()
    with np.errstate(all="ignore"):
        for c in range(0, batch.n_cand, 7):
            A = design(Xfm, batch, c)
            G, byc = gram_from_dots(plan, dots, batch, c, X.shape[0])
            want = A.T @ A
            ok = np.isclose(G, want, rtol=1e-12, atol=0, equal_nan=True) | (~np.isfinite(want) & ~np.isfinite(G))
            assert ok.all(), f"cand {c}"
            wb = A[:, :-1].T @ yc
            okb = np.isclose(byc, wb, rtol=1e-10, atol=1e-9 * np.abs(yc).sum(), equal_nan=True) | (~np.isfinite(wb) & ~np.isfinite(byc))
            assert okb.all(), f"cand {c}"
    if not no_cse:
        # sharing: far fewer evaluations than term instances
        assert plan.n_terms_distinct < batch.n_terms
        if cfg == "cfg1_toy":
            assert plan.n_terms_distinct < 0.5 * batch.n_terms and plan.w_issued < plan.w_contract


def _opcodes(plan):
    return plan.ins["w0"] & 0xFF


@pytest.mark.parametrize("kind", [EMU.KIND_GRAM, EMU.KIND_EVAL])
@pytest.mark.parametrize("cfg,prefix", [("cfg1_toy", "ls3_"), ("cfg3_breast_cancer", "ls0_"), ("cfg5", "")])
def test_super_instructions_are_bit_identical_to_the_sequences_they_replace(golden, monkeypatch, cfg, prefix, kind):
    """rr_plan.cpp close(): the peephole pass (fused pin / constant / column forms, "X; MDOT" carriers), the
    planned ring rows and the RI_COMBINE points change the instruction stream, not one bit of any reduction.
    The emulator also checks every data slot and combine point against its own running counts."""
    if cfg == "cfg5":
        from rils_rols_b200 import workloads as W

        batch = W.cfg5_neighbourhood().subset(range(0, 600))
        X, y = W.cfg5_data(257)
    else:
        z = golden(cfg)
        X, y = z["X"], z["y"]
        batch = B.Batch.load_fields(z, prefix).subset(range(0, 300))
    if kind == EMU.KIND_EVAL:
        # EVAL_ONLY wants one program per candidate: score every term as its own candidate
        progs = [[(batch.code[batch.term_code_begin[t]:batch.term_code_begin[t + 1]], batch.consts)]
                 for t in range(min(batch.n_terms, 400))]
        batch = B.Batch.from_programs(B.MODE_EVAL_ONLY, progs)
    cols = EMU.engine_columns(X, y)
    fused = EMU.Plan(batch, X.shape[1], kind, tile_cols=40)
    monkeypatch.setenv("RR_B200_DEBUG_NO_FUSE", "1")
    monkeypatch.setenv("RR_B200_DEBUG_NO_MDROWS", "1")
    plain = EMU.Plan(batch, X.shape[1], kind, tile_cols=40)
    monkeypatch.delenv("RR_B200_DEBUG_NO_FUSE")
    monkeypatch.delenv("RR_B200_DEBUG_NO_MDROWS")
    assert fused.n_dots == plain.n_dots and np.array_equal(fused.tab, plain.tab)
    assert fused.has_rows and not plain.has_rows
    assert (_opcodes(plain) >= EMU.RI_MULP0).sum() - (_opcodes(plain) >= EMU.RI_FIRST_M).sum() == 0  # no fused register forms
    n_real = lambda p: int((_opcodes(p) != EMU.RI_NOP).sum())  # dispatches: data slots and padding are not dispatched
    if kind == EMU.KIND_GRAM:
        assert n_real(fused) < 0.8 * n_real(plain)
        assert (fused.ins["w0"] & EMU.RR_THEN_MDOT).astype(bool).sum() > 0
    with np.errstate(all="ignore"):
        d_fused, _ = EMU.run(fused, cols)
        d_plain, _ = EMU.run(plain, cols)
    assert np.array_equal(d_fused, d_plain, equal_nan=True)


@pytest.mark.parametrize("cfg", ["cfg1_toy", "cfg2_diabetes", "cfg3_breast_cancer"])
def test_eval_plan_reproduces_fitness(golden, cfg):
    z = golden(cfg)
    X, y = z["X"], z["y"]
    batch = B.Batch.load_fields(z, "pert0_")
    plan = EMU.Plan(batch, X.shape[1], EMU.KIND_EVAL, target_chunks=5)
    dots, _ = EMU.run(plan, EMU.engine_columns(X, y))
    sst = float(((y - y.mean()) ** 2).sum())
    n_ok = 0
    for c in range(batch.n_cand):
        size = int(batch.term_code_begin[c + 1] - batch.term_code_begin[c])
        f = B.fitness_tuple(dots[plan.tab[c]], sst, X.shape[0], size)
        ref = (z["pert0_ref_f0"][c], z["pert0_ref_f1"][c], int(z["pert0_ref_size"][c]))
        assert f[2] == ref[2]
        if np.isfinite(ref[0]):
            assert abs(f[0] - ref[0]) <= 1e-12 * max(1, abs(ref[0])) and abs(f[1] - ref[1]) <= 1e-12 * max(1, abs(ref[1]))
            n_ok += 1
        else:
            assert f[0] == ref[0]
    assert n_ok > 5


def test_materialise_plan_is_bit_exact_for_arithmetic_terms(golden):
    z = golden("cfg2_diabetes")
    X, y = z["X"], z["y"]
    batch = B.Batch.load_fields(z, "ls2_").subset(range(200))
    plan = EMU.Plan(batch, X.shape[1], EMU.KIND_MATERIALISE, target_chunks=3)
    _, V = EMU.run(plan, EMU.engine_columns(X, y), n_stg=plan.n_terms_distinct)
    Xfm = O.feature_major(X)
    transcend = {B.OP_SIN, B.OP_COS, B.OP_LN, B.OP_EXP, B.OP_POW}
    n_exact = 0
    for t in range(batch.n_terms):
        code = batch.code[batch.term_code_begin[t]:batch.term_code_begin[t + 1]]
        want = O.evaluate(Xfm, code, batch.consts)
        got = V[plan.term_ids[t]]
        if set((code & 0xFF).tolist()) & transcend:
            assert np.allclose(got, want, rtol=1e-14, atol=0, equal_nan=True)
        else:
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
            n_exact += 1
    assert n_exact > 50


@pytest.mark.parametrize("tile_cols", [56, 10])  # 10: wide candidates stream their terms
def test_residual_plan_matches_rebuilt_model(golden, tile_cols):
    z = golden("cfg1_toy")
    X, y = z["X"], z["y"]
    batch = B.Batch.load_fields(z, "ls1_").subset(range(120))
    Xfm = O.feature_major(X)
    ores, f0, f1, fs = O.score_batch(Xfm, y, batch)
    cs = ores.coef[: batch.n_coef].copy()
    cs[np.abs(cs) < 1e-12] = 0.0
    for c in range(batch.n_cand):  # value_one applies to term coefficients, not to the free term
        sl = batch.coef_slice(c)
        seg = cs[sl]
        seg[:-1][np.abs(seg[:-1] - 1.0) < 1e-12] = 1.0
    cs[~np.isfinite(cs)] = 0.0
    plan = EMU.Plan(batch, X.shape[1], EMU.KIND_RESIDUAL, coef=cs, tile_cols=tile_cols)
    assert plan.max_tile_cols <= tile_cols
    dots, _ = EMU.run(plan, EMU.engine_columns(X, y))
    n_ok = 0
    for c in range(batch.n_cand):
        if not np.all(np.isfinite(ores.coef[batch.coef_slice(c)])) or not np.isfinite(ores.ssr[c]):
            continue
        idx = plan.tab[plan.tab_begin[c]:plan.tab_begin[c + 1]]
        m = int(batch.cand_term_begin[c + 1] - batch.cand_term_begin[c])
        assert idx.size == m + 2
        assert abs(dots[idx[0]] - ores.ssr[c]) <= 1e-9 * max(ores.ssr[c], 1e-6), c
        # r.t_i and r.1 against the design matrix
        A = design(Xfm, batch, c)
        cz = cs[batch.coef_slice(c)]
        yh = np.zeros(X.shape[0])
        for j in range(m + 1):
            yh = yh + cz[j] * A[:, j]
        r = y - yh
        want = np.concatenate([A[:, :m].T @ r, [r.sum()]])
        got = np.array([dots[i] for i in idx[1:]])
        assert np.allclose(got, want, rtol=1e-9, atol=1e-7 * (np.abs(A).sum(axis=0) * np.abs(r).max()).max()), c
        n_ok += 1
    assert n_ok > 60


def _random_expr(rng, d, depth):
    """Random tree over all 19 opcodes (constants, variables, unary, binary, the rarely generated ones)."""
    if depth == 0 or rng.random() < 0.25:
        return B.Expr.var(int(rng.integers(d))) if rng.random() < 0.6 else B.Expr.const(float(np.round(rng.uniform(-3, 3), 3)))
    k = int(rng.integers(18))
    a = _random_expr(rng, d, depth - 1)
    if k < 6:
        return [B.sin, B.cos, B.ln, B.exp, B.sqrt, B.sqr][k](a)
    b = _random_expr(rng, d, depth - 1)
    if k < 10:
        return [lambda x, y: x + y, lambda x, y: x - y, lambda x, y: x * y, lambda x, y: x / y][k - 6](a, b)
    if k < 14:  # multiplications and divisions dominate real neighbourhoods: give the fused forms more chances
        return a * b if k % 2 else a / b
    return [B.min_, B.max_, B.eq, lambda x, y: x < y][k - 14](a, b)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("tile_cols", [12, 40])
def test_super_instructions_on_random_neighbourhoods(monkeypatch, seed, tile_cols):
    """Fuzz of the peephole pass: neighbourhoods of random candidates that share base terms (so that pins,
    cache registers, spills and every fused form come into play), planned with and without the pass, must
    reduce to the same bits; the fused plan must also agree with the design-matrix products."""
    rng = np.random.default_rng(seed)
    d, n = 5, 97
    X = rng.uniform(0.2, 2.5, (n, d))
    y = rng.normal(size=n)
    base = [_random_expr(rng, d, 3) for _ in range(5)]
    cands = []
    for _ in range(120):
        terms = list(base)
        j = int(rng.integers(len(terms)))
        r = rng.random()
        if r < 0.4:
            terms[j] = _random_expr(rng, d, 4)
        elif r < 0.7:
            terms[j] = terms[j] * B.Expr.var(int(rng.integers(d))) if rng.random() < 0.5 else terms[j] / B.Expr.var(int(rng.integers(d)))
        elif r < 0.85:
            terms[j] = B.Expr.const(float(np.round(rng.uniform(-2, 2), 2))) * terms[j]
        else:
            terms.append(_random_expr(rng, d, 2))
        cands.append(terms)
    batch = B.Batch.from_exprs(B.MODE_OLS_FIT, cands)
    cols = EMU.engine_columns(X, y)
    fused = EMU.Plan(batch, d, EMU.KIND_GRAM, tile_cols=tile_cols)
    monkeypatch.setenv("RR_B200_DEBUG_NO_FUSE", "1")
    plain = EMU.Plan(batch, d, EMU.KIND_GRAM, tile_cols=tile_cols)
    monkeypatch.delenv("RR_B200_DEBUG_NO_FUSE")
    assert int((_opcodes(fused) != EMU.RI_NOP).sum()) < int((_opcodes(plain) != EMU.RI_NOP).sum())
    with np.errstate(all="ignore"):
        d_fused, _ = EMU.run(fused, cols)
        d_plain, _ = EMU.run(plain, cols)
        assert np.array_equal(d_fused, d_plain, equal_nan=True)
        Xfm = O.feature_major(X)
        for c in range(0, batch.n_cand, 5):
            A = design(Xfm, batch, c)
            G, _ = gram_from_dots(fused, d_fused, batch, c, n)
            want = A.T @ A
            ok = np.isclose(G, want, rtol=1e-11, atol=0, equal_nan=True) | (~np.isfinite(want) & ~np.isfinite(G))
            assert ok.all(), f"cand {c}"


def test_halves_planned_concurrently_equal_halves_planned_in_turn():
    """rr_engine.cu run_gram plans the second half of a large neighbourhood on a helper thread: plan_gram must
    be re-entrant (it only reads the analysed batch)."""
    import ctypes as C

    from rils_rols_b200 import workloads as W

    L = E.lib()
    L.rr_debug_plan_concurrency_check.argtypes = [C.POINTER(B.rr_batch), C.c_int32, C.c_int32]
    L.rr_debug_plan_concurrency_check.restype = C.c_int
    nb = W.cfg5_neighbourhood()
    for lo, hi in ((0, 4096), (100, 1300), (2000, 2100)):
        sub = nb.subset(range(lo, hi))
        bs = sub.as_struct()
        for _ in range(3):
            assert L.rr_debug_plan_concurrency_check(C.byref(bs), 20, 23) == 0


def test_planner_rejects_malformed_batches():
    good = B.Batch.from_exprs(B.MODE_OLS_FIT, [[B.Expr.var(0) * B.Expr.var(1)]])
    EMU.Plan(good, 2, EMU.KIND_GRAM)
    with pytest.raises(ValueError):  # feature index out of range
        EMU.Plan(good, 1, EMU.KIND_GRAM)
    bad = B.Batch(B.MODE_OLS_FIT, [0, 1], [0, 2], np.array([B.ins(B.OP_VAR, 0), B.ins(B.OP_VAR, 0)], dtype=np.uint32), np.zeros(0))
    with pytest.raises(ValueError):  # two values left on the stack
        EMU.Plan(bad, 2, EMU.KIND_GRAM)
    bad2 = B.Batch(B.MODE_OLS_FIT, [0, 1], [0, 1], np.array([B.ins(B.OP_PLUS)], dtype=np.uint32), np.zeros(0))
    with pytest.raises(ValueError):  # stack underflow
        EMU.Plan(bad2, 2, EMU.KIND_GRAM)
    bad3 = B.Batch(B.MODE_OLS_FIT, [0, 1], [0, 1], np.array([B.ins(B.OP_CONST, 5)], dtype=np.uint32), np.zeros(2))
    with pytest.raises(ValueError):  # constant index out of range
        EMU.Plan(bad3, 2, EMU.KIND_GRAM)
    bad4 = B.Batch(B.MODE_EVAL_ONLY, [0, 2], [0, 1, 2], np.array([B.ins(B.OP_VAR, 0)] * 2, dtype=np.uint32), np.zeros(0))
    with pytest.raises(ValueError):  # EVAL_ONLY takes one program per candidate
        EMU.Plan(bad4, 2, EMU.KIND_EVAL)


def test_deep_trees_spill_into_slots():
    """A balanced tree needs spill temporaries (no dynamic stack in the kernel)."""
    v = B.Expr.var

    def bal(depth, k=0):
        if depth == 0:
            return B.sin(v(k % 3)), k + 1
        l, k = bal(depth - 1, k)
        r, k = bal(depth - 1, k)
        return (l * r if depth % 2 else l + r), k

    e, _ = bal(5)
    rng = np.random.default_rng(0)
    X = rng.uniform(0.5, 1.5, size=(40, 3))
    y = rng.normal(size=40)
    batch = B.Batch.from_exprs(B.MODE_EVAL_ONLY, [[e]])
    plan = EMU.Plan(batch, 3, EMU.KIND_EVAL)
    assert (plan.ins["w0"] & 0xFF == EMU.RI_ST).sum() >= 5
    dots, _ = EMU.run(plan, EMU.engine_columns(X, y))
    code, consts = e.program()
    want = float(((y - O.evaluate(O.feature_major(X), code, consts)) ** 2).sum())
    assert abs(dots[plan.tab[0]] - want) <= 1e-12 * want


def narrow_subset(batch, limit=None, max_terms=7):
    """candidates with at most 7 terms (what a G8 plan takes: a row meets its other terms, the centred target and the
    column of ones as the eight pins; the engine plans the wider ones the classic way); R8 plans take 8"""
    m = np.diff(batch.cand_term_begin)
    idx = [int(c) for c in np.nonzero(m <= max_terms)[0]]
    return idx[:limit] if limit else idx


@pytest.mark.parametrize("tile_cols", [26, 8, 5])
@pytest.mark.parametrize("cfg,prefix", [("cfg1_toy", "ls3_"), ("cfg3_breast_cancer", "ls0_"), ("cfg5", "")])
def test_g8_plan_reproduces_design_matrix_products(golden, cfg, prefix, tile_cols):
    """G8 plans (rr_isa.h RI_GRAM8: rows in tile slots reduced eight at a time against the pins): every Gram / A^T yc /
    column-sum entry the solver reads must be the design-matrix product, with plenty of tile slots and with very few
    (groups then close early), and the plan must be the classic plan's equal in shared work."""
    if cfg == "cfg5":
        from rils_rols_b200 import workloads as W

        full = W.cfg5_neighbourhood()
        X, y = W.cfg5_data(300)
    else:
        z = golden(cfg)
        X, y = z["X"], z["y"]
        full = B.Batch.load_fields(z, prefix)
    batch = full.subset(narrow_subset(full, 500))
    d = X.shape[1]
    tc = tile_cols + (d if tile_cols < 26 else 0)  # slots on top of the staged features
    if cfg == "cfg3_breast_cancer" and tile_cols == 26:
        tc = 40
    plan = EMU.Plan(batch, d, EMU.KIND_GRAM_G8, tile_cols=tc)
    classic = EMU.Plan(batch, d, EMU.KIND_GRAM, tile_cols=56)
    assert plan.max_tile_cols <= tc
    ops = _opcodes(plan)
    assert (ops == EMU.RI_GRAM8).any() and not (ops == EMU.RI_MDOT).any() and not (ops == EMU.RI_DOTM).any()
    assert plan.n_dots == classic.n_dots  # the same distinct reductions, each once
    G_cols = EMU.engine_columns(X, y)
    dots, _ = EMU.run(plan, G_cols)
    groups, rows = plan.gram_stats
    assert groups >= 1 and rows >= groups
    if tile_cols == 26 and cfg != "cfg1_toy":
        assert rows / groups > 3.0, (groups, rows)  # a local-search neighbourhood fills its groups
    Xfm = O.feature_major(X)
    yc = y - y.mean()
    with np.errstate(all="ignore"):
        for c in range(0, batch.n_cand, 5):
            A = design(Xfm, batch, c)
            G, byc = gram_from_dots(plan, dots, batch, c, X.shape[0])
            want = A.T @ A
            ok = np.isclose(G, want, rtol=1e-12, atol=0, equal_nan=True) | (~np.isfinite(want) & ~np.isfinite(G))
            assert ok.all(), f"cand {c}"
            wb = A[:, :-1].T @ yc
            okb = np.isclose(byc, wb, rtol=1e-10, atol=1e-9 * np.abs(yc).sum(), equal_nan=True) | (~np.isfinite(wb) & ~np.isfinite(byc))
            assert okb.all(), f"cand {c}"
    print(f"\n{cfg} tile_cols={tc}: {len(plan.ins)} instructions (classic {len(classic.ins)}), {groups} GRAM8 with {rows} rows "
          f"({rows / groups:.2f} per group), w_issued {plan.w_issued:.0f} (classic {classic.w_issued:.0f})")


@pytest.mark.parametrize("slots", [32, 10])
@pytest.mark.parametrize("cfg,prefix", [("cfg1_toy", "ls3_"), ("cfg3_breast_cancer", "ls0_"), ("cfg5", "")])
def test_r8_plan_reproduces_design_matrix_products(golden, cfg, prefix, slots):
    """R8 plans (rr_isa.h RQ_*: the row machine - eight rows of one shape per group, evaluated in the DMMA fragment
    layout and reduced straight from registers): every Gram / A^T yc / column-sum entry the solver reads must be the
    design-matrix product, with plenty of tile slots for the stored sub-expressions and with very few; the distinct
    reductions are the classic plan's."""
    if cfg == "cfg5":
        from rils_rols_b200 import workloads as W

        full = W.cfg5_neighbourhood()
        X, y = W.cfg5_data(300)
    else:
        z = golden(cfg)
        X, y = z["X"], z["y"]
        full = B.Batch.load_fields(z, prefix)
    batch = full.subset(narrow_subset(full, 500))
    d = X.shape[1]
    tc = slots + d  # the engine's tile has 52 columns (rr_sweep_r8.cuh)
    plan = EMU.Plan(batch, d, EMU.KIND_GRAM_R8, tile_cols=tc)
    classic = EMU.Plan(batch, d, EMU.KIND_GRAM, tile_cols=56)
    assert plan.max_tile_cols <= tc
    assert plan.n_dots == classic.n_dots
    G_cols = EMU.engine_columns(X, y)
    dots, st = EMU.run_r8(plan, G_cols)
    assert st["groups"] >= 1 and st["rows"] >= st["groups"]
    Xfm = O.feature_major(X)
    yc = y - y.mean()
    with np.errstate(all="ignore"):
        for c in range(0, batch.n_cand, 5):
            A = design(Xfm, batch, c)
            G, byc = gram_from_dots(plan, dots, batch, c, X.shape[0])
            want = A.T @ A
            ok = np.isclose(G, want, rtol=1e-12, atol=0, equal_nan=True) | (~np.isfinite(want) & ~np.isfinite(G))
            assert ok.all(), f"cand {c}"
            wb = A[:, :-1].T @ yc
            okb = np.isclose(byc, wb, rtol=1e-10, atol=1e-9 * np.abs(yc).sum(), equal_nan=True) | (~np.isfinite(wb) & ~np.isfinite(byc))
            assert okb.all(), f"cand {c}"
    print(f"\n{cfg} tile columns {tc}: {len(plan.ins)} instruction slots, {st['groups']} groups with {st['rows']} rows "
          f"({st['rows'] / st['groups']:.2f} per group), {st['stores']} stores, {st['pinb']} pin loads, "
          f"w_issued {plan.w_issued:.0f} (classic {classic.w_issued:.0f})")


def test_r8_plan_two_register_trees_and_forced_stores():
    """Trees whose two sides both need evaluating use the second register; when both sides need it themselves, one side
    becomes a stored sub-expression."""
    v = B.Expr.var
    a = (B.sin(v(0)) + v(1)) * (B.exp(v(1)) - v(2))                      # t / u
    b = ((B.sin(v(0)) + v(1)) * (B.cos(v(1)) - v(2))) / ((B.sqrt(v(2)) + 1.5) * (B.ln(v(0)) - v(1)))  # forced store
    c = 2.5 / (v(0) - B.sqrt(v(1) / v(2)))
    cands = [[a, v(0)], [b, v(1) * v(2)], [c, a], [b, c, a]]
    batch = B.Batch.from_exprs(B.MODE_OLS_FIT, cands)
    rng = np.random.default_rng(1)
    X = rng.uniform(0.5, 2.5, size=(300, 3))
    y = rng.normal(size=300)
    plan = EMU.Plan(batch, 3, EMU.KIND_GRAM_R8, tile_cols=10)
    ops = plan.ins["w0"] & 0xFF
    assert (ops == EMU.RQ_TU).any() and (ops == EMU.RQ_ST).any()
    dots, st = EMU.run_r8(plan, EMU.engine_columns(X, y))
    Xfm = O.feature_major(X)
    for ci in range(batch.n_cand):
        A = design(Xfm, batch, ci)
        G, byc = gram_from_dots(plan, dots, batch, ci, X.shape[0])
        assert np.allclose(G, A.T @ A, rtol=1e-12, atol=0)
        assert np.allclose(byc, A[:, :-1].T @ (y - y.mean()), rtol=1e-9, atol=1e-9 * np.abs(y).sum())


def test_g8_plan_rejects_wide_candidates():
    v = B.Expr.var
    wide = B.Batch.from_exprs(B.MODE_OLS_FIT, [[v(i % 3) * float(i + 2) + B.sin(v(i % 3) * float(i)) for i in range(9)]])
    with pytest.raises(ValueError, match="too wide"):
        EMU.Plan(wide, 3, EMU.KIND_GRAM_G8, tile_cols=26)


def test_terms_constant_by_construction_are_recognised():
    """The planner marks terms whose column is a multiple of the free term's by construction (rr_plan.h Term::exact_const);
    the Gram-path solver uses the marks to keep one of the two parallel columns by the reference's pivot rule instead of
    escalating a singular Gram matrix (GPU: test_constant_terms_follow_the_reference_pivot_rule_without_escalation)."""
    import ctypes as C

    v, k = B.Expr.var, B.Expr.const
    S = B.sin(k(1.0) / v(0))
    cases = [
        (B.sin(k(2.017)), True), (k(3.0) * B.exp(k(0.5)), True),            # no variable
        (S / S, True), (v(1) / v(1), True), (S - S, True),                    # t / t, t - t
        ((k(0.5) * S) / S, True), (S / (S * k(2.0)), True), ((k(0.5) * S) / (k(3.0) * S), True),
        (B.sqrt(v(2) - v(2)) / v(1), True), ((S - S) * v(1), True), (k(0.0) * v(1), True),  # zero columns
        (B.cos(v(2) - v(2)), True),                                            # cos(0) = 1
        (S / B.sin(k(1.0) / v(1)), False), (S * S, False), (v(0) / (v(0) + k(1.0)), False), (S + S, False),
        (B.cos(k(6.2e-05) * S), False), ((S / v(1)) * v(1), False),
    ]
    batch = B.Batch.from_exprs(B.MODE_OLS_FIT, [[v(0), e, v(1) * v(2)] for e, _ in cases])
    L = E.lib()
    L.rr_debug_const_terms.argtypes = [C.POINTER(B.rr_batch), C.c_int32, C.POINTER(C.c_uint32)]
    L.rr_debug_const_terms.restype = C.c_int
    out = (C.c_uint32 * batch.n_cand)()
    bs = batch.as_struct()
    assert L.rr_debug_const_terms(C.byref(bs), 3, out) == 0
    for c, (_, want) in enumerate(cases):
        assert out[c] == (2 if want else 0), (c, out[c], want)


def _design_numpy(Xfm, batch, c):
    """The design matrix of candidate c evaluated with numpy's own functions, operation by operation in tree order -
    the emulators' arithmetic (numpy's transcendentals differ from libm's by an ulp, which sin(exp(...)) of a random
    tree turns into percents)."""
    n = Xfm.shape[1]
    cols = []
    with np.errstate(all="ignore"):
        for t in range(int(batch.cand_term_begin[c]), int(batch.cand_term_begin[c + 1])):
            st = []
            for w in batch.code[batch.term_code_begin[t]:batch.term_code_begin[t + 1]].tolist():
                op, arg = w & 0xFF, w >> 8
                if op == B.OP_CONST: st.append(np.full(n, float(batch.consts[arg])))
                elif op == B.OP_VAR: st.append(Xfm[arg].copy())
                elif B.ARITY[op] == 1:
                    a = st.pop()
                    st.append({B.OP_SIN: np.sin, B.OP_COS: np.cos, B.OP_LN: np.log, B.OP_EXP: np.exp, B.OP_SQRT: np.sqrt,
                               B.OP_SQR: lambda v: v * v}[op](a))
                else:
                    r = st.pop()
                    l = st.pop()
                    if op == B.OP_PLUS: v = l + r
                    elif op == B.OP_MINUS: v = l - r
                    elif op == B.OP_MULTIPLY: v = l * r
                    elif op == B.OP_DIVIDE: v = l / r
                    elif op == B.OP_POW: v = np.power(l, r)
                    elif op == B.OP_LESS_THAN: v = (l < r).astype(float)
                    elif op == B.OP_GREATER_THAN: v = (l > r).astype(float)
                    elif op == B.OP_EQUAL: v = (l == r).astype(float)
                    elif op == B.OP_NOT_EQUAL: v = (l != r).astype(float)
                    elif op == B.OP_MIN: v = np.where(l < r, l, r)
                    else: v = np.where(l > r, l, r)
                    st.append(v)
            cols.append(st[-1])
    cols.append(np.ones(n))
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
@pytest.mark.parametrize("kind", ["r8", "g8"])
def test_row_and_g8_plans_on_random_neighbourhoods(seed, kind):
    """Fuzz of the two large-n Gram planners: neighbourhoods of random candidates over all 19 opcodes that share base
    terms (pins, stored sub-expressions, the second register, forced stores, segments, per-row constants, rare
    operators, constant terms all come into play). Every reduction the solver reads must be the design-matrix product."""
    rng = np.random.default_rng(100 + seed)
    d, n = 5, 200
    X = rng.uniform(0.2, 2.5, (n, d))
    y = rng.normal(size=n)
    base = [_random_expr(rng, d, 3) for _ in range(5)]
    cands = []
    for _ in range(150):
        terms = list(base)
        j = int(rng.integers(len(terms)))
        r = rng.random()
        if r < 0.35:
            terms[j] = _random_expr(rng, d, 5)  # deep trees: two registers, forced stores
        elif r < 0.65:
            v = B.Expr.var(int(rng.integers(d)))
            terms[j] = terms[j] * v if rng.random() < 0.5 else (terms[j] / v if rng.random() < 0.5 else v / terms[j])
        elif r < 0.8:
            terms[j] = B.Expr.const(float(np.round(rng.uniform(-2, 2), 2))) / terms[j]  # per-row constants
        elif r < 0.9:
            terms[j] = terms[j] / terms[j] if rng.random() < 0.5 else B.sin(B.Expr.const(float(np.round(rng.uniform(0.5, 3), 2))))
        else:
            terms.append(_random_expr(rng, d, 2))
        cands.append(terms)
    batch = B.Batch.from_exprs(B.MODE_OLS_FIT, cands)
    cols = EMU.engine_columns(X, y)
    if kind == "r8":
        plan = EMU.Plan(batch, d, EMU.KIND_GRAM_R8, tile_cols=d + 32)
        with np.errstate(all="ignore"):
            dots, st = EMU.run_r8(plan, cols)
        assert st["rows"] >= st["groups"] >= 1
    else:
        plan = EMU.Plan(batch, d, EMU.KIND_GRAM_G8, tile_cols=d + 7)
        with np.errstate(all="ignore"):
            dots, _ = EMU.run(plan, cols)
    Xfm = O.feature_major(X)
    yc = y - y.mean()
    with np.errstate(all="ignore"):
        for c in range(batch.n_cand):
            A = _design_numpy(Xfm, batch, c)
            G, byc = gram_from_dots(plan, dots, batch, c, n)
            want = A.T @ A
            scale = np.sqrt(np.abs(np.outer(np.diag(want), np.diag(want))))
            ok = np.isclose(G, want, rtol=1e-11, atol=0, equal_nan=True) | (np.abs(G - want) <= 1e-12 * scale) | \
                (~np.isfinite(want) & ~np.isfinite(G))
            assert ok.all(), f"cand {c}"
            wb = A[:, :-1].T @ yc
            okb = np.isclose(byc, wb, rtol=1e-9, atol=1e-11 * np.sqrt(np.abs(np.diag(want)[:-1]) * float(yc @ yc)), equal_nan=True) | \
                (~np.isfinite(wb) & ~np.isfinite(byc))
            assert okb.all(), f"cand {c}"
