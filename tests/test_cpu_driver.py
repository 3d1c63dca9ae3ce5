"""CPU suite, part 3: the host ILS driver (no GPU): candidate generation, string forms, factor
selection and tree rebuild against the unmodified reference (where oracle/_ref is available) and
against the committed cfg5 fixture (always)."""
import os
import sys

import numpy as np
import pytest

from oracle import pyoracle as O
from rils_rols_b200 import batch as B

import rils_rols_b200  # noqa: E402

M = rils_rols_b200.driver_module()


def expr_from_postfix(code, consts) -> B.Expr:
    st = []
    for w in np.asarray(code).tolist():
        op, arg = w & 0xFF, w >> 8
        if op == B.OP_CONST:
            st.append(B.Expr.const(float(consts[arg])))
        elif op == B.OP_VAR:
            st.append(B.Expr.var(arg))
        elif B.ARITY[op] == 1:
            st.append(B.Expr(op, st.pop()))
        else:
            r = st.pop()
            st.append(B.Expr(op, st.pop(), r))
    return st[0]


def test_boundary_signature_matches_reference():
    """rils_rols_cpp.cpp:998-1007: class rils_rols with the 8-argument constructor and 7 methods."""
    rr = M.rils_rols(False, 1000, 10, 0.001, 50, 1.0, False, 12345)
    for name in ("fit", "predict", "get_model_string", "get_best_time", "get_fit_calls", "get_total_time"):
        assert callable(getattr(rr, name))
    with pytest.raises(ValueError, match="Size of X"):
        rr.fit(np.zeros((10, 1)), np.zeros(5), 5, 3)
    with pytest.raises(ValueError, match="Size of y"):
        rr.fit(np.zeros((15, 1)), np.zeros(4), 5, 3)


def test_cfg5_neighbourhood_is_regenerated_bit_for_bit(golden):
    """all_candidates(tuned_base, local_search=true) + expand/simplify + factor selection must
    reproduce the batch the UNMODIFIED reference produced (tests/golden/make_golden.py)."""
    z = golden("cfg5_neighbourhood")
    cands = M.debug_all_candidates(z["tuned_code"], z["tuned_consts"], 20, False, True)
    assert len(cands) == int(z["full_neighbourhood"])
    keep, skipped = [], 0
    for c in cands:
        if len(keep) == 4096:
            break
        tb = M.debug_term_batch([(c[0], c[1])])
        if tb["term_code_begin"][-1] > 50:
            skipped += 1
            continue
        keep.append((c[0], c[1]))
    assert skipped == int(z["skipped"])
    tb = M.debug_term_batch(keep)
    g = B.Batch.load_fields(z)
    assert np.array_equal(tb["cand_term_begin"], g.cand_term_begin)
    assert np.array_equal(tb["term_code_begin"], g.term_code_begin)
    assert np.array_equal(tb["code"], g.code)
    assert np.array_equal(tb["consts"].view(np.uint64), g.consts.view(np.uint64))
    assert M.debug_to_string(z["tuned_code"], z["tuned_consts"]) == str(z["tuned_base_str"])


def test_string_forms():
    v = B.Expr.var
    cases = {
        "(x0+3)": v(0) + 3.0, "(x0*0.500000)": v(0) * 0.5, "((x1)**2)": B.sqr(v(1)), "pow(x0,2)": B.pow_(v(0), 2.0),
        "(x0<x1)": v(0) > v(1), "MIN(x0, 3.141593)": B.min_(v(0), 3.14159265358979), "ln(sqrt(x2))": B.ln(B.sqrt(v(2))),
        "(-78.800000*ln(x1))": B.Expr.const(-78.8) * B.ln(v(1)), "1": B.Expr.const(1.0 + 1e-13),
    }
    for want, e in cases.items():
        code, consts = e.program()
        assert M.debug_to_string(code, consts) == want


def test_constant_formatting_is_printf_exact():
    """The dedupe key of every candidate prints constants like std::to_string(double) == printf("%f")
    (node.h:257-261); the driver's fast path must give the same characters, rounding boundaries included."""
    rng = np.random.default_rng(7)
    vals = list(rng.uniform(-10, 10, 4000)) + list(rng.uniform(-1e6, 1e6, 2000)) + list(rng.uniform(-1e-3, 1e-3, 2000))
    k = rng.integers(-10**9, 10**9, 3000)
    for base in ((k + 0.5) / 1e6, k / 1e6):  # ties of the 6th decimal and exact 6-decimal values, with neighbours
        vals += list(base) + list(np.nextafter(base, np.inf)) + list(np.nextafter(base, -np.inf))
    vals += [0.1, 0.2, 0.3, 78.8, 3.31, 1e-7, -1e-7, 5e-7, 999999.9999995, 1e6 + 0.5, -1e6 - 0.25, 1e15 + 0.5, 1e300, -1e300,
             4.9999995e-7, 2.5e-6, 123456.7890125]
    code = np.array([B.ins(B.OP_CONST, 0)], dtype=np.uint32)
    for v in vals:
        v = float(v)
        want = str(int(round(v))) if abs(round(v) - v) < 1e-12 and abs(v) < 2**31 else "%f" % v
        if abs(round(v) - v) < 1e-12 and abs(v) >= 2**31:
            continue  # (int) of an out-of-range double is undefined in the reference too
        assert M.debug_to_string(code, np.array([v])) == want, repr(v)


def test_rebuild_snaps_coefficients():
    """rils_rols_cpp.cpp:488-517"""
    v = B.Expr.var
    code, consts = (v(0) + B.sin(v(1)) + v(2)).program()
    s, size = M.debug_rebuild(code, consts, np.array([1.0 + 1e-13, 2.5, 1e-13, -3.0]))
    assert s == "((x0+(2.500000*sin(x1)))+-3)" and size == 8
    s, size = M.debug_rebuild(code, consts, np.array([0.0, 0.0, 0.0, 1e-14]))
    assert s == "0" and size == 1
    s, size = M.debug_rebuild(code, consts, np.array([0.0, 0.0, 0.0, 1.0]))
    assert s == "1" and size == 1


REF = O.load_ref()


@pytest.mark.skipif(REF is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("classification,d,seed", [(False, 2, 1), (False, 6, 2), (True, 4, 3), (True, 9, 4)])
@pytest.mark.parametrize("local_search", [False, True])
def test_candidate_generation_matches_reference_on_random_walks(classification, d, seed, local_search):
    """Random walks through neighbourhoods: at every step both generators must emit the same
    strings in the same order (this covers expand()'s preserved quirks, normalisation, dedupe)."""
    rng = np.random.default_rng(seed)
    h = REF.RefHarness(classification, 0.001, 50, 0)
    h.set_data(rng.uniform(0.5, 2, (30, d)), rng.normal(size=30))
    code, consts = B.Expr.const(0.0).program()
    steps = 0
    for step in range(7):
        ref = h.all_candidates(code, consts, local_search if step else False)
        mine = M.debug_all_candidates(code, consts, d, classification, local_search if step else False)
        assert [r[2] for r in ref] == [m[2] for m in mine], f"step {step}"
        for r, m in zip(ref[:: max(1, len(ref) // 50)], mine[:: max(1, len(ref) // 50)]):
            assert np.array_equal(r[0], m[0]) and np.array_equal(r[1].view(np.uint64), m[1].view(np.uint64))
        # factor selection + batch assembly vs the reference's own tune_constants
        pick = [ref[i] for i in rng.choice(len(ref), size=min(40, len(ref)), replace=False)]
        g = h.score_list([(c[0], c[1]) for c in pick], True)
        tb = M.debug_term_batch([(c[0], c[1]) for c in pick])
        assert np.array_equal(tb["cand_term_begin"], g["cand_term_begin"]) and np.array_equal(tb["code"], g["code"])
        assert np.array_equal(tb["consts"].view(np.uint64), g["consts"].view(np.uint64))
        # continue from a random, reasonably large candidate
        big = [c for c in ref if 3 <= len(c[0]) <= 25] or ref
        nxt = big[int(rng.integers(len(big)))]
        code, consts = nxt[0], nxt[1]
        steps += 1
    assert steps == 7


def test_reference_front_end_binds_this_module_unmodified():
    """the reference's Python front end stays unchanged and must keep working
    against the new module. Import it as it is (rils_rols/rils_rols.py:8 does `import rils_rols_cpp`) and
    check that it bound the module built here and that its estimators construct the drop-in class with the
    reference's positional signature (rils_rols_cpp.cpp:1000). fit() itself needs the GPU: tests/test_gpu_fit.py."""
    import hashlib
    import importlib

    ref_root = os.environ.get("RR_REFERENCE", "/root/reference")
    installed = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    base = next((b for b in (ref_root, installed) if os.path.isfile(os.path.join(b, "rils_rols", "rils_rols.py"))), None)
    if base is None:
        pytest.skip("reference Python package not available")
    if base not in sys.path:
        sys.path.append(base)
    fe = importlib.import_module("rils_rols.rils_rols")
    assert fe.rils_rols_cpp is M
    if base == installed and os.path.isdir(ref_root):
        for f in ("rils_rols.py", "utils.py", "__init__.py"):  # the installed copy is the reference's file, byte for byte
            a = hashlib.sha256(open(os.path.join(ref_root, "rils_rols", f), "rb").read()).hexdigest()
            b = hashlib.sha256(open(os.path.join(installed, "rils_rols", f), "rb").read()).hexdigest()
            assert a == b, f
    reg = fe.RILSROLSRegressor(max_fit_calls=10, max_time=1, random_state=3)
    clf = fe.RILSROLSBinaryClassifier(max_fit_calls=10, max_time=1, random_state=3)
    assert reg.classification is False and clf.classification is True
    # what fit_inner does at rils_rols.py:100 (positional pybind constructor)
    obj = fe.rils_rols_cpp.rils_rols(reg.classification, reg.max_fit_calls, reg.max_time, reg.complexity_penalty,
                                     reg.max_complexity, reg.sample_size, reg.verbose, reg.random_state)
    for name in ("fit", "predict", "get_model_string", "get_best_time", "get_fit_calls", "get_total_time"):
        assert callable(getattr(obj, name))
    with pytest.raises(Exception, match="not build yet"):
        reg.predict([[1.0, 2.0]])
