"""TEST INFRASTRUCTURE — ctypes front end of oracle/librr_oracle.so (the plain-C
restatement, rr_oracle.c) and loader of oracle/_ref (the unmodified reference built
against the Eigen stand-in). See oracle/Makefile."""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import subprocess
import sys
from typing import Optional, Tuple

import numpy as np

from rils_rols_b200.batch import Batch, Result, rr_batch, rr_result

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None


def build(ref: bool = True) -> None:
    """Compile the C restatement and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    if ref and os.path.isdir(os.environ.get("RR_REFERENCE", "/root/reference")):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "librr_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _LIB = C.CDLL(path)
        dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
        _LIB.rr_oracle_score_batch.argtypes = [dp, dp, C.c_int64, C.c_int32, C.POINTER(rr_batch),
                                               C.POINTER(rr_result), dp, dp, ip]
        _LIB.rr_oracle_score_batch.restype = C.c_int
        _LIB.rr_oracle_eval.argtypes = [up, C.c_int32, dp, C.c_int32, dp, C.c_int64, C.c_int32, dp]
        _LIB.rr_oracle_eval.restype = C.c_int
        _LIB.rr_oracle_classifier_metrics.argtypes = [dp, dp, C.c_int64, C.c_int32, C.POINTER(rr_batch), dp, dp, dp]
        _LIB.rr_oracle_classifier_metrics.restype = C.c_int
    return _LIB


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def feature_major(X_rowmajor: np.ndarray) -> np.ndarray:
    """(n,d) row-major -> (d,n) contiguous, the layout of rils_rols_cpp.cpp:675-698."""
    return np.ascontiguousarray(np.asarray(X_rowmajor, dtype=np.float64).T)


def score_batch(X_fm: np.ndarray, y: np.ndarray, batch: Batch):
    """-> (Result, f0, f1, size): the oracle's answer for one rr_batch."""
    X_fm = np.ascontiguousarray(X_fm, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    d, n = X_fm.shape
    res = Result.alloc(batch)
    f0 = np.zeros(max(batch.n_cand, 1))
    f1 = np.zeros(max(batch.n_cand, 1))
    fs = np.zeros(max(batch.n_cand, 1), dtype=np.int32)
    bs, rs = batch.as_struct(), res.as_struct()
    rc = lib().rr_oracle_score_batch(_dp(X_fm), _dp(y), n, d, C.byref(bs), C.byref(rs), _dp(f0), _dp(f1),
                                     fs.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise ValueError(f"rr_oracle_score_batch failed rc={rc}")
    nc = batch.n_cand
    return res, f0[:nc], f1[:nc], fs[:nc]


def evaluate(X_fm: np.ndarray, code: np.ndarray, consts: np.ndarray) -> np.ndarray:
    X_fm = np.ascontiguousarray(X_fm, dtype=np.float64)
    d, n = X_fm.shape
    code = np.ascontiguousarray(code, dtype=np.uint32)
    consts = np.ascontiguousarray(consts, dtype=np.float64)
    k = consts if consts.size else np.zeros(1)
    out = np.empty(n)
    rc = lib().rr_oracle_eval(code.ctypes.data_as(C.POINTER(C.c_uint32)), code.size, _dp(k), consts.size,
                              _dp(X_fm), n, d, _dp(out))
    if rc != 0:
        raise ValueError(f"rr_oracle_eval failed rc={rc}")
    return out


def classifier_metrics(X_fm: np.ndarray, y: np.ndarray, batch: Batch):
    X_fm = np.ascontiguousarray(X_fm, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    d, n = X_fm.shape
    acc, ll, al = (np.zeros(max(batch.n_cand, 1)) for _ in range(3))
    bs = batch.as_struct()
    rc = lib().rr_oracle_classifier_metrics(_dp(X_fm), _dp(y), n, d, C.byref(bs), _dp(acc), _dp(ll), _dp(al))
    if rc != 0:
        raise ValueError(f"rr_oracle_classifier_metrics failed rc={rc}")
    nc = batch.n_cand
    return acc[:nc], ll[:nc], al[:nc]


def load_ref():
    """Import oracle/_ref/rils_rols_cpp_ref (None if it was never built)."""
    d = os.path.join(_HERE, "_ref")
    if not os.path.isdir(d):
        return None
    for f in os.listdir(d):
        if f.startswith("rils_rols_cpp_ref") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location("rils_rols_cpp_ref", os.path.join(d, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            sys.modules["rils_rols_cpp_ref"] = mod
            return mod
    return None
