"""ctypes front end of the C ABI (include/rr_b200.h) -> rils_rols_b200/librr_b200.so.

This is host plumbing only: every number comes from the CUDA library. There is no CPU
fallback — a missing library or a missing B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .batch import Batch, Result, rr_batch, rr_result

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RR_B200_LIB") or os.path.join(_HERE, "librr_b200.so")  # (RR_B200_LIB: A/B builds of the kernels)
_LIB: Optional[C.CDLL] = None

EXPORTS = [
    "rr_abi_version", "rr_last_error", "rr_engine_create", "rr_engine_create_rowmajor", "rr_engine_create_sharded",
    "rr_engine_destroy", "rr_comm_unique_id", "rr_engine_comm_init", "rr_engine_set_allreduce", "rr_engine_get_info",
    "rr_get_stats", "rr_score_batch", "rr_classifier_metrics", "rr_predict", "rr_predict_rowmajor",
    "rr_predict_proba_rowmajor", "rr_feature_r2", "rr_engine_read_rows", "rr_measure_fp64_peak",
    "rr_debug_plan_batch", "rr_debug_plan_free", "rr_debug_plan_concurrency_check", "rr_debug_const_terms",
]
ABI_VERSION = 3


class EngineError(RuntimeError):
    pass


class rr_engine_info(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_total", C.c_int64), ("d", C.c_int32), ("device", C.c_int32),
                ("y_mean", C.c_double), ("sst", C.c_double), ("sm_count", C.c_int32), ("exact_max_n", C.c_int32),
                ("n_gpus", C.c_int32), ("world", C.c_int32)]


class rr_stats(C.Structure):
    _fields_ = [("batches", C.c_uint64), ("candidates", C.c_uint64), ("sweep_launches", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("refined", C.c_uint64), ("exact", C.c_uint64), ("dd", C.c_uint64),
                ("nonfinite", C.c_uint64), ("distinct_terms", C.c_uint64), ("term_instances", C.c_uint64),
                ("distinct_dots", C.c_uint64), ("dot_instances", C.c_uint64), ("last_sweep_ms", C.c_double),
                ("last_batch_ms", C.c_double), ("w_contract", C.c_double), ("w_shared", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("last_host_ms", C.c_double),
                ("ingest_ms", C.c_double), ("collectives", C.c_uint64), ("row_groups", C.c_uint64),
                ("row_group_rows", C.c_uint64)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p)


def lib() -> C.CDLL:
    """Loads librr_b200.so (built by __graft_entry__.build() / make -C rils_rols_b200/csrc)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    vp = C.c_void_p
    L.rr_abi_version.restype = C.c_int
    L.rr_last_error.argtypes = [vp]
    L.rr_last_error.restype = C.c_char_p
    for name in ("rr_engine_create", "rr_engine_create_rowmajor"):
        f = getattr(L, name)
        f.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(vp)]
        f.restype = C.c_int
    L.rr_engine_create_sharded.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, C.c_int64, C.c_int32, C.c_uint32, C.POINTER(vp)]
    L.rr_engine_create_sharded.restype = C.c_int
    L.rr_engine_destroy.argtypes = [vp]
    L.rr_engine_destroy.restype = None
    L.rr_comm_unique_id.argtypes = [vp]
    L.rr_comm_unique_id.restype = C.c_int
    L.rr_engine_comm_init.argtypes = [vp, vp, C.c_int32, C.c_int32]
    L.rr_engine_comm_init.restype = C.c_int
    L.rr_engine_set_allreduce.argtypes = [vp, ALLREDUCE_FN, vp, C.c_int32, C.c_int32]
    L.rr_engine_set_allreduce.restype = C.c_int
    L.rr_engine_get_info.argtypes = [vp, C.POINTER(rr_engine_info)]
    L.rr_engine_get_info.restype = C.c_int
    L.rr_get_stats.argtypes = [vp, C.POINTER(rr_stats)]
    L.rr_get_stats.restype = C.c_int
    L.rr_score_batch.argtypes = [vp, C.POINTER(rr_batch), C.POINTER(rr_result)]
    L.rr_score_batch.restype = C.c_int
    L.rr_classifier_metrics.argtypes = [vp, C.POINTER(rr_batch), dp, dp, dp]
    L.rr_classifier_metrics.restype = C.c_int
    for name in ("rr_predict", "rr_predict_rowmajor"):
        f = getattr(L, name)
        f.argtypes = [vp, up, C.c_int32, dp, C.c_int32, dp, C.c_int64, C.c_int32, dp]
        f.restype = C.c_int
    L.rr_predict_proba_rowmajor.argtypes = [vp, up, C.c_int32, dp, C.c_int32, dp, C.c_int64, C.c_int32, dp]
    L.rr_predict_proba_rowmajor.restype = C.c_int
    L.rr_feature_r2.argtypes = [vp, dp]
    L.rr_feature_r2.restype = C.c_int
    L.rr_engine_read_rows.argtypes = [vp, C.c_int64, C.c_int64, vp, vp]
    L.rr_engine_read_rows.restype = C.c_int
    L.rr_measure_fp64_peak.argtypes = [vp, dp]
    L.rr_measure_fp64_peak.restype = C.c_int
    _LIB = L
    return L


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Engine:
    """One scoring engine = one (shard of a) data set resident on one B200.

    Mirrors what rils_rols::fit hands to its hot path (X, y of
    /root/reference/rils_rols_cpp/rils_rols_cpp.cpp:717-728): X row-major (n, d) like the
    numpy array the pybind boundary receives, or feature-major (d, n) with rowmajor=False.
    """

    def __init__(self, X: np.ndarray, y: np.ndarray, rowmajor: bool = True, device: int = -1, flags: int = 0):
        L = lib()
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if X.ndim != 2 or y.ndim != 1:
            raise ValueError("X must be 2-D and y 1-D")
        n, d = (X.shape if rowmajor else X.shape[::-1])
        if y.shape[0] != n:
            raise ValueError(f"Size of y {y.shape[0]} is not the same as the data count {n}")
        self._h = C.c_void_p()
        fn = L.rr_engine_create_rowmajor if rowmajor else L.rr_engine_create
        rc = fn(X.ctypes.data, y.ctypes.data, n, d, device, flags, C.byref(self._h))
        if rc != 0:
            raise EngineError(f"rr_engine_create failed ({rc}): {L.rr_last_error(None).decode()}")
        self.n, self.d = int(n), int(d)
        self._cb = None

    @classmethod
    def sharded(cls, X: np.ndarray, y: np.ndarray, n_gpus: int = 0, row_index: Optional[np.ndarray] = None,
                rowmajor: bool = True, flags: int = 0) -> "Engine":
        """ONE engine object over n_gpus devices of this process (0 = all visible): rows in contiguous blocks,
        NCCL inside the engine (rr_engine_create_sharded). row_index selects and orders the rows on the device
        (the shuffle / sub-sample of rils_rols_cpp.cpp:774-795)."""
        from .batch import FLAG_X_ROWMAJOR

        L = lib()
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        n, d = (X.shape if rowmajor else X.shape[::-1])
        if y.shape[0] != n:
            raise ValueError(f"Size of y {y.shape[0]} is not the same as the data count {n}")
        idx, n_rows = None, n
        if row_index is not None:
            idx = np.ascontiguousarray(row_index, dtype=np.int32)
            n_rows = int(idx.size)
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        rc = L.rr_engine_create_sharded(X.ctypes.data, y.ctypes.data, n, d, idx.ctypes.data if idx is not None else None,
                                        n_rows, n_gpus, flags | (FLAG_X_ROWMAJOR if rowmajor else 0), C.byref(self._h))
        if rc != 0:
            raise EngineError(f"rr_engine_create_sharded failed ({rc}): {L.rr_last_error(None).decode()}")
        self.n, self.d = int(n_rows), int(d)
        self._cb = None
        return self

    @classmethod
    def from_device(cls, x_ptr: int, y_ptr: int, n: int, d: int, device: int = -1, flags: int = 0) -> "Engine":
        """X (feature-major, d x n) and y already in device memory (raw pointers)."""
        from .batch import FLAG_X_DEVICE

        L = lib()
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        rc = L.rr_engine_create(x_ptr, y_ptr, n, d, device, flags | FLAG_X_DEVICE, C.byref(self._h))
        if rc != 0:
            raise EngineError(f"rr_engine_create failed ({rc}): {L.rr_last_error(None).decode()}")
        self.n, self.d = int(n), int(d)
        self._cb = None
        return self

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {lib().rr_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            lib().rr_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def info(self) -> rr_engine_info:
        i = rr_engine_info()
        self._check(lib().rr_engine_get_info(self._h, C.byref(i)), "rr_engine_get_info")
        return i

    def stats(self) -> dict:
        s = rr_stats()
        self._check(lib().rr_get_stats(self._h, C.byref(s)), "rr_get_stats")
        return s.as_dict()

    def score(self, batch: Batch, out: Optional[Result] = None) -> Result:
        res = out if out is not None else Result.alloc(batch)
        bs, rs = batch.as_struct(), res.as_struct()
        self._check(lib().rr_score_batch(self._h, C.byref(bs), C.byref(rs)), "rr_score_batch")
        return res

    def classifier_metrics(self, batch: Batch):
        nc = max(batch.n_cand, 1)
        acc, ll, al = np.zeros(nc), np.zeros(nc), np.zeros(nc)
        bs = batch.as_struct()
        self._check(lib().rr_classifier_metrics(self._h, C.byref(bs), _dp(acc), _dp(ll), _dp(al)),
                    "rr_classifier_metrics")
        return acc[: batch.n_cand], ll[: batch.n_cand], al[: batch.n_cand]

    def predict(self, code: np.ndarray, consts: np.ndarray, X: np.ndarray, rowmajor: bool = True) -> np.ndarray:
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, d = (X.shape if rowmajor else X.shape[::-1])
        code = np.ascontiguousarray(code, dtype=np.uint32)
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        k = consts if consts.size else np.zeros(1)
        out = np.empty(n)
        fn = lib().rr_predict_rowmajor if rowmajor else lib().rr_predict
        self._check(fn(self._h, code.ctypes.data_as(C.POINTER(C.c_uint32)), code.size, _dp(k), consts.size, _dp(X),
                       n, d, _dp(out)), "rr_predict")
        return out

    def predict_proba(self, code: np.ndarray, consts: np.ndarray, X: np.ndarray) -> np.ndarray:
        """(n, 2) array [1 - p, p], p = logistic(2 (yhat - 0.5)) (rr_predict_proba_rowmajor)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, d = X.shape
        code = np.ascontiguousarray(code, dtype=np.uint32)
        consts = np.ascontiguousarray(consts, dtype=np.float64)
        k = consts if consts.size else np.zeros(1)
        out = np.empty((n, 2))
        self._check(lib().rr_predict_proba_rowmajor(self._h, code.ctypes.data_as(C.POINTER(C.c_uint32)), code.size, _dp(k),
                                                    consts.size, _dp(X), n, d, _dp(out)), "rr_predict_proba")
        return out

    def feature_r2(self) -> np.ndarray:
        """R2(X[j], y) of relevant_features(), rils_rols_cpp.cpp:753-770, for every feature (one device reduction)."""
        out = np.empty(self.d)
        self._check(lib().rr_feature_r2(self._h, _dp(out)), "rr_feature_r2")
        return out

    def read_rows(self, row0: int, rows: int):
        """(X feature-major (d, rows), y) of the engine's resident rows [row0, row0 + rows)."""
        X = np.empty((self.d, rows))
        y = np.empty(rows)
        self._check(lib().rr_engine_read_rows(self._h, row0, rows, X.ctypes.data, y.ctypes.data), "rr_engine_read_rows")
        return X, y

    def comm_init_torch(self, group=None):
        """One process per GPU: NCCL INSIDE the engine. Rank 0 draws the unique id, torch.distributed only carries
        its 128 bytes to the other ranks; every later all-reduce is issued by the engine itself on its own stream
        (no Python, no host synchronisation in the step)."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            rc = lib().rr_comm_unique_id(buf)
            if rc != 0:
                raise EngineError(f"rr_comm_unique_id failed ({rc}): {lib().rr_last_error(None).decode()}")
        obj = [bytes(buf)]
        dist.broadcast_object_list(obj, src=0, group=group)
        raw = (C.c_ubyte * 128).from_buffer_copy(obj[0])
        self._check(lib().rr_engine_comm_init(self._h, raw, rank, world), "rr_engine_comm_init")

    def fp64_peak(self) -> float:
        v = C.c_double()
        self._check(lib().rr_measure_fp64_peak(self._h, C.byref(v)), "rr_measure_fp64_peak")
        return v.value

    def set_allreduce_torch(self, group=None):
        """Sample-sharded multi-GPU: sum the per-candidate partial reductions over all ranks with
        torch.distributed (NCCL over NVLink), one all-reduce per sweep."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", self.info().device)

        def _cb(ptr, count, stream, _user):
            try:
                from .torch_interop import tensor_from_ptr

                t = tensor_from_ptr(ptr, count, dev)
                ext = torch.cuda.ExternalStream(stream, device=dev)
                with torch.cuda.stream(ext):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                # The collective is stream-ordered already; waiting here makes the host wait for the sweep in
                # front of it, which serialises run_gram's two-half planning pipeline when rows are sharded.
                # RR_B200_HOOK_ASYNC=1 leaves it asynchronous (to be validated on >= 2 GPUs before it
                # becomes the default).
                if os.environ.get("RR_B200_HOOK_ASYNC", "0") in ("", "0"):
                    ext.synchronize()
                return 0
            except Exception as ex:  # pragma: no cover - surfaced as RR_ERR_COLLECTIVE
                print(f"[rils_rols_b200] all-reduce hook failed: {ex!r}", flush=True)
                return 1

        self._cb = ALLREDUCE_FN(_cb)
        self._check(lib().rr_engine_set_allreduce(self._h, self._cb, None, rank, world), "rr_engine_set_allreduce")
