"""Host-side containers for the C ABI of include/rr_b200.h (ctypes mirrors).

A *program* is the postfix bytecode of one expression tree: one uint32 per node,
low 8 bits = opcode (== enum node_type, /root/reference/rils_rols_cpp/node.h:16-38),
high 24 bits = feature index (VAR) or constant-pool index (CONST).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Iterable, List, Sequence, Tuple

import numpy as np

# opcodes (enum rr_opcode)
OP_NONE, OP_CONST, OP_VAR, OP_PLUS, OP_MINUS, OP_MULTIPLY, OP_DIVIDE = range(7)
OP_SIN, OP_COS, OP_LN, OP_EXP, OP_SQRT, OP_SQR, OP_POW = range(7, 14)
OP_LESS_THAN, OP_GREATER_THAN, OP_EQUAL, OP_NOT_EQUAL, OP_MIN, OP_MAX = range(14, 20)
OP_NAMES = ["none", "const", "var", "+", "-", "*", "/", "sin", "cos", "ln", "exp", "sqrt", "sqr", "pow",
            "<", ">", "==", "!=", "min", "max"]
ARITY = [0, 0, 0, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2]

MODE_EVAL_ONLY = 0
MODE_OLS_FIT = 1

FLAG_FORCE_GRAM = 1 << 0
FLAG_FORCE_EXACT = 1 << 1
FLAG_NO_CSE = 1 << 2
FLAG_X_DEVICE = 1 << 3
FLAG_X_ROWMAJOR = 1 << 4

RES_NONFINITE = 1 << 0
RES_RANKDEF = 1 << 1
RES_REFINED = 1 << 2
RES_EXACT = 1 << 3
RES_DD = 1 << 4
RES_SLOWPATH = 1 << 5

# SURVEY.md 8(d) contract weights: FP64-pipe thread-instructions per node
W_OP = [0, 0, 0, 1, 1, 1, 10, 16, 16, 28, 18, 10, 1, 90, 1, 1, 1, 1, 1, 1]


def ins(op: int, arg: int = 0) -> int:
    return (op & 0xFF) | (arg << 8)


class Expr:
    """Tiny expression builder producing postfix programs (tests, bench, examples)."""

    __slots__ = ("op", "a", "b", "val")

    def __init__(self, op, a=None, b=None, val=None):
        self.op, self.a, self.b, self.val = op, a, b, val

    @staticmethod
    def var(j: int) -> "Expr":
        return Expr(OP_VAR, val=int(j))

    @staticmethod
    def const(c: float) -> "Expr":
        return Expr(OP_CONST, val=float(c))

    @staticmethod
    def _wrap(x) -> "Expr":
        return x if isinstance(x, Expr) else Expr.const(x)

    def __add__(self, o): return Expr(OP_PLUS, self, Expr._wrap(o))
    def __radd__(self, o): return Expr(OP_PLUS, Expr._wrap(o), self)
    def __sub__(self, o): return Expr(OP_MINUS, self, Expr._wrap(o))
    def __rsub__(self, o): return Expr(OP_MINUS, Expr._wrap(o), self)
    def __mul__(self, o): return Expr(OP_MULTIPLY, self, Expr._wrap(o))
    def __rmul__(self, o): return Expr(OP_MULTIPLY, Expr._wrap(o), self)
    def __truediv__(self, o): return Expr(OP_DIVIDE, self, Expr._wrap(o))
    def __rtruediv__(self, o): return Expr(OP_DIVIDE, Expr._wrap(o), self)
    def __lt__(self, o): return Expr(OP_LESS_THAN, self, Expr._wrap(o))
    def __gt__(self, o): return Expr(OP_GREATER_THAN, self, Expr._wrap(o))

    def emit(self, code: List[int], consts: List[float]) -> None:
        if self.op == OP_CONST:
            code.append(ins(OP_CONST, len(consts)))
            consts.append(self.val)
        elif self.op == OP_VAR:
            code.append(ins(OP_VAR, self.val))
        else:
            self.a.emit(code, consts)
            if ARITY[self.op] == 2:
                self.b.emit(code, consts)
            code.append(ins(self.op))

    def program(self) -> Tuple[np.ndarray, np.ndarray]:
        code: List[int] = []
        consts: List[float] = []
        self.emit(code, consts)
        return np.asarray(code, dtype=np.uint32), np.asarray(consts, dtype=np.float64)


def unary(op: int, a) -> Expr:
    return Expr(op, Expr._wrap(a))


def binary(op: int, a, b) -> Expr:
    return Expr(op, Expr._wrap(a), Expr._wrap(b))


def sin(a): return unary(OP_SIN, a)
def cos(a): return unary(OP_COS, a)
def ln(a): return unary(OP_LN, a)
def exp(a): return unary(OP_EXP, a)
def sqrt(a): return unary(OP_SQRT, a)
def sqr(a): return unary(OP_SQR, a)
def pow_(a, b): return binary(OP_POW, a, b)
def eq(a, b): return binary(OP_EQUAL, a, b)
def ne(a, b): return binary(OP_NOT_EQUAL, a, b)
def min_(a, b): return binary(OP_MIN, a, b)
def max_(a, b): return binary(OP_MAX, a, b)


class rr_batch(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("n_cand", C.c_int32),
        ("cand_term_begin", C.POINTER(C.c_int32)),
        ("term_code_begin", C.POINTER(C.c_int32)),
        ("code", C.POINTER(C.c_uint32)),
        ("consts", C.POINTER(C.c_double)),
        ("n_consts", C.c_int32),
        ("n_code", C.c_int32),
    ]


class rr_result(C.Structure):
    _fields_ = [
        ("coef", C.POINTER(C.c_double)),
        ("nonzero_pivots", C.POINTER(C.c_int32)),
        ("ssr", C.POINTER(C.c_double)),
        ("flags", C.POINTER(C.c_uint32)),
    ]


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


@dataclass
class Batch:
    """One neighbourhood: the arrays of struct rr_batch."""

    mode: int
    cand_term_begin: np.ndarray  # int32 [n_cand+1]
    term_code_begin: np.ndarray  # int32 [n_terms+1]
    code: np.ndarray             # uint32
    consts: np.ndarray           # float64

    def __post_init__(self):
        self.cand_term_begin = np.ascontiguousarray(self.cand_term_begin, dtype=np.int32)
        self.term_code_begin = np.ascontiguousarray(self.term_code_begin, dtype=np.int32)
        self.code = np.ascontiguousarray(self.code, dtype=np.uint32)
        self.consts = np.ascontiguousarray(self.consts, dtype=np.float64)
        if self.consts.size == 0:  # keep a valid pointer
            self.consts = np.zeros(1, dtype=np.float64)
            self._n_consts = 0
        else:
            self._n_consts = int(self.consts.size)

    @property
    def n_cand(self) -> int:
        return int(self.cand_term_begin.size - 1)

    @property
    def n_terms(self) -> int:
        return int(self.term_code_begin.size - 1)

    @property
    def n_coef(self) -> int:
        return self.n_terms + self.n_cand

    def coef_slice(self, c: int) -> slice:
        b = int(self.cand_term_begin[c]) + c
        return slice(b, b + int(self.cand_term_begin[c + 1] - self.cand_term_begin[c]) + 1)

    def as_struct(self) -> rr_batch:
        s = rr_batch()
        s.mode = self.mode
        s.n_cand = self.n_cand
        s.cand_term_begin = _ptr(self.cand_term_begin, C.c_int32)
        s.term_code_begin = _ptr(self.term_code_begin, C.c_int32)
        s.code = _ptr(self.code, C.c_uint32)
        s.consts = _ptr(self.consts, C.c_double)
        s.n_consts = self._n_consts
        s.n_code = int(self.code.size)
        return s

    def subset(self, idx: Sequence[int]) -> "Batch":
        """Batch holding only the candidates `idx` (constants are kept whole)."""
        ctb, tcb, code = [0], [0], []
        for c in idx:
            for t in range(int(self.cand_term_begin[c]), int(self.cand_term_begin[c + 1])):
                code.extend(self.code[self.term_code_begin[t]:self.term_code_begin[t + 1]].tolist())
                tcb.append(len(code))
            ctb.append(len(tcb) - 1)
        return Batch(self.mode, np.asarray(ctb), np.asarray(tcb), np.asarray(code, dtype=np.uint32),
                     self.consts[: self._n_consts].copy())

    def contract_work(self) -> np.ndarray:
        """no-sharing FP64 thread-instructions per sample, per candidate."""
        w_node = np.asarray(W_OP, dtype=np.float64)[self.code & 0xFF]
        cs = np.concatenate([[0.0], np.cumsum(w_node)])
        term_w = cs[self.term_code_begin[1:]] - cs[self.term_code_begin[:-1]]
        ts = np.concatenate([[0.0], np.cumsum(term_w)])
        w = ts[self.cand_term_begin[1:]] - ts[self.cand_term_begin[:-1]]
        if self.mode == MODE_OLS_FIT:
            k = (self.cand_term_begin[1:] - self.cand_term_begin[:-1]).astype(np.float64) + 1.0
            w = w + k * (k + 1) / 2 + k + k + 2
        else:
            w = w + 2
        return w

    @staticmethod
    def from_programs(mode: int, cands: Iterable[Sequence[Tuple[np.ndarray, np.ndarray]]]) -> "Batch":
        """cands: for each candidate a list of (code, consts) term programs."""
        ctb, tcb, code, consts = [0], [0], [], []
        for terms in cands:
            for tcode, tconsts in terms:
                base = len(consts)
                for w in np.asarray(tcode, dtype=np.uint32).tolist():
                    if (w & 0xFF) == OP_CONST:
                        w = ins(OP_CONST, (w >> 8) + base)
                    code.append(w)
                consts.extend(np.asarray(tconsts, dtype=np.float64).tolist())
                tcb.append(len(code))
            ctb.append(len(tcb) - 1)
        return Batch(mode, np.asarray(ctb), np.asarray(tcb), np.asarray(code, dtype=np.uint32),
                     np.asarray(consts, dtype=np.float64))

    @staticmethod
    def from_exprs(mode: int, cands: Iterable[Sequence[Expr]]) -> "Batch":
        return Batch.from_programs(mode, [[e.program() for e in terms] for terms in cands])

    def save_fields(self, prefix: str = "") -> dict:
        return {
            prefix + "mode": np.int32(self.mode),
            prefix + "cand_term_begin": self.cand_term_begin,
            prefix + "term_code_begin": self.term_code_begin,
            prefix + "code": self.code,
            prefix + "consts": self.consts[: self._n_consts],
        }

    @staticmethod
    def load_fields(z, prefix: str = "") -> "Batch":
        return Batch(int(z[prefix + "mode"]), z[prefix + "cand_term_begin"], z[prefix + "term_code_begin"],
                     z[prefix + "code"], z[prefix + "consts"])


@dataclass
class Result:
    coef: np.ndarray
    nonzero_pivots: np.ndarray
    ssr: np.ndarray
    flags: np.ndarray
    _keep: list = field(default_factory=list, repr=False)

    @staticmethod
    def alloc(batch: Batch) -> "Result":
        return Result(np.zeros(max(batch.n_coef, 1)), np.zeros(max(batch.n_cand, 1), dtype=np.int32),
                      np.zeros(max(batch.n_cand, 1)), np.zeros(max(batch.n_cand, 1), dtype=np.uint32))

    def as_struct(self) -> rr_result:
        s = rr_result()
        s.coef = _ptr(self.coef, C.c_double)
        s.nonzero_pivots = _ptr(self.nonzero_pivots, C.c_int32)
        s.ssr = _ptr(self.ssr, C.c_double)
        s.flags = _ptr(self.flags, C.c_uint32)
        return s


def fitness_tuple(ssr: float, sst: float, n: int, size: int) -> Tuple[float, float, int]:
    """(1-R2, RMSE, size) exactly as fitness() forms it, rils_rols_cpp.cpp:520-541, :40-49."""
    with np.errstate(all="ignore"):
        r2 = 1.0 - np.float64(ssr) / np.float64(sst)
        rmse = np.sqrt(np.float64(ssr) / np.float64(n))
    if r2 != r2 or rmse != rmse:
        return (1000.0, 1000.0, 1000)
    return (float(1.0 - r2), float(rmse), int(size))
