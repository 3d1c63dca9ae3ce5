"""CPU emulator of the engine's internal instruction set (rils_rols_b200/csrc/rr_isa.h).

Test infrastructure: executes the instruction stream the planner produced (obtained through the
host-only rr_debug_plan_batch entry point) with numpy, one n-vector per tile column, so the
compiler half of the engine — term hashing, slot allocation, spills, fused reductions, chunking —
can be checked against the oracle without a GPU. It is NOT a fallback of the product: nothing in
rils_rols_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from rils_rols_b200 import engine as E
from rils_rols_b200.batch import Batch, rr_batch

RR_NPIN = 8
RR_NREG = 10
(RI_END, RI_WINEND, RI_LOAD_C, RI_ST, RI_STG, RI_LDG, RI_NOP, RI_COMBINE, RI_ADD_C, RI_SUB_C, RI_RSUB_C, RI_MUL_C, RI_DIV_C,
 RI_RDIV_C, RI_SIN, RI_COS, RI_LN, RI_EXP, RI_SQRT, RI_SQR, RI_RARE, RI_MDOT, RI_GRAM8, RI_PINBG, RI_MDOTDD, RI_CLSMET,
 RI_PIN0) = range(27)
RR_INS_WINDOW = 64
RI_LDP0 = RI_PIN0 + RR_NREG
RI_USEP0 = RI_LDP0 + RR_NREG
RI_MULP0 = RI_USEP0 + RR_NREG
RI_DIVP0 = RI_MULP0 + RR_NREG
RI_RDIVP0 = RI_DIVP0 + RR_NREG
RI_CMULP0 = RI_RDIVP0 + RR_NREG
RI_CDIVP0 = RI_CMULP0 + RR_NREG
RI_FIRST_M = RI_CDIVP0 + RR_NREG
(RI_LOAD_M, RI_ADD_M, RI_SUB_M, RI_RSUB_M, RI_MUL_M, RI_DIV_M, RI_RDIV_M, RI_AXPY, RI_DOTM,
 RI_DOTMDD, RI_CMUL_M, RI_CDIV_M, RI_MUL_MM, RI_MUL_M_ST, RI_MUL_MMM, RI_LDPMUL_M0) = range(RI_FIRST_M, RI_FIRST_M + 16)
RI_LDPDIV_M0 = RI_LDPMUL_M0 + RR_NREG
RI_LDMDIVP0 = RI_LDPDIV_M0 + RR_NREG
RI_PINB0 = RI_LDMDIVP0 + RR_NREG
RI_OPCOUNT = RI_PINB0 + RR_NPIN
RR_MDOT_MAX_OUT = 8
RR_POW, RR_LT, RR_GT, RR_EQ, RR_NE, RR_MIN, RR_MAX = range(7)
RB_CONST, RB_SWAP = 1 << 4, 1 << 5
RR_THEN_MDOT = 0x8000
RR_THEN_ST = 0x8000  # G8 plans: the same bit means "X; ST c", c in bits 16-23 of w0
RR_MDOT_ROWS = 1
RR_GRAM_COLS = 2


def ring_rows(aux: int, cnt: int):
    """Rows (count & 15) of the 10 potential outputs self, one, pins 0..7 of an RI_MDOT with this aux when
    `cnt` reductions went through the ring before it; an unwanted output takes the row of the next wanted
    one or the first free row (rr_plan.cpp close())."""
    want = (aux & 3) | (((aux >> 8) & 0xFF) << 2)
    rows, r = [0] * 10, cnt
    for o in range(10):
        if want >> o & 1:
            rows[o] = r & 15
            r += 1
    nxt = r & 15
    for o in range(9, -1, -1):
        if want >> o & 1: nxt = rows[o]
        else: rows[o] = nxt
    return rows, r


def md_fusable(op: int) -> bool:
    return (op in (RI_MUL_M, RI_DIV_M, RI_RDIV_M, RI_DIV_C, RI_RDIV_C, RI_CMUL_M, RI_CDIV_M, RI_MUL_MM, RI_MUL_MMM)
            or RI_MULP0 <= op < RI_FIRST_M or RI_LDPMUL_M0 <= op < RI_PINB0)


INS_DT = np.dtype([("w0", "<u4"), ("w1", "<u4"), ("imm", "<f8")])
CHUNK_DT = np.dtype([("pc_begin", "<i4"), ("n_ins", "<i4"), ("dot_base", "<i4"), ("n_dots", "<i4"),
                     ("col_begin", "<i4"), ("n_cols", "<i4"), ("r0", "<i4"), ("r1", "<i4")])

KIND_GRAM, KIND_GRAM_DD, KIND_EVAL, KIND_EVAL_METRICS, KIND_MATERIALISE, KIND_RESIDUAL, KIND_GRAM_G8, KIND_GRAM_R8 = range(8)


class rr_debug_plan(C.Structure):
    _fields_ = [("n_ins", C.c_int64), ("n_chunks", C.c_int64), ("n_cols", C.c_int64), ("n_tab", C.c_int64),
                ("n_tab_begin", C.c_int64), ("n_term_ids", C.c_int64), ("ins", C.c_void_p), ("chunks", C.c_void_p),
                ("cols", C.POINTER(C.c_int32)), ("tab", C.POINTER(C.c_int32)), ("tab_begin", C.POINTER(C.c_int32)),
                ("term_ids", C.POINTER(C.c_int32)), ("n_dots", C.c_int32), ("max_tile_cols", C.c_int32),
                ("n_terms_distinct", C.c_int32), ("reserved", C.c_int32), ("w_issued", C.c_double),
                ("w_contract", C.c_double), ("error", C.c_char * 256)]


class Plan:
    def __init__(self, batch: Batch, d: int, kind: int, tile_cols: int = 56, max_slots: int = 0,
                 target_chunks: int = 1, no_cse: bool = False, coef=None, n_pins: int = RR_NPIN):
        L = E.lib()
        L.rr_debug_plan_batch.argtypes = [C.POINTER(rr_batch), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(rr_debug_plan)]
        L.rr_debug_plan_batch.restype = C.c_int
        L.rr_debug_plan_free.argtypes = [C.POINTER(rr_debug_plan)]
        L.rr_debug_plan_free.restype = None
        out = rr_debug_plan()
        bs = batch.as_struct()
        cp = None
        if coef is not None:
            coef = np.ascontiguousarray(coef, dtype=np.float64)
            cp = coef.ctypes.data_as(C.POINTER(C.c_double))
        rc = L.rr_debug_plan_batch(C.byref(bs), d, kind, tile_cols, max_slots, target_chunks, int(no_cse),
                                   int(n_pins), cp, C.byref(out))
        if rc != 0:
            raise ValueError(out.error.decode())

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dtype=dt)
            buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr if isinstance(ptr, int) else C.addressof(ptr.contents))
            return np.frombuffer(buf, dtype=dt).copy()

        self.ins = arr(out.ins, out.n_ins, INS_DT)
        self.chunks = arr(out.chunks, out.n_chunks, CHUNK_DT)
        self.cols = arr(out.cols, out.n_cols, np.int32)
        self.tab = arr(out.tab, out.n_tab, np.int32)
        self.tab_begin = arr(out.tab_begin, out.n_tab_begin, np.int32)
        self.term_ids = arr(out.term_ids, out.n_term_ids, np.int32)
        self.n_dots, self.max_tile_cols = out.n_dots, out.max_tile_cols
        self.n_terms_distinct = out.n_terms_distinct
        self.w_issued, self.w_contract = out.w_issued, out.w_contract
        self.kind = kind
        self.has_rows = bool(((self.ins["w0"] & 0xFFFF) == (RI_NOP | RR_MDOT_ROWS << 8)).any())
        L.rr_debug_plan_free(C.byref(out))


def run(plan: Plan, cols_global: np.ndarray, n_stg: int = 0):
    """cols_global: (d+2, n) = features, y, y - mean(y). Returns (dots[n_dots], stg[n_stg, n])."""
    n = cols_global.shape[1]
    dots = np.zeros(max(plan.n_dots, 1))
    stg = np.zeros((n_stg, n))
    n_gram = [0, 0]  # GRAM8 instructions, rows
    plan.gram_stats = n_gram

    def do_mdot(op, aux, t, pins, out):
        step = 2 if op == RI_MDOTDD else 1
        vals = []
        if aux & 1: vals.append(float(np.dot(t, t)))
        if aux & 2: vals.append(float(np.sum(t)))
        mask = (aux >> 8) & 0xFF
        for j in range(RR_NPIN):
            if mask >> j & 1:
                assert pins[j] is not None, "MDOT against an empty pin"
                vals.append(float(np.dot(t, pins[j])))
        assert 0 < len(vals) <= RR_MDOT_MAX_OUT
        for v in vals:
            dots[out] += v
            out += step
        if aux >> 16:  # fused "then pin t"
            j = (aux >> 16) - 1
            assert j < RR_NREG and not (mask >> j & 1)
            pins[j] = t.copy()
        return out

    with np.errstate(all="ignore"):
        for ch in plan.chunks:
            ncols = int(ch["n_cols"])
            tile = {i: cols_global[plan.cols[ch["col_begin"] + i]] for i in range(ncols)}
            t = np.zeros(n)
            pins = [None] * RR_NREG
            use_pin = -1
            ring_cnt = 0          # reductions pushed through the warp ring so far (kernel: cnt)
            ring_fl = 0           # ... and flushed (kernel: fl); the core flushes 8 rows at a time
            want_combine = False  # the last flush filled the fourth staging slot: RI_COMBINE must follow
            rows_expected = None  # set by an RI_MDOT carrier, checked against the data slot behind it
            out = int(ch["dot_base"])
            pc = int(ch["pc_begin"])
            end = pc + int(ch["n_ins"])
            while pc < end:
                assert rows_expected is None or (int(plan.ins["w0"][pc]) & 0xFFFF) == (RI_NOP | RR_MDOT_ROWS << 8) \
                    or not plan.has_rows, "RI_MDOT carrier without its data slot"
                w0, w1, imm = int(plan.ins["w0"][pc]), int(plan.ins["w1"][pc]), float(plan.ins["imm"][pc])
                if want_combine and plan.has_rows and (w0 & 0xFFFF) != (RI_NOP | RR_MDOT_ROWS << 8):
                    assert (w0 & 0xFF) == RI_COMBINE, "missing RI_COMBINE behind the flush that filled the staging rows"
                col2 = int(plan.ins["imm"][pc:pc + 1].view(np.uint64)[0] & 0xFFFFFFFF)  # second column of the fused forms
                pc += 1
                op, aux = w0 & 0xFF, w0 >> 8
                src = None
                if op >= RI_FIRST_M:
                    if use_pin >= 0:
                        src = pins[use_pin]
                        assert src is not None, "USEP of an empty pin"
                    else:
                        assert w1 < plan.max_tile_cols, f"tile column {w1} out of range"
                        src = tile[w1]
                else:
                    assert use_pin < 0, "USEP must be followed by a tile-column operand form"
                use_pin = -1
                if op == RI_END:
                    break
                elif op == RI_COMBINE:
                    assert want_combine, "RI_COMBINE without a full group of staged reductions"
                    want_combine = False
                elif op == RI_NOP:
                    if aux == RR_MDOT_ROWS:  # data slot: the ring rows of the preceding instruction's reductions
                        assert rows_expected is not None, "MDOT data slot without an MDOT in front"
                        got = list(plan.ins[pc - 1:pc].view(np.uint8)[4:14])
                        assert got == rows_expected, f"ring rows {got} != {rows_expected}"
                        assert (pc - 1 - int(ch["pc_begin"])) % RR_INS_WINDOW != 0, "data slot at a window start"
                        rows_expected = None
                        continue
                elif op == RI_LOAD_C: t = np.full(n, imm)
                elif op == RI_LOAD_M: t = src.copy()
                elif op == RI_ST:
                    assert w1 < plan.max_tile_cols, f"tile column {w1} out of range"
                    tile[w1] = t.copy()
                elif op == RI_STG: stg[w1] = t
                elif op == RI_LDG: t = cols_global[w1].copy()
                elif op == RI_ADD_C: t = t + imm
                elif op == RI_ADD_M: t = t + src
                elif op == RI_SUB_C: t = t - imm
                elif op == RI_SUB_M: t = t - src
                elif op == RI_RSUB_C: t = imm - t
                elif op == RI_RSUB_M: t = src - t
                elif op == RI_MUL_C: t = t * imm
                elif op == RI_MUL_M: t = t * src
                elif op == RI_DIV_C: t = t / imm
                elif op == RI_DIV_M: t = t / src
                elif op == RI_RDIV_C: t = imm / t
                elif op == RI_RDIV_M: t = src / t
                elif op == RI_AXPY: t = t + imm * src
                elif op in (RI_DOTM, RI_DOTMDD):
                    dots[out] += float(np.dot(t, src))
                    out += 2 if op == RI_DOTMDD else 1
                    if op == RI_DOTM:  # pushes, then flushes behind itself
                        ring_cnt += 1
                        if ring_cnt - ring_fl >= 8:
                            ring_fl += 8
                            want_combine = ring_fl % 32 == 0
                elif RI_PIN0 <= op < RI_PIN0 + RR_NREG: pins[op - RI_PIN0] = t.copy()
                elif RI_LDP0 <= op < RI_LDP0 + RR_NREG:
                    assert pins[op - RI_LDP0] is not None, "LDP of an empty pin"
                    t = pins[op - RI_LDP0].copy()
                elif RI_USEP0 <= op < RI_USEP0 + RR_NREG:
                    assert (pc - 1 - int(ch["pc_begin"])) % RR_INS_WINDOW != RR_INS_WINDOW - 1, "USEP at a window end"
                    use_pin = op - RI_USEP0
                elif RI_MULP0 <= op < RI_FIRST_M:  # fused value-register forms
                    j = (op - RI_MULP0) % RR_NREG
                    assert pins[j] is not None, "fused form reads an empty value register"
                    kind = (op - RI_MULP0) // RR_NREG
                    if kind == 0: t = t * pins[j]
                    elif kind == 1: t = t / pins[j]
                    elif kind == 2: t = pins[j] / t
                    elif kind == 3: t = imm * pins[j]
                    else: t = imm / pins[j]
                elif op == RI_CMUL_M: t = imm * src
                elif op == RI_CDIV_M: t = imm / src
                elif op == RI_MUL_MM:
                    assert col2 < plan.max_tile_cols, f"tile column {col2} out of range"
                    t = src * tile[col2]
                elif op == RI_MUL_MMM:
                    col3 = int(plan.ins["imm"][pc - 1:pc].view(np.uint64)[0] >> 32)
                    assert col2 < plan.max_tile_cols and col3 < plan.max_tile_cols and plan.kind == KIND_GRAM_G8
                    t = (src * tile[col2]) * tile[col3]
                elif op == RI_MUL_M_ST:
                    assert col2 < plan.max_tile_cols, f"tile column {col2} out of range"
                    t = t * src
                    tile[col2] = t.copy()
                elif op == RI_GRAM8:
                    # up to 8 rows (tile columns named by the column slot behind the instruction) against the 8 pins,
                    # themselves and ones; wanted outputs in bit order take consecutive ids
                    assert plan.kind == KIND_GRAM_G8
                    n_rows = aux & 0xFF
                    assert 1 <= n_rows <= 8
                    assert (pc - 1 - int(ch["pc_begin"])) % RR_INS_WINDOW != RR_INS_WINDOW - 1, "GRAM8 at a window end"
                    bits = w1 | (int(plan.ins["imm"][pc - 1:pc].view(np.uint64)[0]) << 32)
                    d0 = int(plan.ins["w0"][pc])
                    assert (d0 & 0xFFFF) == (RI_NOP | RR_GRAM_COLS << 8), "GRAM8 without its column slot"
                    cols8 = list(plan.ins[pc:pc + 1].view(np.uint8)[4:12])
                    pc += 1
                    assert bits >> (10 * n_rows) == 0, "wanted bits beyond the last row"
                    for g in range(n_rows):
                        assert cols8[g] < plan.max_tile_cols, f"row column {cols8[g]} out of range"
                        a = tile[cols8[g]]
                        for o in range(10):
                            if bits >> (10 * g + o) & 1:
                                if o < 8:
                                    assert pins[o] is not None, "GRAM8 against an empty pin"
                                    v = float(np.dot(a, pins[o]))
                                else:
                                    v = float(np.dot(a, a)) if o == 8 else float(np.sum(a))
                                dots[out] += v
                                out += 1
                    n_gram[0] += 1
                    n_gram[1] += n_rows
                elif op == RI_PINBG:
                    pins[aux & 0xFF] = cols_global[w1].copy()
                elif RI_PINB0 <= op < RI_PINB0 + RR_NPIN:
                    pins[op - RI_PINB0] = src.copy()
                elif RI_LDPMUL_M0 <= op < RI_PINB0:
                    j = (op - RI_LDPMUL_M0) % RR_NREG
                    assert pins[j] is not None, "fused form reads an empty value register"
                    kind = (op - RI_LDPMUL_M0) // RR_NREG
                    t = pins[j] * src if kind == 0 else (pins[j] / src if kind == 1 else src / pins[j])
                elif op == RI_SIN: t = np.sin(t)
                elif op == RI_COS: t = np.cos(t)
                elif op == RI_LN: t = np.log(t)
                elif op == RI_EXP: t = np.exp(t)
                elif op == RI_SQRT: t = np.sqrt(t)
                elif op == RI_SQR: t = t * t
                elif op == RI_RARE:
                    u = np.full(n, imm) if aux & RB_CONST else tile[w1]
                    x, v = (u, t) if aux & RB_SWAP else (t, u)
                    r = aux & 0xF
                    if r == RR_POW: t = np.power(x, v)
                    elif r == RR_LT: t = (x < v).astype(float)
                    elif r == RR_GT: t = (x > v).astype(float)
                    elif r == RR_EQ: t = (x == v).astype(float)
                    elif r == RR_NE: t = (x != v).astype(float)
                    elif r == RR_MIN: t = np.where(x < v, x, v)
                    else: t = np.where(x > v, x, v)
                elif op in (RI_MDOT, RI_MDOTDD):
                    out = do_mdot(op, aux, t, pins, out)
                    if op == RI_MDOT:
                        if ring_cnt - ring_fl >= 8:  # flushes on entry, then pushes
                            ring_fl += 8
                            want_combine = ring_fl % 32 == 0
                        rows_expected, ring_cnt = ring_rows(aux, ring_cnt)
                elif op == RI_CLSMET:
                    y = tile[w1]
                    ypb, yb = (t >= 0.5).astype(float), (y >= 0.5).astype(float)
                    prob = 1.0 / (1.0 + np.exp(-2.0 * (t - 0.5)))
                    dots[out] += float(np.sum(ypb == yb))
                    dots[out + 1] += float(-np.sum((1.0 - yb) * np.log(1.0 - prob) + yb * np.log(prob)))
                    dots[out + 2] += float(np.sum(np.abs(yb - t)))
                    out += 3
                    ring_cnt += 3
                else:
                    raise AssertionError(f"bad opcode {op}")
                if (w0 & RR_THEN_ST) and md_fusable(op) and plan.kind == KIND_GRAM_G8:  # "X; ST c" in one instruction
                    c = (w0 >> 16) & 0xFF
                    assert c < plan.max_tile_cols, f"tile column {c} out of range"
                    tile[c] = t.copy()
                elif (w0 & RR_THEN_MDOT) and md_fusable(op):  # "X; MDOT" in one instruction
                    out = do_mdot(RI_MDOT, aux & ~(RR_THEN_MDOT >> 8), t, pins, out)
                    if ring_cnt - ring_fl >= 8:
                        ring_fl += 8
                        want_combine = ring_fl % 32 == 0
                    rows_expected, ring_cnt = ring_rows(aux, ring_cnt)
            assert out == int(ch["dot_base"]) + int(ch["n_dots"]), "chunk dot count mismatch"
    return dots[: plan.n_dots], stg


# ---- R8 plans: the row machine (rr_isa.h RQ_*) ----
(RQ_END, RQ_WINEND, RQ_NOP, RQ_LD, RQ_ADD, RQ_SUB, RQ_RSUB, RQ_MUL, RQ_DIV, RQ_RDIV, RQ_RARE, RQ_SIN, RQ_COS, RQ_LN,
 RQ_EXP, RQ_SQRT, RQ_SQR, RQ_TU, RQ_ST, RQ_GRAM, RQ_PINB, RQ_OPCOUNT) = range(22)
RQ_M, RQ_K, RQ_U, RQ_C = 0, 1, 2, 3
RQ_SWAP, RQ_PIN_GLOBAL = 1 << 11, 1 << 24


def run_r8(plan: Plan, cols_global: np.ndarray):
    """Emulates an R8 plan: eight rows of one shape per group. Returns (dots[n_dots], stats dict)."""
    assert plan.kind == KIND_GRAM_R8
    n = cols_global.shape[1]
    dots = np.zeros(max(plan.n_dots, 1))
    m = np.ones(n, dtype=bool)
    stats = dict(groups=0, rows=0, stores=0, pinb=0, ops=0)
    with np.errstate(all="ignore"):
        for ch in plan.chunks:
            ncols = int(ch["n_cols"])
            tile = {i: cols_global[plan.cols[ch["col_begin"] + i]].copy() for i in range(ncols)}
            t = np.zeros((8, n))
            u = np.zeros((8, n))
            pins = [None] * RR_NPIN
            D = np.zeros((8, 10))
            pc = int(ch["pc_begin"])
            end = pc + int(ch["n_ins"])
            base = int(ch["dot_base"])
            n_out = 0
            ended = False
            while pc < end:
                w0, w1 = int(plan.ins["w0"][pc]), int(plan.ins["w1"][pc])
                imm = float(plan.ins["imm"][pc])
                immbits = int(plan.ins["imm"][pc:pc + 1].view(np.uint64)[0])
                c8 = list(plan.ins[pc:pc + 1].view(np.uint8)[8:16])
                pc += 1
                op, mode = w0 & 0xFF, (w0 >> 8) & 3
                if op == RQ_END:
                    ended = True
                    break
                if op == RQ_NOP:
                    continue
                stats["ops"] += 1
                b = None
                if RQ_LD <= op <= RQ_RARE:
                    if mode == RQ_M:
                        for c in c8:
                            assert c < plan.max_tile_cols and c in tile, f"tile column {c} not available"
                        b = np.stack([tile[c] for c in c8])
                    elif mode == RQ_K:
                        b = np.full((8, n), imm)
                    elif mode == RQ_C:
                        assert (pc - 1 - int(ch["pc_begin"])) % RR_INS_WINDOW <= RR_INS_WINDOW - 5, "constants straddle a window"
                        k8 = plan.ins[pc:pc + 4].view(np.float64)
                        assert k8.shape == (8,)
                        pc += 4
                        b = np.repeat(k8[:, None], n, axis=1)
                    else:
                        assert mode == RQ_U
                        b = u
                if op == RQ_LD: t[:, m] = b[:, m]
                elif op == RQ_ADD: t[:, m] = (t + b)[:, m]
                elif op == RQ_SUB: t[:, m] = (t - b)[:, m]
                elif op == RQ_RSUB: t[:, m] = (b - t)[:, m]
                elif op == RQ_MUL: t[:, m] = (t * b)[:, m]
                elif op == RQ_DIV: t[:, m] = (t / b)[:, m]
                elif op == RQ_RDIV: t[:, m] = (b / t)[:, m]
                elif op == RQ_RARE:
                    x, v = (b, t) if w0 & RQ_SWAP else (t, b)
                    r = (w0 >> 12) & 0xF
                    if r == RR_POW: res = np.power(x, v)
                    elif r == RR_LT: res = (x < v).astype(float)
                    elif r == RR_GT: res = (x > v).astype(float)
                    elif r == RR_EQ: res = (x == v).astype(float)
                    elif r == RR_NE: res = (x != v).astype(float)
                    elif r == RR_MIN: res = np.where(x < v, x, v)
                    else: res = np.where(x > v, x, v)
                    t[:, m] = res[:, m]
                elif op == RQ_SIN: t[:, m] = np.sin(t[:, m])
                elif op == RQ_COS: t[:, m] = np.cos(t[:, m])
                elif op == RQ_LN: t[:, m] = np.log(t[:, m])
                elif op == RQ_EXP: t[:, m] = np.exp(t[:, m])
                elif op == RQ_SQRT: t[:, m] = np.sqrt(t[:, m])
                elif op == RQ_SQR: t[:, m] = (t * t)[:, m]
                elif op == RQ_TU: u[:, m] = t[:, m]
                elif op == RQ_ST:
                    assert mode == RQ_M
                    for g in range(8):
                        c = c8[g]
                        assert ncols <= c < plan.max_tile_cols, f"store to tile column {c}"
                        if c not in tile:
                            tile[c] = np.zeros(n)
                        tile[c][m] = t[g, m]
                    stats["stores"] += 1
                elif op == RQ_PINB:
                    j = (w0 >> 16) & 0xFF
                    assert j < RR_NPIN
                    if w0 & RQ_PIN_GLOBAL:
                        pins[j] = cols_global[w1].copy()
                    else:
                        assert w1 in tile, f"PINB from tile column {w1}"
                        pins[j] = tile[w1].copy()
                    stats["pinb"] += 1
                elif op == RQ_GRAM:
                    n_rows = (w0 >> 16) & 0xFF
                    assert 1 <= n_rows <= 8
                    for g in range(n_rows):
                        a = t[g, m]
                        for o in range(8):
                            if pins[o] is not None:
                                D[g, o] += float(np.dot(a, pins[o][m]))
                        D[g, 8] += float(np.dot(a, a))
                        D[g, 9] += float(np.sum(a))
                    if True:
                        assert (pc - 1 - int(ch["pc_begin"])) % RR_INS_WINDOW != RR_INS_WINDOW - 1, "GRAM at a window end"
                        d0 = int(plan.ins["w0"][pc])
                        assert (d0 & 0xFF) == RQ_NOP, "GRAM without its data slot"
                        bits = immbits | (int(plan.ins["w1"][pc]) << 64)
                        pc += 1
                        assert bits >> (10 * n_rows) == 0, "wanted bits beyond the last row"
                        out = base + w1
                        assert w1 == n_out, "group outputs are not consecutive"
                        for g in range(n_rows):
                            for o in range(10):
                                if bits >> (10 * g + o) & 1:
                                    if o < 8:
                                        assert pins[o] is not None, "GRAM against an empty pin"
                                    dots[out] += D[g, o]
                                    out += 1
                        n_out = out - base
                        D[:] = 0.0
                        stats["groups"] += 1
                        stats["rows"] += n_rows
                else:
                    raise AssertionError(f"bad R8 opcode {op}")
            assert ended, "chunk without END"
            assert n_out == int(ch["n_dots"]), "chunk dot count mismatch"
    return dots[: plan.n_dots], stats


def engine_columns(X_rowmajor: np.ndarray, y: np.ndarray) -> np.ndarray:
    Xfm = np.ascontiguousarray(np.asarray(X_rowmajor, dtype=np.float64).T)
    return np.vstack([Xfm, y[None, :], (y - y.mean())[None, :]])
