#!/usr/bin/env python
"""Which candidates of the headline neighbourhood leave the plain Gram solve (refinement / double-double / rank-deficient)?
Prints their term lists. Usage: python tools/list_escalated.py [n_rows]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rils_rols_b200 import batch as B  # noqa: E402
from rils_rols_b200 import workloads as W  # noqa: E402
from rils_rols_b200.engine import Engine  # noqa: E402

NAMES = {1: 'c', 2: 'x', 3: '+', 4: '-', 5: '*', 6: '/', 7: 'sin', 8: 'cos', 9: 'ln', 10: 'exp', 11: 'sqrt', 12: 'sqr'}


def dec(code, consts):
    st = []
    for w in code:
        op, arg = int(w) & 0xff, int(w) >> 8
        if op == 1: st.append('%.4g' % consts[arg])
        elif op == 2: st.append('x%d' % arg)
        elif op in (3, 4, 5, 6):
            b_ = st.pop(); a = st.pop(); st.append('(%s%s%s)' % (a, NAMES[op], b_))
        else:
            a = st.pop(); st.append('%s(%s)' % (NAMES.get(op, str(op)), a))
    return st[-1]


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
X, y = W.cfg5_data(n)
b = W.cfg5_neighbourhood()
with Engine(X, y) as eng:
    r = eng.score(b)
fl = np.asarray(r.flags)
for name, bit in (("dd", B.RES_DD), ("refined", B.RES_REFINED), ("rankdef", B.RES_RANKDEF)):
    idx = np.nonzero(fl & bit)[0]
    print(f"== {name}: {len(idx)} candidates")
    for c in idx[:60]:
        ts = [dec(b.code[b.term_code_begin[t]:b.term_code_begin[t + 1]], b.consts) for t in range(b.cand_term_begin[c], b.cand_term_begin[c + 1])]
        print(int(c), hex(int(fl[c])), ' | '.join(ts))
