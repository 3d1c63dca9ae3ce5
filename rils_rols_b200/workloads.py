"""Synthetic / bundled inputs of the BASELINE.json configs.

configs 1-3 follow /root/reference/test_example.py:18-57; config 4/5 data follow the
generators SURVEY.md 8(d) fixes (numpy default_rng(12345), X ~ U(0.1, 3.0)).
"""
from __future__ import annotations

import math
import os
from typing import Tuple

import numpy as np

from . import batch as B

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cfg5_data(n: int = 1 << 24, d: int = 20) -> Tuple[np.ndarray, np.ndarray]:
    """(X row-major (n,d), y): y = sum_{i<5} sin(1/x_i) + 0.5 x5 x6 - 2 ln x7 + exp(-x8) + 0.1 N(0,1).

    X comes from default_rng(12345), the noise from default_rng(12346) (SURVEY.md 8(d) draws it
    from the same generator after X; a second generator makes every prefix of the data set
    independent of n, so tests at small n and the 2^24-row bench see the same rows)."""
    rng = np.random.default_rng(12345)
    X = rng.uniform(0.1, 3.0, size=(n, d))
    y = np.zeros(n)
    for i in range(5):
        y += np.sin(1.0 / X[:, i])
    y += 0.5 * X[:, 5] * X[:, 6] - 2.0 * np.log(X[:, 7]) + np.exp(-X[:, 8])
    y += 0.1 * np.random.default_rng(12346).standard_normal(n)
    return X, y


def cfg5_base_expr() -> B.Expr:
    """sin(1/x0)+sin(1/x1)+x5*x6+ln(x7)+exp(x8)*x9+sqrt(x10)/x11 (26 nodes)."""
    v = B.Expr.var
    return (((((B.sin(1.0 / v(0)) + B.sin(1.0 / v(1))) + v(5) * v(6)) + B.ln(v(7))) + B.exp(v(8)) * v(9))
            + B.sqrt(v(10)) / v(11))


def cfg4_data(n: int = 1_000_000, d: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    """test_large.py-style: y = sum_{i<5} sin(1/x_i)."""
    rng = np.random.default_rng(12345)
    X = rng.uniform(0.1, 3.0, size=(n, d))
    y = np.zeros(n)
    for i in range(5):
        y += np.sin(1.0 / X[:, i])
    return X, y


def cfg4_truth_expr() -> B.Expr:
    """sin(1/x0)+sin(1/x1)+sin(1/x2)+sin(1/x3)+sin(1/x4): the generating formula of cfg4_data."""
    v = B.Expr.var
    e = B.sin(1.0 / v(0))
    for i in range(1, 5):
        e = e + B.sin(1.0 / v(i))
    return e


def cfg5_neighbourhood() -> B.Batch:
    """The committed 4096-candidate OLS_FIT batch (tests/golden/cfg5_neighbourhood.npz)."""
    z = np.load(os.path.join(GOLDEN_DIR, "cfg5_neighbourhood.npz"))
    return B.Batch.load_fields(z)


def config_data(name: str, test: bool = False):
    """Training split of BASELINE configs 1-3 exactly as test_example.py builds it."""
    from sklearn.model_selection import train_test_split

    rs = 12345
    if name == "cfg1_toy":
        from random import randint, seed

        seed(rs)
        X = list(zip([randint(1, 100) for _ in range(200)], [randint(1, 100) for _ in range(200)]))
        y = [math.sin(x1) - 78.8 * math.log(x2) + 4 * x1 + 3.31 * x2 for x1, x2 in X]
    elif name == "cfg2_diabetes":
        from sklearn.datasets import load_diabetes

        X, y = load_diabetes(return_X_y=True)
    elif name == "cfg3_breast_cancer":
        from sklearn.datasets import load_breast_cancer

        X, y = load_breast_cancer(return_X_y=True)
    else:
        raise KeyError(name)
    Xtr, Xte, ytr, yte = train_test_split(X, y, train_size=0.75, test_size=0.25, random_state=rs)
    if test:
        return (np.asarray(Xtr, dtype=np.float64), np.asarray(ytr, dtype=np.float64),
                np.asarray(Xte, dtype=np.float64), np.asarray(yte, dtype=np.float64))
    return np.asarray(Xtr, dtype=np.float64), np.asarray(ytr, dtype=np.float64)
