"""sklearn-style front end: RILSROLSRegressor / RILSROLSBinaryClassifier.

Same constructor parameters, methods and behaviour as the reference's front end
(/root/reference/rils_rols/rils_rols.py:16-188, utils.py:6-35) so that code written against
`rils_rols.rils_rols` runs unchanged; underneath, `rils_rols_cpp` is this repo's pybind11 module
(host ILS driver + B200 scoring engine). The reference front end itself also works unmodified
against that module: put rils_rols_b200/ on PYTHONPATH.
"""
from __future__ import annotations

import math
import multiprocessing.pool
import time
import warnings
from math import inf

import numpy as np
from sklearn.base import BaseEstimator
from sklearn.metrics import accuracy_score, r2_score
from sklearn.model_selection import train_test_split

from . import rils_rols_cpp

warnings.filterwarnings("ignore")


def complexity_sympy(model) -> int:  # utils.py:6-10
    from sympy import preorder_traversal

    return sum(1 for _ in preorder_traversal(model))


def noisefy(y, noise_level, random_state):  # utils.py:12-19
    yn = np.array(y, dtype=float)
    rng = np.random.RandomState(random_state)
    return yn + rng.normal(0, np.sqrt(np.mean(np.square(yn))) * noise_level, size=len(yn))


def logistic(x):  # utils.py:21-24
    return 1.0 / (1.0 + math.exp(-x))


def binarize(yp):  # utils.py:29-31
    return np.array([1 if v >= 0.5 else 0 for v in yp])


def proba(yp):  # utils.py:33-35
    return np.array([[1 - logistic(v), logistic(v)] for v in yp])


def _with_timeout(seconds, fn, *args):
    pool = multiprocessing.pool.ThreadPool(processes=1)
    try:
        return pool.apply_async(fn, args).get(seconds)
    finally:
        pool.terminate()


class RILSROLSBase(BaseEstimator):
    def __init__(self, classification=None, max_fit_calls=100000, max_time=100, complexity_penalty=0.001,
                 max_complexity=50, sample_size=1, verbose=False, random_state=0):
        print(f'Calling with max_fit_calls={max_fit_calls} max_time={max_time} complexity_penalty={complexity_penalty} '
              f'max_complexity={max_complexity} sample_size={sample_size} verbose={verbose} random_state={random_state}')
        self.classification = classification
        self.max_time = max_time
        self.max_fit_calls = max_fit_calls
        self.max_complexity = max_complexity
        self.complexity_penalty = complexity_penalty
        self.sample_size = sample_size
        self.verbose = verbose
        self.random_state = random_state
        self.rr_cpp = None
        self.model = None
        self.model_simp = None

    def fit(self, X, y):
        if self.sample_size == 0:  # automatic tuning of the sample size, rils_rols.py:51-85
            print('Automatically tuning sample size:')
            if len(X) <= 10000:
                print('Training size is smaller than 10000 so setting it to full sample 1 (100%).')
                self.sample_size = 1
            else:
                start = time.time()
                total_max_fit_calls = self.max_fit_calls
                self.max_fit_calls = self.max_fit_calls / 100
                tuning_fit_calls = 0
                best_ss = 1
                for ss in [0.1]:
                    X_sample, _, y_sample, _ = train_test_split(X, y, train_size=ss, random_state=self.random_state)
                    self.sample_size = 1
                    self.fit_inner(X_sample, y_sample)
                    tuning_fit_calls += self.max_fit_calls
                    yp_sample = self.predict(X_sample)
                    yp = self.predict(X)
                    try:
                        r2_sample = r2_score(y_sample, yp_sample)
                        r2 = r2_score(y, yp)
                        print(f'Sample size {ss} --> R2={r2} R2_sample={r2_sample}')
                        if abs(r2 - r2_sample) < 0.01:
                            best_ss = ss
                            break
                    except Exception:
                        print('Error while calculating R2.')
                        r2 = -inf
                print(f'Setting sample_size={best_ss}')
                self.sample_size = best_ss
                self.max_time -= time.time() - start
                self.max_fit_calls = total_max_fit_calls - tuning_fit_calls
        elif self.sample_size < 0 or self.sample_size > 1:
            raise Exception('Sample size parameter must belong to interval (0, 1], while value 0 means it is automatically tuned.')
        self.fit_inner(X, y)

    def fit_inner(self, X, y):
        import pandas as pd
        from sympy import simplify, sympify

        if isinstance(X, pd.DataFrame):
            X = X.values.tolist()
        if isinstance(y, pd.DataFrame):
            y = y.values.tolist()
        self.rr_cpp = rils_rols_cpp.rils_rols(self.classification, int(self.max_fit_calls), int(self.max_time),
                                              self.complexity_penalty, self.max_complexity, self.sample_size,
                                              self.verbose, int(self.random_state))
        X = np.array(X)
        data_cnt, feat_cnt = X.shape[0], X.shape[1]
        X = X.reshape(data_cnt * feat_cnt, 1)
        y = np.array(y)
        self.rr_cpp.fit(X, y, data_cnt, feat_cnt)
        self.model = self.rr_cpp.get_model_string()
        self.best_time = self.rr_cpp.get_best_time()
        self.total_time = self.rr_cpp.get_total_time()
        self.fit_calls = self.rr_cpp.get_fit_calls()
        if self.classification is True:
            self.model_simp = self.model
        else:
            try:  # sympy simplify under a 2 s budget, rils_rols.py:46-48,119-125
                self.model_simp = _with_timeout(2.0, lambda: simplify(self.model, ratio=1))
            except Exception:
                self.model_simp = sympify(self.model)
        return (self.model, self.model_simp)

    def check_model(self):
        if self.model is None or self.rr_cpp is None:
            raise Exception("Cannot predict because model is not build yet. First call fit().")

    def predict(self, X):
        import pandas as pd

        if isinstance(X, pd.DataFrame):
            X = X.values.tolist()
        self.check_model()
        X = np.array(X)
        data_cnt, feat_cnt = X.shape[0], X.shape[1]
        return self.rr_cpp.predict(X.reshape(data_cnt * feat_cnt, 1), data_cnt, feat_cnt)

    def model_string(self):
        self.check_model()
        return self.model_simp

    def fit_report_string(self):
        self.check_model()
        return ("maxTime={0}\tmaxFitCalls={1}\tseed={2}\tsizePenalty={3}\tmaxComplexity={4}\tsampleShare={5}\t"
                "totalTime={6:.1f}\tbestTime={7}\tfitCalls={8}\tsimpSize={9}\texpr={10}\texprSimp={11}").format(
            self.max_time, self.max_fit_calls, self.random_state, self.complexity_penalty, self.max_complexity,
            self.sample_size, self.total_time, self.best_time, self.fit_calls, complexity_sympy(self.model_simp),
            self.model, self.model_simp)


class RILSROLSRegressor(RILSROLSBase):
    def __init__(self, max_fit_calls=100000, max_time=100, complexity_penalty=0.001, max_complexity=50, sample_size=1,
                 verbose=False, random_state=0):
        super().__init__(False, max_fit_calls, max_time, complexity_penalty, max_complexity, sample_size, verbose,
                         random_state)

    def score(self, X, y):
        return r2_score(y, self.predict(X))


class RILSROLSBinaryClassifier(RILSROLSBase):
    def __init__(self, max_fit_calls=100000, max_time=100, complexity_penalty=0.001, max_complexity=50, sample_size=1,
                 verbose=False, random_state=0):
        super().__init__(True, max_fit_calls, max_time, complexity_penalty, max_complexity, sample_size, verbose,
                         random_state)

    def check_binary_targets(self, y):
        for yi in y:
            if yi != 0 and yi != 1:
                raise Exception('The classifier works only for binary targets, so allowed target values are 0 or 1.')

    def predict(self, X):
        return binarize(super().predict(X))

    def predict_proba(self, X):
        return proba(self.predict(X))

    def score(self, X, y):
        self.check_binary_targets(y)
        return accuracy_score(y, self.predict(X))

    def fit(self, X, y):
        self.check_binary_targets(y)
        return super().fit(X, y)
