"""rils_rols_b200 — B200-native scoring engine for the RILS-ROLS hot path.

Layout: csrc/ (CUDA kernels, planner, C ABI -> librr_b200.so; host ILS driver ->
rils_rols_cpp pybind11 module), engine.py (ctypes front end of the C ABI), batch.py
(host containers). The sklearn-style front end is the reference's own, unmodified
(/root/reference/rils_rols/rils_rols.py): it does `import rils_rols_cpp`, and with this
directory on sys.path that name resolves to the module built here.
There is no CPU fallback: every scoring entry point fails loudly without the CUDA
library and a B200.
"""
import os
import sys

__version__ = "0.2.0"

_HERE = os.path.dirname(os.path.abspath(__file__))


def driver_module():
    """The pybind11 module `rils_rols_cpp` (csrc/driver/rr_pymodule.cpp) under the top-level name the
    reference's front end imports (rils_rols/rils_rols.py:8). One name only: a pybind11 extension
    cannot be imported twice under two names in one process."""
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    import rils_rols_cpp

    if os.path.dirname(os.path.abspath(rils_rols_cpp.__file__)) != _HERE:
        raise ImportError(f"rils_rols_cpp resolves to {rils_rols_cpp.__file__}, not to the module built in {_HERE}")
    return rils_rols_cpp
