"""rils_rols_b200 — B200-native scoring engine for the RILS-ROLS hot path.

Layout: csrc/ (CUDA kernels, planner, C ABI -> librr_b200.so; host ILS driver ->
rils_rols_cpp pybind11 module), engine.py (ctypes front end of the C ABI), batch.py
(host containers), rils_rols.py (sklearn-style front end mirroring the reference).
There is no CPU fallback: every scoring entry point fails loudly without the CUDA
library and a B200.
"""
__version__ = "0.1.0"
