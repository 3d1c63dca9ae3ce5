"""TEST INFRASTRUCTURE — CPU oracle of the RILS-ROLS scoring hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package. The product (rils_rols_b200/) never does.
"""
