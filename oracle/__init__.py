"""CPU oracle for the MAPF step/observation path (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product package (mapf_marl_b200) never does.
See oracle/mapf_oracle.c for the restated algorithms and their reference citations.
"""
from .oracle import Oracle, build_oracle, oracle_max_threads  # noqa: F401
