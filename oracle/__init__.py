"""CPU oracle for the render hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  Nothing under raingun_b200/ does.
"""
from .oracle import Oracle, build_oracle, oracle  # noqa: F401
