"""CPU oracle of the CL-SLAM adaptation hot path.  TEST INFRASTRUCTURE ONLY: importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never
from clslam_b200/ (the product path has no CPU fallback)."""
