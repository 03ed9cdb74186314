"""sp1_b200 — B200-native (sm_100a) core-shard prover hot path for SP1 v6 "Hypercube".

The product is the C-ABI shared library `libsp1b200.so` (include/sp1b200.h); this package is the thin
host-side mirror used by tests and bench.py.  There is no CPU fallback: importing `sp1_b200.lib` without the
built library raises.
"""
from .lib import Lib, load, SO_PATH, DEFAULT_CORE_PARAMS  # noqa: F401
