"""CPU oracle (test infrastructure only): NumPy/SymPy restatement of the reference's IEKS hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs import this.
"""
from . import ivps, pof_oracle  # noqa: F401
