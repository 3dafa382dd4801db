"""TEST INFRASTRUCTURE ONLY -- single-process stand-in for `mpi4py` so the reference's
`algs/` and `utils/` modules import (rollout-collector goldens).  world size is 1."""
from . import MPI   # noqa: F401
