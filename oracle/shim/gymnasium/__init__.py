"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for `gymnasium` so that the reference
package imports unmodified (oracle/gen_golden.py).  Not used by the product package."""
from . import spaces, core, envs          # noqa: F401
from .core import Env, Wrapper            # noqa: F401
from .envs.registration import make, register, registry   # noqa: F401
