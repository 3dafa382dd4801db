"""TEST INFRASTRUCTURE ONLY (see gymnasium/__init__.py)."""
from .registration import register, registry, make   # noqa: F401
