"""TEST INFRASTRUCTURE ONLY -- stand-in for `pybullet_utils` (see oracle/shim/pybullet.py)."""
