"""TEST INFRASTRUCTURE ONLY -- stand-in for `pybullet_data` (see oracle/shim/pybullet.py)."""
import os


def getDataPath():
    return os.path.dirname(__file__)
