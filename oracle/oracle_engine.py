"""Python loader for the CPU oracle (TEST INFRASTRUCTURE ONLY — see mcs_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])
    return os.path.join(_HERE, "libmcs_oracle.so")


def load_oracle_library():
    global _LIB
    if _LIB is None:
        import mcs_b200
        p = os.path.join(_HERE, "libmcs_oracle.so")
        if not os.path.exists(p):
            build()
        _LIB = mcs_b200.abi.bind(C.CDLL(p))
        assert _LIB.mcs_backend().decode() == "cpu-oracle"
    return _LIB


def load_oracle_engine(cfg):
    import mcs_b200
    return mcs_b200.abi.Engine(load_oracle_library(), cfg)
