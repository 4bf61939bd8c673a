"""Loader for the CUDA library. There is no CPU fallback: a missing or non-CUDA library is an error."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path() -> str:
    """In-tree CUDA library. MCS_LIB may point at another build of the SAME sources (tuning variants from
    `make -C csrc variants`); it is never a CPU library: load_cuda_library() checks the backend string."""
    return os.environ.get("MCS_LIB") or os.path.join(_HERE, "libmcs_b200.so")


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libmcs_b200.so (in-tree, so it travels to the GPU box)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("nvcc build of libmcs_b200.so failed")
    return lib_path()


def load_cuda_library() -> C.CDLL:
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise RuntimeError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                               "There is no CPU fallback for the transport loop.")
        lib = abi.bind(C.CDLL(p))
        if not lib.mcs_backend().decode().startswith("cuda"):
            raise RuntimeError("libmcs_b200.so does not report a CUDA backend")
        _LIB = lib
    return _LIB


def load_cuda_engine(cfg: abi.McsConfig) -> abi.Engine:
    return abi.Engine(load_cuda_library(), cfg)
