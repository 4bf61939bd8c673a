"""B200-native per-particle transport loop for MonteCarloScattering.jl (hot path only).

Layout:
  csrc/        CUDA kernels (sm_100a) + the C-ABI of include/mcs.h  -> libmcs_b200.so
  abi.py       ctypes mirror of include/mcs.h
  engine.py    loads libmcs_b200.so (fails loudly when it is missing: there is NO CPU fallback)
  problem.py   host-side producer of the kernel's inputs (grid, profile, PSD scalars, init_pop)
  driver.py    mirror of main_loops.jl's iteration / ion / pcut nest around the C-ABI calls

Import as `import mcs_b200` (the repo-root shim) because the directory name is not an identifier.
"""
from . import abi, problem  # noqa: F401
from .engine import load_cuda_library, load_cuda_engine, lib_path, build  # noqa: F401
from .driver import make_config, species_struct, main_loops, run_ion_host_comm  # noqa: F401
