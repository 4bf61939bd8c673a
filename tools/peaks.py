"""Print the roofline denominators measured by the library on the current GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcs_b200
from mcs_b200 import problem, driver, abi
run = problem.setup_run(problem.planar_test_particle_input(1000))
lib = mcs_b200.load_cuda_library()
e = abi.Engine(lib, driver.make_config(lib, run))
print("fp64 DFMA peak  [TFLOP/s]:", e.measure_fp64_peak())
print("scatter-only    [steps/s]:", e.measure_scatter_peak())
for n in (1 << 10, 1 << 16, 2_757_447, 1 << 26):
    print(f"fp64 red.global scattered over {n:>9d} cells [Gop/s]:", e.measure_atomic_peak(n))
