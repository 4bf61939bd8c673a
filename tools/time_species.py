"""Per-species device time of one iteration of a multi-species configuration (development aid).
usage: python tools/time_species.py [n_per_pcut] [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mcs_b200
from mcs_b200 import abi, driver, problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
workload = sys.argv[2] if len(sys.argv) > 2 else "multi"
mk = {"planar": problem.planar_test_particle_input, "relativistic": problem.relativistic_input,
      "nonlinear": problem.nonlinear_input, "multi": problem.multi_species_input}[workload]
inp = mk(n)
inp.num_iterations = 1
run = problem.setup_run(inp)
lib = mcs_b200.load_cuda_library()
eng = abi.Engine(lib, driver.make_config(lib, run, na_cr=1_000_000))
prof = run.profile
eps = problem.populate_eps_target(run, prof)
pool = np.zeros(run.n_grid)
for rep in range(2):
    pool[:] = 0
    for i_ion in range(1, run.n_ions + 1):
        sp = run.species[i_ion - 1]
        spec = problem.injection_spec(run, prof, i_ion)
        eng.set_profile(prof, eps, pool.copy())
        eng.timing(reset=True)
        eng.begin_ion_generate(1, i_ion, driver.species_struct(run, i_ion), spec, shuffle=True)
        p_hi = problem.pcut_hi(inp.en_pcut_hi, sp.mass)
        n_run, n_used, n_saved = eng.run_ion(run.pcuts, p_hi, inp.n_pts_pcut, inp.n_pts_pcut_hi)
        t = eng.end_ion(want_psd=False, want_log=False)
        tm = eng.timing()
        pool = pool + t.energy_transfer_pool
        steps = t.stats["n_helix_steps"] + t.stats["n_retro_steps"]
        if rep == 1:
            print(f"ion {i_ion} aa={sp.aa:.3g} zz={sp.charge:.3g}: pcuts {n_run} steps {steps:.3e} "
                  f"kernel {tm['transport_ms']:.0f} ms  -> {steps / (tm['transport_ms'] * 1e-3):.3e} steps/s  fates {t.stats['n_fate']}")
