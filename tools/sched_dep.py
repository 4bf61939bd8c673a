"""Diagnostic: do per-particle outcomes depend on the particle-to-warp schedule?  Runs the same population with the static
and the dynamic schedule (enough particles for lanes to be refilled) and reports which particles differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mcs_b200
from mcs_b200 import abi, driver, problem
from helpers import start_ion

wl = sys.argv[1] if len(sys.argv) > 1 else "relativistic"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
npc = int(sys.argv[3]) if len(sys.argv) > 3 else 8
mk = {"planar": problem.planar_test_particle_input, "relativistic": problem.relativistic_input, "nonlinear": problem.nonlinear_input}[wl]
run = problem.setup_run(mk(n, num_iterations=1))
prof = problem.synthetic_precursor(run) if wl == "nonlinear" else run.profile
run.pcuts = run.pcuts[:npc]
lib = mcs_b200.load_cuda_library()
res = []
for dyn in (0, 1):
    cfg = driver.make_config(lib, run, na_cr=1000, n_pts_cap=n + 8)
    cfg.dynamic_queue = dyn
    e = abi.Engine(lib, cfg)
    start_ion(e, run, prof=prof)
    out = []
    for k, pcut in enumerate(run.pcuts, start=1):
        m = e.population_size()
        ns, st = e.run_pcut(k, pcut, run.pcuts[k - 2] if k > 1 else 0.0)
        out.append((m, ns, st, e.get_fates(m), e.get_population(1, m)))
        if ns == 0:
            break
        e.split(run.inp.n_pts_pcut)
    res.append(out)
for k, (a, b) in enumerate(zip(*res), start=1):
    fa, fb = a[3], b[3]
    d = np.nonzero((fa["fate"] != fb["fate"]) | (fa["helix_count"] != fb["helix_count"]) | (fa["n_draws"] != fb["n_draws"]))[0]
    print(f"pcut {k}: n {a[0]} saved {a[1]}/{b[1]} steps {a[2]}/{b[2]} particles that differ: {len(d)}")
    for i in d[:6]:
        print("   ", i, {kk: (int(fa[kk][i]), int(fb[kk][i])) for kk in ("fate", "helix_count", "retro_steps", "n_draws")})
    if len(d):
        sa, sb = a[4], b[4]
        both = (sa["l_save"] == 1) & (sb["l_save"] == 1)
        for nm in ("ptot_pf", "pb_pf", "x_cm", "phi_rad"):
            sc = np.maximum(np.abs(sa["ptot_pf"][both]) if nm != "x_cm" else np.abs(sa["x_cm"][both]), 1e-300)
            if nm == "phi_rad": sc = 2 * np.pi
            print("    saved", nm, "max scaled diff", float(np.max(np.abs(sa[nm][both] - sb[nm][both]) / sc)))
        break
