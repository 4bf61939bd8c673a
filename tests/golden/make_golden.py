"""Regenerate tests/golden/*.npz.

IMPORTANT: these vectors are produced by the CPU ORACLE (oracle/mcs_oracle.c), not by the reference:
abhro/MonteCarloScattering.jl ships no golden vectors and no Julia runtime exists in this image
(SURVEY.md 8c), so parity is "unpinned" by reference outputs.  The fixtures freeze the oracle's
behaviour (pinned by tests/test_oracle_known_answers.py) so that both the oracle and the CUDA kernel
are regression-checked against the same committed numbers.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(HERE)):
    sys.path.insert(0, p)

import oracle_engine  # noqa: E402
from helpers import LADDER, make_engine, start_ion  # noqa: E402
from mcs_b200 import abi, problem  # noqa: E402

CASES = {
    "planar_small": (lambda: problem.planar_test_particle_input(300, momentum_cutoffs=LADDER[:4]), 1),
    "relativistic_small": (lambda: problem.relativistic_input(200, momentum_cutoffs=problem.DEFAULT_PCUTS[:7]), 1),
    "electrons_small": (lambda: problem.multi_species_input(150, momentum_cutoffs=problem.DEFAULT_PCUTS[:5]), 3),
    "bundled": (problem.bundled_input, 1),
}


def run_case(lib, inp, i_ion):
    """Returns a flat dict of arrays: per-pcut fates + saved state, and the per-ion tallies."""
    run = problem.setup_run(inp)
    e = make_engine(lib, run)
    pool = None
    if i_ion > 1:  # electrons read the pool the ions donated (main_loops.jl:164): run ion 1 first
        start_ion(e, run, i_ion=1)
        e.run_ion(run.pcuts, problem.pcut_hi(inp.en_pcut_hi, run.species[0].mass), inp.n_pts_pcut, inp.n_pts_pcut_hi)
        pool = e.end_ion(want_psd=False, want_log=False).energy_transfer_pool
    start_ion(e, run, i_ion=i_ion, pool=pool)
    sp = run.species[i_ion - 1]
    p_hi = problem.pcut_hi(inp.en_pcut_hi, sp.mass)
    out = {}
    for k, pcut in enumerate(run.pcuts, start=1):
        n = e.population_size()
        ns, steps = e.run_pcut(k, pcut, run.pcuts[k - 2] if k > 1 else 0.0)
        f = e.get_fates(n)
        s = e.get_population(1, n)
        out[f"p{k}_n"] = np.array([n, ns, steps])
        for nm, v in f.items():
            out[f"p{k}_{nm}"] = v
        for nm, v in s.items():
            out[f"p{k}_saved_{nm}"] = v
        if ns == 0:
            break
        e.split(inp.n_pts_pcut if pcut < p_hi else inp.n_pts_pcut_hi)
    t = e.end_ion()
    for nm in ("pxx_flux", "pxz_flux", "energy_flux", "num_crossings", "esc_energy_eff", "esc_num_eff",
               "weight_coupled", "energy_transfer_pool"):
        out["t_" + nm] = getattr(t, nm)
    for nm in ("psd", "esc_psd_feb_upstream", "esc_psd_feb_downstream", "spectra_coupled"):
        a = getattr(t, nm).ravel()
        idx = np.nonzero(a)[0]
        out[f"t_{nm}_idx"], out[f"t_{nm}_val"] = idx, a[idx]
    k = np.lexsort((t.therm_weight, t.therm_ptot_sk, t.therm_px_sk, t.therm_grid))
    out["t_log_grid"], out["t_log_px"], out["t_log_w"] = t.therm_grid[k][:5000], t.therm_px_sk[k][:5000], t.therm_weight[k][:5000]
    out["t_scalars"] = np.array([t.scalars[k2] for k2 in sorted(t.scalars)])
    out["t_stats"] = np.array([t.stats[k2] for k2 in sorted(t.stats) if k2 != "n_fate"] + t.stats["n_fate"])
    return out


if __name__ == "__main__":
    lib = oracle_engine.load_oracle_library()
    for name, (mk, ion) in CASES.items():
        d = run_case(lib, mk(), ion)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: v for k, v in d.items() if k.endswith("_n")}, os.path.getsize(os.path.join(HERE, name + ".npz")))
