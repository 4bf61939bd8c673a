"""Full-run statistical parity at size: CUDA at 1e6 particles per pcut against 8 independent oracle
replicas at 4000 per pcut, on the planar test-particle shock AND on the smoothed (nonlinear) precursor profile."""
import os

import numpy as np
import pytest

import stat_parity
from helpers import make_engine, start_ion
from mcs_b200 import driver, problem

pytestmark = pytest.mark.gpu
P_MIN = 1e-3   # stated p: a spectrum or profile is rejected below this


def _one(lib, run, prof, seed, threads=1, generate=False):
    e = make_engine(lib, run, seed=seed, threads=threads, na_cr=1000, n_pts_cap=max(run.inp.n_pts_pcut, run.inp.n_pts_inj) + 8)
    eps = problem.populate_eps_target(run, prof)
    e.set_profile(prof, eps, np.zeros(run.n_grid))
    sp = driver.species_struct(run, 1)
    if generate:
        e.begin_ion_generate(1, 1, sp, problem.injection_spec(run, prof, 1), shuffle=True)
    else:
        e.begin_ion(1, 1, sp, problem.init_pop(run, prof, 1, np.random.default_rng(seed)).pop)
    e.run_ion(run.pcuts, problem.pcut_hi(run.inp.en_pcut_hi, run.species[0].mass), run.inp.n_pts_pcut, run.inp.n_pts_pcut_hi)
    t = e.end_ion(want_log=False)
    e.close()
    return stat_parity.observables(t, run)


@pytest.mark.parametrize("workload", ["planar", "nonlinear"])
def test_spectra_and_profiles_at_size(olib, clib, workload):
    n_big, n_rep, R = 1_000_000, 4000, 8
    mk = {"planar": problem.planar_test_particle_input, "nonlinear": problem.nonlinear_input}[workload]
    run_big = problem.setup_run(mk(n_big, num_iterations=1))
    run_rep = problem.setup_run(mk(n_rep, num_iterations=1))
    prof = (lambda r: problem.synthetic_precursor(r)) if workload == "nonlinear" else (lambda r: r.profile)
    big = _one(clib, run_big, prof(run_big), seed=210, generate=True)
    cores = os.cpu_count() or 1
    reps = [_one(olib, run_rep, prof(run_rep), seed=5000 + r, threads=cores) for r in range(R)]
    res = stat_parity.compare(big, reps, n_ratio=n_big / n_rep)
    print()
    for k, v in res.items():
        print(f"[{workload}] {k}: {v}")
    checked = 0
    for k, v in res.items():
        if "skipped" in v:
            continue
        checked += 1
        assert v["p_chi2"] > P_MIN, (workload, k, v)
    assert checked >= 3
    # the KS p-values of the z scores against Student-t are printed above; they are not asserted: neighbouring flux zones
    # (and, less so, momentum bins fed by the same long trajectories) are positively correlated
