"""Host-side input producers against the numbers SURVEY.md section 8 derives from the reference source."""
import math

import numpy as np
import pytest

from mcs_b200 import problem


def test_bundled_run_shapes():
    run = problem.setup_run(problem.bundled_input())
    assert run.n_grid == 99                      # 1+27+35+5+16+16+1 = 101 nodes
    assert run.i_grid_feb == 27                  # FEB = -100 rg0
    assert run.i_shock == 65 and run.profile.x_grid_rg[65] == 0.0
    assert run.num_psd_mom_bins == 171 and run.num_psd_theta_bins == 159
    assert run.rg0 == pytest.approx(1.5333e12, rel=1e-4)
    assert run.psd_mom_min / (problem.MP * problem.CL) == pytest.approx(1.0002e-6, rel=1e-3)
    assert run.gam0 == 5.0 and run.do_retro and run.do_tcuts and len(run.pcuts) == 45
    assert run.r_comp == pytest.approx(4.0, rel=1e-3)   # as-written calc_rRH (SURVEY B-12)


def test_psd_bins_other_species_lists():
    a = problem.setup_run(problem.ShockInput(aa_ion=[1.0], zz_ion=[1.0], tz_ion=[1e6], denz_ion=[1.0]))
    assert a.num_psd_mom_bins == 155             # protons only, gamma0 = 5
    b = problem.setup_run(problem.ShockInput(shock_speed=10.0))
    assert b.num_psd_mom_bins == 175             # with electrons, gamma0 = 10


def test_grid_as_written_is_non_monotonic_and_fixed_is_not():
    g, _, _ = problem.setup_grid(-1e7, 10.0, True, 0.0, 1.0)
    assert len(g) == 101 and g[0] == -1e30 and g[-1] == 1e30
    assert g[1] == pytest.approx(-1e7) and g[27] == pytest.approx(-1.7e27, rel=0.05) and g[28] == -9.0
    assert not np.all(np.diff(g) >= 0)
    assert g[83] == 1.0 and g[84] == pytest.approx(1.0)   # downstream log block repeats x = 1
    f, _, _ = problem.setup_grid(-1e7, 10.0, True, 0.0, 1.0, fixed=True)
    assert np.all(np.diff(f) > 0) and f[-2] == pytest.approx(10.0)


def test_profile_unmodified_shock():
    run = problem.setup_run(problem.planar_test_particle_input(1000))
    p = run.profile
    up = p.x_grid_cm < 0
    assert np.all(p.ux_sk[up] == run.u0) and np.all(p.ux_sk[~up] == run.u0 / run.r_comp)
    assert np.all(p.gam_ef[up] == 1.0) and np.all(p.btot == run.bmag0)   # turbulence 0, no custom eps_B
    assert run.u2 == pytest.approx(run.u0 / run.r_comp)
    assert run.beta0 == pytest.approx(1e9 / problem.CL)


def test_inj_dist_equal_weight():
    p, w = problem.set_inj_dist(True, 1000, 1, 1e6, problem.MP, 1.0)
    assert abs(len(p) - 1000) < 15 and np.all(np.diff(p) >= 0)
    assert w.sum() == pytest.approx(1.0) and np.all(w == w[0])
    pz, _ = problem.set_inj_dist(True, 1000, 1, 1e6, problem.MP, 1.0, compat_zero_first=True)
    assert pz[0] == 0.0 and len(pz) == len(p) + 1           # SURVEY B-7
    pb, wb = problem.set_inj_dist(False, 1500, 1, 1e6, problem.MP, 2.0)
    assert len(pb) == 1500 and wb.sum() == pytest.approx(2.0)
    # thermal peak of p^2 exp(-p^2/2mkT) sits at sqrt(2 m k T)
    peak = math.sqrt(2 * problem.MP * problem.KB * 1e6)
    h, e = np.histogram(p, bins=30)
    assert e[np.argmax(h)] < peak < e[np.argmax(h) + 2]


def test_init_pop_fast_push_and_plain():
    run = problem.setup_run(problem.planar_test_particle_input(500))
    ip = problem.init_pop(run, run.profile, 1, np.random.default_rng(0))
    n = len(ip.pop["weight"])
    assert np.all(ip.pop["grid"] == 43) and np.all(ip.pop["x_cm"] == -1.0 * run.rg0)
    assert np.all(np.abs(ip.pop["pb_pf"]) <= ip.pop["ptot_pf"] * (1 + 1e-12))
    assert np.all((ip.pop["phi_rad"] >= 0) & (ip.pop["phi_rad"] < 2 * np.pi))
    assert np.all(ip.pxx_flux[:43] == pytest.approx(run.F_px_upstream, rel=1e-3)) and np.all(ip.pxx_flux[43:] == 0)
    run2 = problem.setup_run(problem.planar_test_particle_input(500, fast_upstream_transport=False))
    ip2 = problem.init_pop(run2, run2.profile, 1, np.random.default_rng(0))
    assert np.all(ip2.pop["grid"] == 0) and np.all(ip2.pxx_flux == 0)
    assert np.all(ip2.pop["x_cm"] == run2.x_grid_start - 10 * run2.rg0) and len(ip2.pop["weight"]) == n


def test_eps_target_and_cutoffs():
    run = problem.setup_run(problem.bundled_input())
    eps = problem.populate_eps_target(run, run.profile)
    assert np.all(eps[:64] == 0) and np.all(eps[65:] == pytest.approx(0.1))   # energy-transfer-frac downstream
    assert problem.get_pmax_cutoff(run, 1.0) == pytest.approx(1e10 * problem.MP * problem.CL)
    assert problem.pcut_hi(1e6, problem.MP) / (problem.MP * problem.CL) == pytest.approx(1.81, rel=0.01)
