"""Shared helpers for the parity tests."""
import numpy as np

import mcs_b200
from mcs_b200 import abi, driver, problem

LADDER = [0.01, 0.04, 0.06, 0.09, 0.13, 0.2, 0.3, 0.45, 0.6, 1.0, 1.6, 2.5]  # m_p c; suits a 1e4 km/s shock


def small_inputs():
    """name -> ShockInput: the BASELINE.json configs at sizes the oracle finishes in seconds."""
    return {
        "bundled": problem.bundled_input(),
        "planar": problem.planar_test_particle_input(1500, momentum_cutoffs=LADDER[:6]),
        "relativistic": problem.relativistic_input(600, momentum_cutoffs=problem.DEFAULT_PCUTS[:9]),
        "nonlinear": problem.nonlinear_input(800, momentum_cutoffs=LADDER[:5], num_iterations=1),
        "multi": problem.multi_species_input(500, momentum_cutoffs=problem.DEFAULT_PCUTS[:6]),
        # feature coverage beyond the five headline configs (each switches on one more branch of particle_loop)
        "tcuts_age": problem.planar_test_particle_input(800, momentum_cutoffs=LADDER[:4], maximum_age=3.0e4,
                                                        tcuts=[1e2, 3e2, 1e3, 3e3, 1e4, 1e6]),
        "xspec": problem.planar_test_particle_input(600, momentum_cutoffs=LADDER[:3], x_spec=[-0.5, -0.05, 0.2, 3.0]),
        "injfrac": problem.planar_test_particle_input(800, momentum_cutoffs=LADDER[:4], inj_fracs=[0.3]),
        "feb_down": problem.planar_test_particle_input(800, momentum_cutoffs=LADDER[:4], feb_downstream=(4.0, 0.0)),
        "bundled_scatter": problem.ShockInput(no_scatter=False, no_dsa=False, n_pts_inj=300, n_pts_pcut=400, n_pts_pcut_hi=400,
                                              momentum_cutoffs=problem.DEFAULT_PCUTS[:6]),
        "no_retro_error": problem.planar_test_particle_input(300, momentum_cutoffs=LADDER[:2], use_retro=False),
    }


def make_engine(lib, run, **kw):
    kw.setdefault("na_cr", 3_000_000)
    return abi.Engine(lib, driver.make_config(lib, run, **kw))


def start_ion(eng, run, i_ion=1, i_iter=1, prof=None, pool=None, pop=None, first_global=0):
    prof = run.profile if prof is None else prof
    eps = problem.populate_eps_target(run, prof)
    if pop is None:
        pop = problem.init_pop(run, prof, i_ion, np.random.default_rng((i_iter - 1) * run.n_ions + i_ion - 1)).pop
    eng.set_profile(prof, eps, np.zeros(run.n_grid) if pool is None else pool)
    eng.begin_ion(i_iter, i_ion, driver.species_struct(run, i_ion), pop, first_global=first_global)
    return pop


def natural_scales(run, sp, pop):
    """Magnitudes against which "1e-12 relative" is measured for each continuous field: a component of a
    vector is compared on the scale of the vector (pb vs ptot), a position on max(|x|, gyroradius),
    an angle on 2 pi."""
    ptot = np.abs(pop["ptot_pf"])
    rg = ptot * problem.CL / (abs(sp.charge) * run.bmag0)
    return {
        "weight": np.abs(pop["weight"]), "ptot_pf": ptot, "pb_pf": ptot,
        "x_cm": np.maximum(np.abs(pop["x_cm"]), rg), "xn_per": np.abs(pop["xn_per"]),
        "prp_x_cm": np.maximum(np.abs(pop["prp_x_cm"]), rg), "acctime_sec": None, "phi_rad": np.full_like(ptot, 2 * np.pi),
    }


def compare_saved(run, sp, a, b, tol, tol_phi=None):
    """a, b: get_population(1, n) of oracle and device. Returns dict field -> max scaled difference."""
    assert np.array_equal(a["l_save"], b["l_save"]), "l_save differs"
    m = a["l_save"].astype(bool)
    out = {}
    for nm in abi.POP_I64 + abi.POP_U8:
        assert np.array_equal(a[nm][m], b[nm][m]), f"integer field {nm} differs"
    sc = natural_scales(run, sp, {k: v[m] for k, v in a.items()})
    for nm in abi.POP_F64:
        x, y = a[nm][m], b[nm][m]
        s = sc[nm]
        if s is None:
            s = np.maximum(np.abs(x), 1e-300)
        s = np.where(s > 0, s, 1.0)
        out[nm] = float(np.max(np.abs(x - y) / s)) if x.size else 0.0
        lim = tol_phi if (nm == "phi_rad" and tol_phi is not None) else tol
        assert out[nm] <= lim, f"{nm}: scaled difference {out[nm]:.3e} > {lim:g}"
    return out


def rel_close(a, b, rtol, atol_frac=0.0):
    """max |a-b| / max(|a|,|b|) over cells, ignoring cells below atol_frac * max|a|."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    s = np.maximum(np.abs(a), np.abs(b))
    floor = atol_frac * (np.abs(a).max() if a.size else 0.0)
    m = s > floor
    if not m.any():
        return 0.0
    return float((np.abs(a - b)[m] / s[m]).max())


def sorted_log(t):
    k = np.lexsort((t.therm_weight, t.therm_ptot_sk, t.therm_px_sk, t.therm_grid))
    return t.therm_grid[k], t.therm_px_sk[k], t.therm_ptot_sk[k], t.therm_weight[k]
