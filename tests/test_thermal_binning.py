"""device-side forms of what the host consumers build from the thermal-crossing log.

CPU: the oracle's on-the-fly histograms equal a post-hoc numpy binning of its own crossing log with the consumers'
formulas (particle_counter.jl:426-445 shock frame; thermo_calcs.jl:133-164 plasma frame; particle_counter.jl:81-85).
GPU: the kernel's histograms against the oracle's."""
import numpy as np
import pytest

from helpers import LADDER, make_engine, rel_close, start_ion
from mcs_b200 import problem


def _bins(run, px, pt):
    """get_psd_bins.jl:16-97, vectorised."""
    M, T = run.num_psd_mom_bins, run.num_psd_theta_bins
    bpd_p, bpd_t = run.inp.num_psd_bins_per_decade
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(pt < run.psd_mom_min, 0, np.trunc(np.log10(pt / run.psd_mom_min) * bpd_p).astype(np.int64) + 1)
        k = np.minimum(k, M)
        pc = -px / pt
        lin = T - np.trunc((pc + 1) / run.delta_cos).astype(np.int64)
        th = np.arccos(np.clip(pc, -1, 1))
        lg = np.where(th < run.psd_theta_min, 0, np.trunc(np.log10(np.maximum(th, 1e-300) / run.psd_theta_min) * bpd_t).astype(np.int64) + 1)
        j = np.minimum(np.where(pc < run.psd_cos_fine, lin, lg), T)
    return k, j


def _run(lib, run, n_cut=3):
    e = make_engine(lib, run, bin_thermal=True, na_cr=4_000_000)
    start_ion(e, run)
    for k, pcut in enumerate(run.pcuts[:n_cut], start=1):
        ns, _ = e.run_pcut(k, pcut, run.pcuts[k - 2] if k > 1 else 0.0)
        if ns == 0:
            break
        e.split(run.inp.n_pts_pcut)
    return e.end_ion()


@pytest.mark.parametrize("cfg", ["planar", "relativistic"])
def test_oracle_histograms_equal_binned_log(olib, cfg):
    mk = {"planar": lambda: problem.planar_test_particle_input(500, momentum_cutoffs=LADDER[:4]),
          "relativistic": lambda: problem.relativistic_input(300, momentum_cutoffs=problem.DEFAULT_PCUTS[:7])}[cfg]
    run = problem.setup_run(mk())
    t = _run(olib, run)
    assert t.n_cr_overflow == 0 and len(t.therm_grid) > 1000
    ng, M2, T2 = run.n_grid, run.num_psd_mom_bins + 2, run.num_psd_theta_bins + 2
    i = t.therm_grid - 1
    px, pt, w = t.therm_px_sk, t.therm_ptot_sk, t.therm_weight
    k, j = _bins(run, px, pt)
    sf = np.zeros((ng, M2, T2))
    np.add.at(sf, (i, k, j), w)
    assert np.array_equal(sf != 0, t.therm_d2N_sf != 0) and rel_close(sf, t.therm_d2N_sf, 0) < 1e-12
    # plasma frame of the zone whose boundary was crossed
    sp = run.species[0]
    E0 = sp.mass * problem.CL**2
    g, b = run.profile.gam_sf[t.therm_grid], run.profile.ux_sk[t.therm_grid] / problem.CL
    etot = np.hypot(pt * problem.CL, E0)
    pxX = g * (px - b * etot / problem.CL)
    ptX = np.sqrt((pt**2 - px**2) + pxX**2)
    pxX = np.where(np.abs(pxX) > ptX, np.copysign(ptX, pxX), pxX)
    kX, jX = _bins(run, pxX, ptX)
    pf = np.zeros((ng, M2, T2))
    np.add.at(pf, (i, kX, jX), w)
    # a crossing whose boosted momentum sits within rounding of a bin edge may land one bin over in numpy's log10
    agree = (np.abs(pf - t.therm_d2N_pf) <= 1e-12 * np.maximum(pf, t.therm_d2N_pf)).mean()
    assert agree > 0.9999 and pf.sum() == pytest.approx(t.therm_d2N_pf.sum(), rel=1e-12)
    assert t.therm_d2N_sf.sum() == pytest.approx(w.sum(), rel=1e-12)
    # dN(p) of the cosmic rays in the shock frame = PSD summed over angle
    assert rel_close(t.dNdp_cr_sf, t.psd.sum(axis=1), 0) < 1e-13
    assert np.array_equal(np.bincount(i, minlength=ng), t.num_crossings)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["planar", "relativistic", "nonlinear"])
def test_device_histograms_match_oracle(olib, clib, cfg):
    mk = {"planar": lambda: problem.planar_test_particle_input(1500, momentum_cutoffs=LADDER[:4]),
          "relativistic": lambda: problem.relativistic_input(600, momentum_cutoffs=problem.DEFAULT_PCUTS[:7]),
          "nonlinear": lambda: problem.nonlinear_input(800, momentum_cutoffs=LADDER[:4], num_iterations=1)}[cfg]
    run = problem.setup_run(mk())
    if cfg == "nonlinear":
        run.profile = problem.synthetic_precursor(run)
    a, b = _run(olib, run), _run(clib, run)
    for nm in ("therm_d2N_sf", "therm_d2N_pf", "dNdp_cr_sf"):
        x, y = getattr(a, nm), getattr(b, nm)
        assert (x.sum() > 0 or nm == "dNdp_cr_sf") and np.array_equal(x != 0, y != 0), nm
        assert rel_close(x, y, 0) < 1e-9, nm
    assert b.therm_d2N_sf.sum() == pytest.approx(b.therm_weight.sum(), rel=1e-11)


@pytest.mark.gpu
def test_binning_survives_log_overflow(clib):
    """With a tiny log (the 1e7-particle case in miniature) the histograms still hold every crossing."""
    run = problem.setup_run(problem.planar_test_particle_input(20_000, momentum_cutoffs=LADDER[:2]))
    e = make_engine(clib, run, bin_thermal=True, na_cr=100)
    start_ion(e, run)
    e.run_pcut(1, run.pcuts[0], 0.0)
    t = e.end_ion()
    assert t.n_cr_overflow > 0 and t.stats["n_cr_count"] == 100
    full = make_engine(clib, run, bin_thermal=True, na_cr=5_000_000)
    start_ion(full, run)
    full.run_pcut(1, run.pcuts[0], 0.0)
    tf = full.end_ion()
    assert tf.n_cr_overflow == 0
    assert rel_close(t.therm_d2N_sf, tf.therm_d2N_sf, 0) < 1e-12
    assert t.therm_d2N_sf.sum() == pytest.approx(tf.therm_weight.sum(), rel=1e-11)
