"""the pressure consumer `mcs_thermo` (thermo_calcs, /root/reference/src/thermo_calcs.jl:31-355).

CPU: the oracle against (a) a second, cell-by-cell restatement written here with the reference's loop structure, on the
tallies of a real run, and (b) hand-computed answers for the three normalisation cases of :242-300 on synthetic cells.
GPU: the kernel against the oracle, on identical host arrays (rounding-level agreement) and on the tallies each backend
holds after its own run of the same population."""
import math

import numpy as np
import pytest

from helpers import LADDER, make_engine, rel_close, start_ion
from mcs_b200 import abi, problem

KB = 1.380649e-16


def _bin_mom(run, pt):
    bpd_p = run.inp.num_psd_bins_per_decade[0]
    b = 0 if pt < run.psd_mom_min else int(math.trunc(math.log10(pt / run.psd_mom_min) * bpd_p)) + 1
    return min(b, run.num_psd_mom_bins)


def _bin_ang(run, px, pt):
    T, bpd_t = run.num_psd_theta_bins, run.inp.num_psd_bins_per_decade[1]
    if pt == 0.0:
        return 0
    pc = -px / pt
    if pc < run.psd_cos_fine:
        b = T - int(math.trunc((pc + 1) / run.delta_cos))
    else:
        th = math.acos(pc)
        b = 0 if th < run.psd_theta_min else int(math.trunc(math.log10(th / run.psd_theta_min) * bpd_t)) + 1
    return min(b, T)


def thermo_restated(run, prof, sp, cosc, ptc, zone_pop, T0, psd, thpf, ncross):
    """thermo_calcs.jl:166-352 zone by zone; psd [ng, T+2, M+2], thpf [ng, M+2, T+2] as abi.Tallies holds them."""
    ng, T, M = run.n_grid, run.num_psd_theta_bins, run.num_psd_mom_bins
    c, m = problem.CL, sp.mass
    mc, E0 = m * c, m * c * c
    out = np.zeros((4, ng))
    for i in range(1, ng + 1):
        g, b = float(prof.gam_sf[i]), float(prof.ux_sk[i]) / c
        d2N = 1.0e-99 + thpf[i - 1].T.copy()                      # [jth, k]
        for jt, k in zip(*np.nonzero(psd[i - 1][: T + 1, : M + 1] > 1.0e-66)):
            pt = float(ptc[k]); px = pt * float(cosc[jt])
            etot = math.hypot(pt * c, E0)
            pxX = g * (px - b * etot / c)
            ptX = math.sqrt(pt * pt - px * px + pxX * pxX)
            d2N[_bin_ang(run, pxX, ptX), _bin_mom(run, ptX)] += psd[i - 1][jt, k]
        nf = d2N[d2N > 1.0e-66].sum()
        if ncross[i - 1] == 0 and nf > 0:
            nf += sp.n0 / float(prof.ux_sk[i])
        if nf > 0:
            nf = zone_pop[i - 1] / nf
        d2N[d2N > 1.0e-66] *= nf
        pop = d2N[d2N > 1.0e-66].sum()
        dens = run.gam0 * run.beta0 * sp.n0 / math.sqrt(g * g - 1)
        par = perp = en = 0.0
        if d2N.max() < 1.0e-66 and ncross[i - 1] == 0:
            pl = dens ** (5 / 3) * KB * T0
            out[:, i - 1] = pl / 3, 2 * pl / 3, 1.5 * pl, pop
            continue
        if ncross[i - 1] == 0:
            pl = dens ** (5 / 3) * KB * T0 * (1 - pop / zone_pop[i - 1])
            par, perp, en = pl / 3, 2 * pl / 3, 1.5 * pl
        norm = dens / zone_pop[i - 1] if zone_pop[i - 1] != 0 else math.inf
        gt = np.hypot(1.0, ptc / mc)
        vel = ptc * c / (mc * gt)
        cells = d2N[: T + 1, : M + 1]
        live = cells >= 1.0e-66
        pf = ptc * vel * norm / 3
        c2 = np.broadcast_to((cosc**2)[:, None], cells.shape)
        pfb, efb = np.broadcast_to(pf[None, :], cells.shape), np.broadcast_to(((gt - 1) * E0)[None, :], cells.shape)
        par += float((cells[live] * pfb[live] * c2[live]).sum()); perp += float((cells[live] * pfb[live] * (1 - c2[live])).sum())
        en += float((efb[live] * cells[live] * norm).sum())
        out[:, i - 1] = par, perp, en, pop
    return out


def _ion_tallies(lib, run, n_cut=4, prof=None):
    e = make_engine(lib, run, bin_thermal=True, na_cr=4_000_000)
    start_ion(e, run, prof=prof)
    for k, pcut in enumerate(run.pcuts[:n_cut], start=1):
        ns, _ = e.run_pcut(k, pcut, run.pcuts[k - 2] if k > 1 else 0.0)
        if ns == 0:
            break
        e.split(run.inp.n_pts_pcut)
    return e, e.end_ion(want_log=False)


def _case(cfg):
    mk = {"planar": lambda: problem.planar_test_particle_input(800, momentum_cutoffs=LADDER[:5], fixed_grid=True),
          "as_written_grid": lambda: problem.planar_test_particle_input(500, momentum_cutoffs=LADDER[:4]),
          "relativistic": lambda: problem.relativistic_input(400, momentum_cutoffs=problem.DEFAULT_PCUTS[:8], fixed_grid=True),
          "nonlinear": lambda: problem.nonlinear_input(600, momentum_cutoffs=LADDER[:4], num_iterations=1, fixed_grid=True)}[cfg]
    run = problem.setup_run(mk())
    prof = problem.synthetic_precursor(run) if cfg == "nonlinear" else run.profile
    return run, prof


@pytest.mark.parametrize("cfg", ["planar", "as_written_grid", "relativistic"])
def test_oracle_thermo_equals_restatement(olib, cfg):
    run, prof = _case(cfg)
    sp = run.species[0]
    e, t = _ion_tallies(olib, run, n_cut=8 if cfg == "relativistic" else 4, prof=prof)
    cosc, ptc, zp = problem.thermo_inputs(run, prof, 0)
    got = np.array(e.thermo(cosc, ptc, zp, sp.T))
    want = thermo_restated(run, prof, sp, cosc, ptc, zp, sp.T, t.psd, t.therm_d2N_pf, t.num_crossings)
    for k, nm in enumerate(("P_psd_par", "P_psd_perp", "energy_density_psd", "d2N_pop")):
        assert np.all(np.isfinite(got[k])), nm
        assert rel_close(got[k], want[k], 0) < 1e-12, nm
    # the same call on explicit host arrays
    again = np.array(e.thermo(cosc, ptc, zp, sp.T, psd=t.psd, therm_d2N_pf=t.therm_d2N_pf, num_crossings=t.num_crossings))
    assert np.array_equal(got, again)
    if cfg == "planar":
        # zones the thermal gas crossed hold the whole population (case 3): d2N_pop = zone_pop, pressures positive
        seen = t.num_crossings > 0
        assert seen.sum() > run.n_grid // 2
        assert rel_close(got[3][seen], zp[seen], 0) < 1e-12
        assert np.all(got[0][seen] > 0) and np.all(got[1][seen] > 0) and np.all(got[2][seen] > 0)


def test_three_normalisation_cases_by_hand(olib):
    """Zone 1: nothing seen (case 1, :250-271).  Zone 2: one CR cell, no thermal crossing (case 2, :273-297).
    Zone 3: one thermal cell (case 3, :299-306).  Cells sit where the boost cannot move them across a bin edge check:
    the expected values are computed from the bin the restated boost lands in."""
    run = problem.setup_run(problem.planar_test_particle_input(100, momentum_cutoffs=LADDER[:2], fixed_grid=True))
    prof, sp = run.profile, run.species[0]
    ng, T2, M2 = run.n_grid, run.num_psd_theta_bins + 2, run.num_psd_mom_bins + 2
    e = make_engine(olib, run, bin_thermal=True)
    start_ion(e, run)
    cosc, ptc, zp = problem.thermo_inputs(run, prof, 0)
    psd, thp, ncr = np.zeros((ng, T2, M2)), np.zeros((ng, M2, T2)), np.zeros(ng, np.int64)
    psd[1, 30, 40] = 7.0
    thp[2, 12, 50] = 3.0
    ncr[2] = 5
    par, perp, en, pop = e.thermo(cosc, ptc, zp, sp.T, psd=psd, therm_d2N_pf=thp, num_crossings=ncr)
    c, mc, E0 = problem.CL, sp.mass * problem.CL, sp.mass * problem.CL**2
    dens = lambda i: run.gam0 * run.beta0 * sp.n0 / math.sqrt(float(prof.gam_sf[i]) ** 2 - 1)
    # case 1
    pl = dens(1) ** (5 / 3) * KB * sp.T
    assert (par[0], perp[0], en[0]) == pytest.approx((pl / 3, 2 * pl / 3, 1.5 * pl), rel=1e-14) and pop[0] == 0
    # case 2: the cell is scaled by zone_pop / (7 + n0/ux), boosted to the plasma frame of zone 2
    g, b = float(prof.gam_sf[2]), float(prof.ux_sk[2]) / c
    pt = ptc[40]; px = pt * cosc[30]
    pxX = g * (px - b * math.hypot(pt * c, E0) / c); ptX = math.sqrt(pt * pt - px * px + pxX * pxX)
    kX, jX = _bin_mom(run, ptX), _bin_ang(run, pxX, ptX)
    d = 7.0 * zp[1] / (7.0 + sp.n0 / float(prof.ux_sk[2]))
    assert pop[1] == pytest.approx(d, rel=1e-14)
    pl = dens(2) ** (5 / 3) * KB * sp.T * (1 - d / zp[1])
    norm = dens(2) / zp[1]
    gt = math.hypot(1, ptc[kX] / mc); vel = ptc[kX] * c / (mc * gt)
    pf = ptc[kX] * vel * norm / 3
    assert par[1] == pytest.approx(pl / 3 + d * pf * cosc[jX] ** 2, rel=1e-13)
    assert perp[1] == pytest.approx(2 * pl / 3 + d * pf * (1 - cosc[jX] ** 2), rel=1e-13)
    assert en[1] == pytest.approx(1.5 * pl + (gt - 1) * E0 * d * norm, rel=1e-13)
    # case 3: the thermal cell is already in the plasma frame; it becomes the whole zone population
    assert pop[2] == pytest.approx(zp[2], rel=1e-14)
    norm = dens(3) / zp[2]
    gt = math.hypot(1, ptc[12] / mc); vel = ptc[12] * c / (mc * gt)
    pf = ptc[12] * vel * norm / 3
    assert par[2] == pytest.approx(zp[2] * pf * cosc[50] ** 2, rel=1e-13)
    assert perp[2] == pytest.approx(zp[2] * pf * (1 - cosc[50] ** 2), rel=1e-13)
    assert en[2] == pytest.approx((gt - 1) * E0 * zp[2] * norm, rel=1e-13)
    # every other zone is case 1
    rest = np.arange(3, ng)
    assert np.all(pop[rest] == 0) and np.all(par[rest] > 0) and np.allclose(perp[rest], 2 * par[rest], rtol=1e-14)


def test_thermo_argument_errors(olib):
    run = problem.setup_run(problem.planar_test_particle_input(100, momentum_cutoffs=LADDER[:2]))
    cosc, ptc, zp = problem.thermo_inputs(run, run.profile, 0)
    e = make_engine(olib, run)                      # bin_thermal off
    start_ion(e, run)
    with pytest.raises(abi.McsError, match="bin_thermal"):
        e.thermo(cosc, ptc, zp, 1e6)
    e = make_engine(olib, run, bin_thermal=True)
    start_ion(e, run)
    with pytest.raises(abi.McsError, match="mcs_end_ion first"):
        e.thermo(cosc, ptc, zp, 1e6)                # tallies of an ion still in flight
    with pytest.raises(abi.McsError, match="together"):
        e.thermo(cosc, ptc, zp, 1e6, psd=np.zeros((run.n_grid, run.num_psd_theta_bins + 2, run.num_psd_mom_bins + 2)))
    with pytest.raises(abi.McsError, match="cos_center"):
        e.thermo(cosc[:-1], ptc, zp, 1e6)
    e.end_ion()
    e.thermo(cosc, ptc, zp, 1e6)


def test_thermo_inputs_follow_reference_bin_centres():
    run = problem.setup_run(problem.planar_test_particle_input(100))
    mom, th = problem.psd_bounds(run, as_written=False)
    T, M, lin = run.num_psd_theta_bins, run.num_psd_mom_bins, run.inp.psd_linear_cosine_bins
    assert len(mom) == M + 2 and len(th) == T + 2 and mom[0] == -99.0 and th[0] == 1e-99
    assert th[T - lin + 1] == run.psd_cos_fine and th[-1] == pytest.approx(-1.0, abs=1e-12)
    assert np.all(np.diff(th[: T - lin + 1]) > 0) and np.all(np.diff(th[T - lin + 1:]) < 0)   # the docstring's warning
    assert np.all(np.diff(problem.psd_bounds(run)[1]) >= 0)                                   # sort! as written
    cosc, ptc, zp = problem.thermo_inputs(run, run.profile, 0)
    # finest gradations point upstream: cos_center runs from -1 (bin 0, theta ~ 0 flipped) to +1
    assert cosc[0] == pytest.approx(-1.0, abs=1e-6) and cosc[-1] == pytest.approx(1 - run.delta_cos / 2, rel=1e-12)
    assert np.all(np.diff(cosc) >= 0)
    # bin k >= 1 spans [psd_mom_min 10^((k-1)/bpd), psd_mom_min 10^(k/bpd)): its centre bins back to k
    for k in (1, 7, M - 1):
        assert _bin_mom(run, ptc[k]) == k
    raw = problem.thermo_inputs(run, run.profile, 0, as_written=True)[1]
    assert np.allclose(raw * problem.MP * problem.CL, ptc, rtol=1e-15)
    assert zp.shape == (run.n_grid,)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["planar", "relativistic", "nonlinear"])
def test_device_thermo_matches_oracle(olib, clib, cfg):
    run, prof = _case(cfg)
    sp = run.species[0]
    n_cut = 8 if cfg == "relativistic" else 4      # the first cosmic rays of the gamma0 = 10 ladder appear at pcut 5
    eo, to = _ion_tallies(olib, run, n_cut=n_cut, prof=prof)
    ec, tc = _ion_tallies(clib, run, n_cut=n_cut, prof=prof)
    cosc, ptc, zp = problem.thermo_inputs(run, prof, 0)
    want = np.array(eo.thermo(cosc, ptc, zp, sp.T))
    # (a) identical inputs: only the order of the FP64 adds differs
    same_in = np.array(ec.thermo(cosc, ptc, zp, sp.T, psd=to.psd, therm_d2N_pf=to.therm_d2N_pf, num_crossings=to.num_crossings))
    # (b) the tallies the device holds after its own run of the same population
    resident = np.array(ec.thermo(cosc, ptc, zp, sp.T))
    assert (to.num_crossings > 0).sum() > 10 and (to.psd > 0).sum() > 100
    for k, nm in enumerate(("P_psd_par", "P_psd_perp", "energy_density_psd", "d2N_pop")):
        assert np.all(np.isfinite(same_in[k])), nm
        # energy_density sums (hypot(1, p/mc) - 1) E0: for p/mc ~ 1e-3 one ulp of the device hypot against glibc's is
        # 1e-16 / 5e-7 of the term — the reference's own formula is that ill-conditioned
        assert rel_close(same_in[k], want[k], 0) < (1e-8 if k == 2 else 1e-12), nm
        assert rel_close(resident[k], want[k], 0) < (1e-8 if k == 2 else 1e-9), nm
    # run to run on the device: the re-binning uses FP64 atomics, so agreement to rounding, not bitwise
    again = np.array(ec.thermo(cosc, ptc, zp, sp.T))
    assert rel_close(again, resident, 0) < 1e-13


@pytest.mark.gpu
def test_device_thermo_state_errors(clib):
    run = problem.setup_run(problem.planar_test_particle_input(200, momentum_cutoffs=LADDER[:2]))
    cosc, ptc, zp = problem.thermo_inputs(run, run.profile, 0)
    e = make_engine(clib, run, bin_thermal=True)
    start_ion(e, run)
    with pytest.raises(abi.McsError, match="mcs_end_ion first"):
        e.thermo(cosc, ptc, zp, 1e6)
    e2 = make_engine(clib, run)
    start_ion(e2, run)
    with pytest.raises(abi.McsError, match="bin_thermal"):
        e2.thermo(cosc, ptc, zp, 1e6)


def test_thermo_with_the_reference_bin_centres_as_written(olib):
    """`thermo_inputs(as_written=True)`: theta bounds sorted (initializers.jl:283) and pt_center = exp10(log10(p / m_p c)) taken
    as g cm/s (thermo_calcs.jl:77-82) — momenta 1/(m_p c) = 2e13 too large, so a bin centre lands ~130 bins above its own bin.
    The library does what it is given: finite output, equal to the restatement."""
    run, prof = _case("planar")
    sp = run.species[0]
    e, t = _ion_tallies(olib, run, prof=prof)
    cosc, ptc, zp = problem.thermo_inputs(run, prof, 0, as_written=True)
    assert ptc[1] > 1e10 * run.psd_mom_min and np.all(np.diff(problem.psd_bounds(run, as_written=True)[1]) >= 0)
    got = np.array(e.thermo(cosc, ptc, zp, sp.T))
    want = thermo_restated(run, prof, sp, cosc, ptc, zp, sp.T, t.psd, t.therm_d2N_pf, t.num_crossings)
    assert np.all(np.isfinite(got))
    for k in range(4):
        assert rel_close(got[k], want[k], 0) < 1e-12
    for k in (1, 3, 10):
        assert _bin_mom(run, float(ptc[k])) > k + 100
