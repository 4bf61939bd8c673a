"""Pin the CPU oracle against the known answers the reference source itself documents.

The reference ships no golden vectors (test/runtests.jl is Aqua only), so these analytic facts are what
stands between the oracle and "unpinned": each test names the reference line that states the fact.
"""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import make_engine, start_ion
from mcs_b200 import abi, driver, problem

D = C.c_double


@pytest.fixture(scope="module")
def eng(olib):
    run = problem.setup_run(problem.planar_test_particle_input(200))
    e = make_engine(olib, run)
    e.run = run
    for f in ("mcso_radiation_loss", "mcso_mod2pi"):
        getattr(olib, f).restype = D
    olib.mcso_transform_p_PS.argtypes = [C.c_void_p] + [D] * 9 + [C.POINTER(D * 5)]
    olib.mcso_transform_p_PSP.argtypes = [C.c_void_p, D, C.POINTER(D * 5), C.POINTER(D * 6), C.POINTER(D * 6)]
    olib.mcso_scattering.argtypes = [C.c_void_p, C.c_uint32, C.c_int] + [D] * 5 + [C.POINTER(D * 4)]
    olib.mcso_psd_bin_momentum.argtypes = [C.c_void_p, D]
    olib.mcso_psd_bin_angle.argtypes = [C.c_void_p, D, D]
    olib.mcso_radiation_loss.argtypes = [C.c_void_p, D, D, D]
    olib.mcso_mod2pi.argtypes = [D]
    olib.mcso_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(D)]
    return e


def test_philox_random123_known_answers(olib):
    """Philox4x32-10 against the published Random123 kat_vectors (the only external golden vectors on this path)."""
    def ph(ctr, key):
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        olib.mcso_philox_raw(c, k, o)
        return list(o)
    assert ph([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert ph([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert ph([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_stream_is_53_bit_in_unit_interval(eng, olib):
    n = 200_000
    u = np.zeros(n)
    olib.mcso_philox(210, 7, 3, 1, n, u.ctypes.data_as(C.POINTER(D)))
    assert u.min() >= 0.0 and u.max() < 1.0                     # Random.rand: [0,1) (App. E)
    assert np.all(u * 2.0**53 == np.floor(u * 2.0**53))
    assert abs(u.mean() - 0.5) < 4 / math.sqrt(12 * n) and abs(u.var() - 1 / 12) < 2e-3


def test_transform_ps_invariant_and_psp_identity(eng, olib):
    """E^2 - p^2 c^2 is invariant under transform_p_PS (transformers.jl:440-476);
    transform_p_PSP between identical zones is the identity (:523-607)."""
    run, h = eng.run, eng._h
    m, c = problem.MP, problem.CL
    rng = np.random.default_rng(1)
    for gam0 in (1.0005, 1.5, 10.0):
        beta = math.sqrt(1 - 1 / gam0**2)
        ux = beta * c
        for _ in range(20):
            ptot = m * c * 10 ** rng.uniform(-3, 3)
            mu = rng.uniform(-1, 1)
            pb, pperp = ptot * mu, ptot * math.sqrt(1 - mu * mu)
            gam = math.hypot(1, ptot / (m * c))
            phi = rng.uniform(0, 2 * math.pi)
            out = (D * 5)()
            olib.mcso_transform_p_PS(h, 1.0, pb, pperp, gam, phi, ux, gam0, 1.0, 0.0, C.byref(out))
            ptot_sk, sx, sy, sz, gam_sk = out
            assert ptot_sk == pytest.approx(math.sqrt(sx**2 + sy**2 + sz**2), rel=1e-14)
            assert gam_sk**2 - (ptot_sk / (m * c)) ** 2 == pytest.approx(1.0, abs=1e-9 * gam_sk**2)
            # Lorentz boost along x: p_x' = gam0 (p_x + beta E/c), E' = gam0 (E + beta p_x c)
            assert sx == pytest.approx(gam0 * (pb + beta * gam * m * c), rel=1e-12)
            assert gam_sk == pytest.approx(gam0 * (gam + beta * pb / (m * c)), rel=1e-10)
            assert sy**2 + sz**2 == pytest.approx(pperp**2, rel=1e-12)
            io = (D * 5)(ptot, pb, pperp, gam, phi)
            z = (D * 6)(ux, 0.0, ux, gam0, 1.0, 0.0)
            olib.mcso_transform_p_PSP(h, 1.0, C.byref(io), C.byref(z), C.byref(z))
            assert io[0] == pytest.approx(ptot, rel=1e-9) and io[1] == pytest.approx(pb, abs=1e-9 * ptot)
            assert io[3] == pytest.approx(gam, rel=1e-9)
            assert math.cos(io[4]) == pytest.approx(math.cos(phi), abs=1e-7)


def test_psp_composition_matches_single_boost(eng, olib):
    """plasma(old) -> shock -> plasma(new) equals one boost by the relative velocity (parallel shock)."""
    h, m, c = eng._h, problem.MP, problem.CL
    b1, b2 = 0.6, 0.2
    g1, g2 = 1 / math.sqrt(1 - b1 * b1), 1 / math.sqrt(1 - b2 * b2)
    ptot, mu, phi = 3.0 * m * c, 0.3, 1.1
    pb, pperp, gam = ptot * mu, ptot * math.sqrt(1 - mu * mu), math.hypot(1, 3.0)
    io = (D * 5)(ptot, pb, pperp, gam, phi)
    olib.mcso_transform_p_PSP(h, 1.0, C.byref(io), C.byref((D * 6)(b1 * c, 0, b1 * c, g1, 1, 0)),
                              C.byref((D * 6)(b2 * c, 0, b2 * c, g2, 1, 0)))
    br = (b1 - b2) / (1 - b1 * b2)
    gr = 1 / math.sqrt(1 - br * br)
    assert io[1] == pytest.approx(gr * (pb + br * gam * m * c), rel=1e-11)
    assert io[2] == pytest.approx(pperp, rel=1e-11)
    assert io[3] == pytest.approx(gr * (gam + br * pb / (m * c)), rel=1e-11)


def test_scattering_max_kick_and_isotropisation(eng, olib):
    """one kick changes the pitch by at most acos(cos_max), cos_max = cos sqrt(12 pi/(xn_per eta))
    (scattering.jl:46-60); many kicks isotropise: <mu> -> 0, <mu^2> -> 1/3."""
    h, m, c = eng._h, problem.MP, problem.CL
    ptot = 0.01 * m * c
    gam = math.hypot(1, 0.01)
    gd = 1 / (problem.QCGS * 1e-5)
    for xn in (2000.0, 100.0):
        dmax = math.sqrt(12 * math.pi / (xn * 1.0))
        worst = 0.0
        for s in range(300):
            mu0 = -0.9 + 1.8 * s / 299
            io = (D * 4)(0.0, ptot * mu0, ptot * math.sqrt(1 - mu0 * mu0), 0.5)
            olib.mcso_scattering(h, s, 1, 1.0, gd, ptot, gam, xn, C.byref(io))
            assert io[1] ** 2 + io[2] ** 2 == pytest.approx(ptot**2, rel=1e-12)
            worst = max(worst, abs(math.acos(io[1] / ptot) - math.acos(mu0)))
            assert io[0] == pytest.approx(2 * math.pi * gam * m * c * gd, rel=1e-14)   # gyro period
        assert 0.5 * dmax < worst <= dmax * (1 + 1e-9)
    mus = []
    for s in range(400):
        io = (D * 4)(0.0, ptot, 0.0, 0.0)
        olib.mcso_scattering(h, 1000 + s, 600, 1.0, gd, ptot, gam, 100.0, C.byref(io))
        mus.append(io[1] / ptot)
    mus = np.array(mus)
    assert abs(mus.mean()) < 0.1 and abs((mus**2).mean() - 1 / 3) < 0.05


def test_radiation_loss_limits(eng, olib):
    """SURVEY 8c-6, particle_loop.jl:583-590."""
    h = eng._h
    f = problem.RAD_LOSS_FAC
    p, B2 = 1e-14, 1e-6
    dt_small = 1e-4 / (f * B2 * p)
    assert olib.mcso_radiation_loss(h, B2, p, dt_small) == pytest.approx(p * (1 - 1e-4), rel=1e-14)
    dt_big = 3.0 / (f * B2 * p)
    assert olib.mcso_radiation_loss(h, B2, p, dt_big) == pytest.approx(p / 4.0, rel=1e-13)
    assert olib.mcso_radiation_loss(h, B2, p, 0.0) == p
    assert f == pytest.approx(4 / 3 * problem.SIGMA_T / (problem.CL**2 * problem.ME**2 * 8 * math.pi), rel=1e-14)


def test_psd_bin_edges(eng, olib):
    """SURVEY 8c-7, get_psd_bins.jl:16-97."""
    run, h = eng.run, eng._h
    pmin, M, T = run.psd_mom_min, run.num_psd_mom_bins, run.num_psd_theta_bins
    bm, ba = olib.mcso_psd_bin_momentum, olib.mcso_psd_bin_angle
    assert bm(h, pmin * 0.999) == 0 and bm(h, pmin) == 1 and bm(h, pmin * 10**0.0999) == 1
    assert bm(h, pmin * 10**0.1001) == 2 and bm(h, pmin * 10.0**5.05) == 51
    assert bm(h, pmin * 1e300) == M                                   # clamp (+ warning counter)
    assert ba(h, 0.0, 0.0) == 0                                        # zero momentum
    assert ba(h, 1.0, 1.0) == T                                        # p_cos = -1: downstream-pointing, last linear bin
    assert ba(h, -1.0, 1.0) == 0                                       # theta = 0 < theta_min
    th = run.psd_theta_min * 10**0.25
    assert ba(h, -math.cos(th), 1.0) == 3
    assert ba(h, -(run.psd_cos_fine - 1e-9), 1.0) == T - int((run.psd_cos_fine - 1e-9 + 1) / run.delta_cos)
    assert T - 119 <= ba(h, -(run.psd_cos_fine - 1e-9), 1.0) <= T - 118
    assert ba(h, -run.psd_cos_fine, 1.0) <= T - 119 + 1                # log-theta side of the seam


def test_mod2pi(olib):
    f = olib.mcso_mod2pi
    tp = 2 * math.pi
    assert f(1.0) == 1.0 and f(0.0) == 0.0
    assert f(-1.0) == pytest.approx(tp - 1.0, rel=1e-15) and 0 <= f(-1e-20) < tp
    assert f(tp + 0.5) == pytest.approx(0.5, abs=1e-15) and f(5 * tp + 3.0) == pytest.approx(3.0, abs=3e-15)


def test_return_probability(olib):
    """P_ret = ((v-u2)/(v+u2))^2 (prob_return.jl:89-90) for a mono-energetic beam hitting the PRP."""
    run = problem.setup_run(problem.planar_test_particle_input(200))
    e = make_engine(olib, run, n_pts_cap=40_000)
    n = 30_000
    m, c = problem.MP, problem.CL
    v = 4.0 * run.u2
    ptot = m * v                                  # non-relativistic (gamma - 1 ~ 1e-3)
    prp = 2.0 * run.x_grid_stop
    pop = dict(weight=np.full(n, 1.0 / n), ptot_pf=np.full(n, ptot), pb_pf=np.full(n, ptot * 0.999),
               x_cm=np.full(n, prp * (1 - 1e-12)), grid=np.full(n, run.n_grid, np.int64), phi_rad=np.zeros(n),
               downstream=np.ones(n, np.uint8), inj=np.ones(n, np.uint8), prp_x_cm=np.full(n, prp))
    start_ion(e, run, pop=pop)
    e.run_pcut(1, 1e30, 0.0)
    f = e.get_fates(n)
    crossed_first = (f["helix_count"] == 1) | (f["retro_steps"] > 0)
    returned = f["retro_steps"] > 0
    gam = math.hypot(1, ptot / (m * c))
    vt = ptot / (gam * m)
    p_ret = ((vt - run.u2) / (vt + run.u2)) ** 2
    # every particle crosses the PRP on its first move; those that fail the test leave with helix_count == 1
    assert crossed_first.mean() > 0.999
    first_escape = (f["helix_count"] == 1) & (f["fate"] == abi.FATE_DOWNSTREAM)
    frac = 1.0 - first_escape.mean()
    assert frac == pytest.approx(p_ret, abs=4 * math.sqrt(p_ret * (1 - p_ret) / n))
    assert returned.mean() == pytest.approx(frac, abs=1e-12)


def test_flux_conservation_without_dsa(olib):
    """with acceleration switched off every particle crosses every boundary once, so the tallied
    momentum and energy fluxes equal the far-upstream fluxes (smoothers.jl:173-177, initializers.jl:513-549)."""
    run = problem.setup_run(problem.planar_test_particle_input(3000, no_dsa=True, momentum_cutoffs=[1e9]))
    e = make_engine(olib, run)
    r = driver.main_loops(run, e, n_iters=1)[0][0]
    px = r["pxx_flux"] / run.F_px_upstream
    en = r["energy_flux"] / run.F_energy_upstream
    up = slice(43, 64)                   # between the fast-push stop and the shock
    far = slice(80, 99)                  # isotropised downstream flow
    assert np.all(np.abs(px[up] - 1) < 0.01) and np.all(np.abs(en[up] - 1) < 0.02)
    assert np.all(np.abs(px[far] - 1) < 0.03) and np.all(np.abs(en[far] - 1) < 0.03)
    assert r["tallies"].stats["n_fate"][abi.FATE_DOWNSTREAM] == r["n_pts_inj"]


def test_weight_bookkeeping(olib):
    """new_pcut conserves weight (cuts.jl:77) and every injected weight ends in exactly one fate."""
    from helpers import LADDER
    run = problem.setup_run(problem.planar_test_particle_input(800, momentum_cutoffs=LADDER[:5]))
    e = make_engine(olib, run)
    pop = start_ion(e, run)
    w_in = pop["weight"].sum()
    w_gone = 0.0
    for k, pcut in enumerate(run.pcuts, start=1):
        n = e.population_size()
        cur = e.get_population(0, n)
        ns, _ = e.run_pcut(k, pcut, run.pcuts[k - 2] if k > 1 else 0.0)
        f = e.get_fates(n)
        sv = e.get_population(1, n)
        assert ns == int(sv["l_save"].sum()) == int((f["fate"] == abi.FATE_SAVED).sum())
        assert np.all(sv["weight"][sv["l_save"] == 1] == cur["weight"][sv["l_save"] == 1])
        assert np.all(sv["ptot_pf"][sv["l_save"] == 1] > pcut) and np.all(sv["downstream"][sv["l_save"] == 1] == 1)
        w_gone += cur["weight"][f["fate"] != abi.FATE_SAVED].sum()
        w_saved = sv["weight"][sv["l_save"] == 1].sum()
        if ns == 0:
            break
        n_new, _, i_mult = e.split(inp_target := run.inp.n_pts_pcut)
        assert i_mult == max(inp_target // ns, 1) and n_new == ns * i_mult
        new = e.get_population(0, n_new)
        assert new["weight"].sum() == pytest.approx(w_saved, rel=1e-13)
        # order-preserving clones: child k of saved j sits at i_mult*rank(j)+k
        j = np.nonzero(sv["l_save"])[0]
        assert np.array_equal(new["ptot_pf"], np.repeat(sv["ptot_pf"][j], i_mult))
        assert np.array_equal(new["grid"], np.repeat(sv["grid"][j], i_mult))
    assert w_gone + w_saved == pytest.approx(w_in, rel=1e-12)
    t = e.end_ion()
    assert sum(t.stats["n_fate"]) == t.stats["n_fate"][0] + t.stats["n_fate"][1] + t.stats["n_fate"][2]
    assert t.stats["n_errors"] == 0 and t.stats["n_neg_sqrt"] == 0


def test_helix_cap_counts_as_downstream_escape(olib):
    """SURVEY B-1 (K): helix_count > 10_000 ends the particle with i_reason = 1 (particle_loop.jl:162-165);
    the bundled no-scatter run is entirely in this branch."""
    run = problem.setup_run(problem.bundled_input())
    e = make_engine(olib, run)
    pop = start_ion(e, run)
    n = len(pop["weight"])
    ns, steps = e.run_pcut(1, run.pcuts[0], 0.0)
    f = e.get_fates(n)
    assert ns == 0 and np.all(f["helix_count"] == 10_001) and np.all(f["fate"] == abi.FATE_DOWNSTREAM)
    assert steps == n * 10_001 and np.all(f["n_draws"] == 0)
