"""init_pop in run-length form, generated inside the library (mcs_begin_ion_generate).

CPU: the oracle's generator against the host mirror (problem.expand_injection with the same Philox stream) bit for bit,
over the three modes (upstream start, non-relativistic and relativistic fast push), with and without the strided
permutation, whole and in shards; the numpy Philox against the oracle's.  GPU: the CUDA generator against the oracle's
bit for bit, and a transport run started from a generated population against one started from the same population
passed through mcs_begin_ion."""
import ctypes as C

import numpy as np
import pytest

from helpers import LADDER, make_engine
from mcs_b200 import abi, driver, problem

MODES = {
    "upstream": lambda n: problem.planar_test_particle_input(n, momentum_cutoffs=LADDER[:3], fast_upstream_transport=False),
    # beta0 = 0.01 < beta_rel_fl: the non-relativistic branch of the fast push (the transport itself is not run here)
    "fastpush_nonrel": lambda n: problem.planar_test_particle_input(n, momentum_cutoffs=LADDER[:3], shock_speed=3e3),
    "fastpush_rel": lambda n: problem.planar_test_particle_input(n, momentum_cutoffs=LADDER[:3]),
    "fastpush_gamma5": lambda n: problem.ShockInput(no_scatter=False, no_dsa=False, n_pts_inj=n, n_pts_pcut=n, n_pts_pcut_hi=n,
                                                    momentum_cutoffs=problem.DEFAULT_PCUTS[:4]),
}
FIELDS = ("weight", "ptot_pf", "pb_pf", "x_cm", "phi_rad", "grid")


def _generate(lib, run, spec, seed, i_iter, i_ion, shuffle, lo=0, n_local=None):
    e = make_engine(lib, run, seed=seed)
    e.set_profile(run.profile, problem.populate_eps_target(run, run.profile), np.zeros(run.n_grid))
    e.begin_ion_generate(i_iter, i_ion, driver.species_struct(run, i_ion), spec, first_global=lo, n_local=n_local,
                         shuffle=shuffle)
    return e, e.get_population(0)


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64) if a.dtype == np.float64 else a


def test_numpy_philox_equals_oracle(olib):
    olib.mcso_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_double)]
    olib.mcso_philox.restype = None
    seed, i_iter, i_ion = 0x1234_5678_9ABC_DEF0, 7, 3
    rng = problem.PhiloxInjectionRng(seed, i_iter, i_ion)
    u1, u2 = rng.random(50), rng.random(50)
    out = np.zeros(2)
    for j in (0, 1, 17, 49):
        olib.mcso_philox(seed, j, i_ion << 16, i_iter, 2, out.ctypes.data_as(C.POINTER(C.c_double)))
        assert out[0] == u1[j] and out[1] == u2[j]


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("shuffle", [False, True])
def test_oracle_generator_equals_host_mirror(olib, mode, shuffle):
    run = problem.setup_run(MODES[mode](1003))
    spec = problem.injection_spec(run, run.profile, 1)
    assert spec.mode == {"upstream": 0, "fastpush_nonrel": 1, "fastpush_rel": 2, "fastpush_gamma5": 2}[mode]
    seed, i_iter, i_ion = 98765, 2, 1
    ref = problem.expand_injection(spec, problem.PhiloxInjectionRng(seed, i_iter, i_ion), shuffle=shuffle).pop
    n = spec.n
    assert len(ref["weight"]) == n
    _, pop = _generate(olib, run, spec, seed, i_iter, i_ion, shuffle)
    for f in FIELDS:
        assert np.array_equal(_bits(pop[f]), _bits(ref[f])), f
    assert np.all(pop["xn_per"] == run.inp.fine_scattering_Ng) and np.all(pop["prp_x_cm"] == run.x_grid_stop)
    assert not pop["downstream"].any() and not pop["inj"].any() and np.all(pop["tcut"] == 1) and not pop["acctime_sec"].any()
    # shards: any contiguous slice of slots is the same slice of the whole
    for lo, hi in ((0, n // 3), (n // 3, n - 5), (n - 5, n)):
        _, part = _generate(olib, run, spec, seed, i_iter, i_ion, shuffle, lo=lo, n_local=hi - lo)
        for f in FIELDS:
            assert np.array_equal(_bits(part[f]), _bits(ref[f][lo:hi])), (f, lo, hi)


def test_permutation_is_a_fair_sample_per_shard():
    for n in (64, 1000, 4097):
        perm = problem.injection_permutation(n)
        assert np.array_equal(np.sort(perm), np.arange(n))
        for w in (2, 8):
            for r in range(w):
                part = perm[n * r // w: n * (r + 1) // w]
                assert abs(part.mean() - (n - 1) / 2) < 0.02 * n + 64  # every shard spans the whole momentum range


def test_generate_rejects_bad_arguments(olib):
    run = problem.setup_run(MODES["fastpush_nonrel"](600))
    spec = problem.injection_spec(run, run.profile, 1)
    e = make_engine(olib, run)
    with pytest.raises(abi.McsError):
        e.begin_ion_generate(1, 1, driver.species_struct(run, 1), spec)  # no profile yet
    e.set_profile(run.profile, problem.populate_eps_target(run, run.profile), np.zeros(run.n_grid))
    with pytest.raises(abi.McsError):
        e.begin_ion_generate(1, 1, driver.species_struct(run, 1), spec, first_global=10, n_local=spec.n)  # beyond the end


def test_main_loops_with_generated_population(olib):
    """driver.main_loops(generate_in_library=True) == the same nest fed with the host mirror's population."""
    inp = problem.planar_test_particle_input(700, momentum_cutoffs=LADDER[:3])
    run = problem.setup_run(inp)
    a = driver.main_loops(run, make_engine(olib, run, seed=77), n_iters=1, generate_in_library=True, shuffle_population=True)[0][0]
    spec = problem.injection_spec(run, run.profile, 1)
    pop = problem.expand_injection(spec, problem.PhiloxInjectionRng(77, 1, 1), shuffle=True).pop
    e = make_engine(olib, run, seed=77)
    e.set_profile(run.profile, problem.populate_eps_target(run, run.profile), np.zeros(run.n_grid))
    e.begin_ion(1, 1, driver.species_struct(run, 1), pop)
    _, n_used, n_saved = e.run_ion(run.pcuts, problem.pcut_hi(inp.en_pcut_hi, run.species[0].mass), inp.n_pts_pcut, inp.n_pts_pcut_hi)
    t = e.end_ion()
    assert np.array_equal(a["n_used"], n_used) and np.array_equal(a["n_saved"], n_saved) and a["tallies"].stats == t.stats
    assert np.array_equal(a["tallies"].pxx_flux, t.pxx_flux) and np.array_equal(a["tallies"].psd, t.psd)
    assert a["n_pts_inj"] == spec.n and a["weight_running"] == pop["weight"][0]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(MODES))
def test_cuda_generator_equals_oracle(olib, clib, mode):
    run = problem.setup_run(MODES[mode](20_011))
    spec = problem.injection_spec(run, run.profile, 1)
    for shuffle, lo, n_local in ((False, 0, None), (True, 0, None), (True, 777, 9000)):
        _, a = _generate(olib, run, spec, 4242, 3, 1, shuffle, lo, n_local)
        _, b = _generate(clib, run, spec, 4242, 3, 1, shuffle, lo, n_local)
        for f in abi.POP_F64 + abi.POP_I64 + abi.POP_U8:
            assert np.array_equal(_bits(a[f]), _bits(b[f])), (f, shuffle, lo)


@pytest.mark.gpu
def test_transport_from_generated_population_equals_uploaded(clib):
    """Same particles, same counters: a run whose population was generated on the device equals, bit for bit in every
    integer and to rounding in the sums, the run that received that population through mcs_begin_ion."""
    inp = problem.planar_test_particle_input(20_000, momentum_cutoffs=LADDER[:5])
    run = problem.setup_run(inp)
    spec = problem.injection_spec(run, run.profile, 1)
    res = []
    for generated in (True, False):
        if generated:
            e, _ = _generate(clib, run, spec, 210, 1, 1, True)
        else:
            pop = problem.expand_injection(spec, problem.PhiloxInjectionRng(210, 1, 1), shuffle=True).pop
            e = make_engine(clib, run, seed=210)
            e.set_profile(run.profile, problem.populate_eps_target(run, run.profile), np.zeros(run.n_grid))
            e.begin_ion(1, 1, driver.species_struct(run, 1), pop)
        n_run, n_used, n_saved = e.run_ion(run.pcuts, problem.pcut_hi(inp.en_pcut_hi, run.species[0].mass), inp.n_pts_pcut,
                                           inp.n_pts_pcut_hi)
        res.append((n_used, n_saved, e.end_ion()))
    (u1, s1, t1), (u2, s2, t2) = res
    assert np.array_equal(u1, u2) and np.array_equal(s1, s2) and t1.stats == t2.stats
    assert np.array_equal(t1.num_crossings, t2.num_crossings)
    assert np.array_equal(t1.pxx_flux, t2.pxx_flux) and np.array_equal(t1.energy_flux, t2.energy_flux)
