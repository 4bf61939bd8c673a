"""Edge cases of the boundary: empty and ragged populations, the size cap, calls out of order, a pcut nobody survives.
The same checks run against the CPU oracle (here) and against the CUDA library (-m gpu), where the two are also
compared with each other."""
import numpy as np
import pytest

from helpers import LADDER, make_engine, start_ion
from mcs_b200 import abi, driver, problem


def _run(n=600, **kw):
    return problem.setup_run(problem.planar_test_particle_input(n, momentum_cutoffs=LADDER[:3], **kw))


def _slice(pop, lo, hi):
    return {k: np.ascontiguousarray(v[lo:hi]) for k, v in pop.items()}


def _one_pcut(lib, run, pop, first_global=0):
    e = make_engine(lib, run)
    start_ion(e, run, pop=pop, first_global=first_global)
    n = e.population_size()
    ns, steps = e.run_pcut(1, run.pcuts[0], 0.0)
    return e, n, ns, steps


def _empty_population(lib):
    run = _run()
    pop = _slice(problem.init_pop(run, run.profile, 1, np.random.default_rng(0)).pop, 0, 0)
    e, n, ns, steps = _one_pcut(lib, run, pop)
    assert (n, ns, steps) == (0, 0, 0)
    n_run, n_used, n_saved = e.run_ion(run.pcuts, 1e30, 600, 600)
    assert n_run == 1 and n_saved[0] == 0            # pcut_finalize: break on an empty pcut (cuts.jl:105-110)
    t = e.end_ion()
    assert not t.pxx_flux.any() and not t.psd.any() and t.stats["n_helix_steps"] == 0 and sum(t.stats["n_fate"]) == 0


def _ragged_sizes(lib, other=None):
    """1, 31, 33 and 257 particles: partial warps and partial blocks give the per-particle results of the full run."""
    run = _run()
    full = problem.init_pop(run, run.profile, 1, np.random.default_rng(0)).pop
    e, n, ns, _ = _one_pcut(lib, run, full)
    ref = e.get_fates(n)
    for m in (1, 31, 33, 257):
        e2, n2, _, _ = _one_pcut(lib, run, _slice(full, 0, m))
        f = e2.get_fates(n2)
        assert n2 == m
        for key in ("fate", "helix_count", "n_draws"):
            assert np.array_equal(f[key], ref[key][:m]), (m, key)
        if other is not None:
            e3, _, _, _ = _one_pcut(other, run, _slice(full, 0, m))
            g = e3.get_fates(m)
            for key in ("fate", "helix_count", "n_draws"):
                assert np.array_equal(f[key], g[key]), (m, key)
    # a shard that starts in the middle keeps the global RNG indices
    e4, n4, _, _ = _one_pcut(lib, run, _slice(full, 100, 163), first_global=100)
    assert np.array_equal(e4.get_fates(n4)["helix_count"], ref["helix_count"][100:163])


def _size_cap_and_call_order(lib):
    run = _run(n=200)
    pop = problem.init_pop(run, run.profile, 1, np.random.default_rng(0)).pop
    n = len(pop["weight"])
    e = make_engine(lib, run, n_pts_cap=n)            # exactly at the cap: accepted
    start_ion(e, run, pop=pop)
    assert e.population_size() == n
    small = make_engine(lib, run, n_pts_cap=n - 1)    # one above the cap: refused, with a message
    with pytest.raises(abi.McsError, match="n_pts"):
        start_ion(small, run, pop=pop)
    fresh = make_engine(lib, run)
    with pytest.raises(abi.McsError):
        fresh.run_pcut(1, run.pcuts[0], 0.0)          # no ion yet
    with pytest.raises(abi.McsError):
        fresh.begin_ion(1, 1, driver.species_struct(run, 1), pop)  # no profile yet
    bad = dict(pop)
    bad["grid"] = np.full(n, run.n_grid + 5, np.int64)
    with pytest.raises(abi.McsError, match="grid"):
        start_ion(make_engine(lib, run), run, pop=bad)
    e5 = make_engine(lib, run)
    start_ion(e5, run, pop=pop)
    with pytest.raises(abi.McsError):
        e5.split(100)                                 # split before any pcut: nothing saved
    e.run_pcut(1, run.pcuts[0], 0.0)
    with pytest.raises(abi.McsError):
        e.split_explicit(0, 0)                        # i_mult < 1
    # widths of the Philox counter fields: errors at the boundary instead of a silent wrap that would reuse streams
    e6 = make_engine(lib, run)
    with pytest.raises(abi.McsError, match="2\\^32"):
        start_ion(e6, run, pop=pop, first_global=2**32 - n + 1)   # last particle would get global index 2^32
    start_ion(e6, run, pop=pop, first_global=2**32 - n)           # the last index that fits: accepted
    with pytest.raises(abi.McsError, match="16-bit"):
        e6.run_pcut(0x10000, run.pcuts[0], 0.0)
    ns, _ = e6.run_pcut(0xFFFF, run.pcuts[0], 0.0)                # the largest pcut number that fits
    assert ns >= 0


def _nobody_survives(lib):
    """A cut-off above every reachable momentum: the pcut loop stops after the first pcut with nothing saved and every
    particle accounted for by an escape fate."""
    run = _run(n=300)
    run.pcuts = [1.0e6 * run.pcuts[-1]]
    e = make_engine(lib, run)
    pop = start_ion(e, run)
    n = len(pop["weight"])
    n_run, n_used, n_saved = e.run_ion(run.pcuts, 1e30, 300, 300)
    t = e.end_ion()
    assert n_run == 1 and n_used[0] == n and n_saved[0] == 0
    assert t.stats["n_fate"][0] == 0 and sum(t.stats["n_fate"]) == n


def test_empty_population(olib):
    _empty_population(olib)


def test_ragged_sizes(olib):
    _ragged_sizes(olib)


def test_size_cap_and_call_order(olib):
    _size_cap_and_call_order(olib)


def test_pcut_nobody_survives(olib):
    _nobody_survives(olib)


@pytest.mark.gpu
def test_empty_population_cuda(clib):
    _empty_population(clib)


@pytest.mark.gpu
def test_ragged_sizes_cuda(olib, clib):
    _ragged_sizes(clib, other=olib)


@pytest.mark.gpu
def test_size_cap_and_call_order_cuda(clib):
    _size_cap_and_call_order(clib)


@pytest.mark.gpu
def test_pcut_nobody_survives_cuda(clib):
    _nobody_survives(clib)
