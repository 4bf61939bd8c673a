"""Host driver: the iteration / ion / pcut nest of main_loops.jl around the C-ABI calls.

Mirrors /root/reference/src/main_loops.jl:52-341 with the particle loop (:228-292) and the
between-pcut population management (:297-313) replaced by mcs_* calls.  What stays on the host is
what the reference keeps on the host: eps_target (iter_init.jl), init_pop (initializers.jl),
the energy pools hand-over (main_loops.jl:164) and the 1e-99 floors (main_loops.jl:59-72).
Profile smoothing (smoothers.jl) is out of scope: the profile is constant unless the caller
passes `profile_update`.
"""
from __future__ import annotations

import math

import numpy as np

from . import abi, problem


def make_config(lib, run: problem.Run, *, n_pts_cap: int | None = None, seed: int = 210, na_cr: int | None = None,
                rng_mode: int = abi.RNG_PHILOX, compat: int = abi.COMPAT_DEFAULT, threads: int = 1,
                device: int = -1, helix_cap: int = 10_000, bin_thermal: bool = False) -> abi.McsConfig:
    """Scalars of particle_loop's argument list (particle_loop.jl:1-31) as one POD struct."""
    inp = run.inp
    c = abi.default_config(lib)
    c.device = device
    c.mp_g, c.c_cms, c.qcgs_esu = problem.MP, problem.CL, problem.QCGS
    c.E_rel_pt, c.rad_loss_fac = problem.E_REL_PT, problem.RAD_LOSS_FAC
    c.gam0, c.beta0, c.u0, c.u2, c.bmag2 = run.gam0, run.beta0, run.u0, run.u2, run.bmag2
    c.pe_crit, c.gam_e_crit, c.eta_mfp = run.pe_crit, run.gam_e_crit, inp.gyrofactor
    c.psd_mom_min, c.psd_cos_fine, c.delta_cos, c.psd_theta_min = (run.psd_mom_min, run.psd_cos_fine,
                                                                   run.delta_cos, run.psd_theta_min)
    c.psd_bins_per_dec_mom, c.psd_bins_per_dec_theta = inp.num_psd_bins_per_decade
    c.num_psd_mom_bins, c.num_psd_theta_bins = run.num_psd_mom_bins, run.num_psd_theta_bins
    c.energy_transfer_frac = inp.energy_transfer_frac
    c.feb_upstream, c.feb_downstream, c.x_grid_stop = run.feb_upstream, run.feb_downstream, run.x_grid_stop
    c.B_CMBz = run.B_CMBz
    c.xn_per_fine, c.xn_per_coarse = inp.fine_scattering_Ng, inp.coarse_scattering_Ng
    c.age_max = run.age_max
    c.n_grid, c.i_grid_feb, c.i_shock, c.n_ions = run.n_grid, run.i_grid_feb, run.i_shock, run.n_ions
    cap = max(inp.n_pts_inj, inp.n_pts_pcut, inp.n_pts_pcut_hi) + 8
    c.n_pts_max = int(n_pts_cap if n_pts_cap is not None else cap)
    c.na_cr = int(na_cr if na_cr is not None else 10 * problem.NA_PARTICLES)
    c.n_xspec = len(run.x_spec_cm)
    for i, x in enumerate(run.x_spec_cm):
        c.x_spec[i] = x
    c.n_tcuts = len(run.tcuts)
    for i, t in enumerate(run.tcuts):
        c.tcuts[i] = t
    for i, f in enumerate(run.inj_fracs):
        c.inj_fracs[i] = f
    c.do_rad_losses, c.do_retro, c.do_tcuts = int(inp.radiation_losses), int(run.do_retro), int(run.do_tcuts)
    c.dont_DSA, c.dont_scatter = int(inp.no_dsa), int(inp.no_scatter)
    c.use_custom_frg, c.use_custom_epsB = int(inp.use_custom_frg), int(inp.use_custom_epsB)
    c.helix_cap, c.seed, c.compat, c.rng_mode, c.threads = helix_cap, seed, compat, rng_mode, threads
    c.bin_thermal = int(bin_thermal)
    return c


def species_struct(run: problem.Run, i_ion: int) -> abi.McsSpecies:
    """main_loops.jl:97-102 (aa, zz, pmax_cutoff) + MonteCarloScattering.jl:493 (electron_weight_fac)."""
    sp = run.species[i_ion - 1]
    zz = abs(sp.charge) if run.inp.abs_charge else sp.charge
    ewf = run.electron_weight_fac if math.isfinite(run.electron_weight_fac) else 0.0
    return abi.McsSpecies(aa=sp.aa, zz_esu=zz, n0=sp.n0, pmax_cutoff=problem.get_pmax_cutoff(run, sp.aa),
                          electron_weight_fac=ewf)


class TorchComm:
    """Rank plumbing over an initialised torch.distributed group (gloo on CPU, nccl on GPU)."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else "cpu"

    def allgather_i64(self, v: int):
        t = self.torch.tensor([int(v)], dtype=self.torch.int64, device=self.device)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [int(o.item()) for o in out]

    def allreduce_sum_(self, a: np.ndarray):
        t = self.torch.from_numpy(a).to(self.device)
        self.dist.all_reduce(t)
        a[...] = t.cpu().numpy()

    def broadcast_bytes(self, b: bytes | None, n: int, src: int = 0) -> bytes:
        t = self.torch.zeros(n, dtype=self.torch.uint8, device=self.device)
        if self.rank == src:
            t = self.torch.tensor(list(b), dtype=self.torch.uint8, device=self.device)
        self.dist.broadcast(t, src)
        return bytes(t.cpu().tolist())


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block of the pcut's particle index range per rank."""
    return (n * rank) // world, (n * (rank + 1)) // world


def run_ion_host_comm(engine: abi.Engine, pcuts, p_pcut_hi, n_pts_pcut, n_pts_pcut_hi, comm=None):
    """loop_pcut (main_loops.jl:179-317) with the rank exchange done by the caller: one allgather of
    n_saved per pcut; i_mult from the global count (cuts.jl:42); children keep global indices."""
    n_used, n_saved_g = [], []
    for k, pcut in enumerate(pcuts, start=1):
        n_loc = engine.population_size()
        ns, _ = engine.run_pcut(k, float(pcut), float(pcuts[k - 2]) if k > 1 else 0.0)
        per_rank = comm.allgather_i64(ns) if comm is not None else [ns]
        n_used.append(sum(comm.allgather_i64(n_loc)) if comm is not None else n_loc)
        tot = sum(per_rank)
        n_saved_g.append(tot)
        if tot == 0:  # pcut_finalize: break_pcut
            break
        target = n_pts_pcut if pcut < p_pcut_hi else n_pts_pcut_hi
        i_mult = max(target // tot, 1)
        rank = comm.rank if comm is not None else 0
        engine.split_explicit(i_mult, i_mult * sum(per_rank[:rank]))
    return len(n_used), np.array(n_used), np.array(n_saved_g)


_REDUCE_F = ("pxx_flux", "pxz_flux", "energy_flux", "psd", "esc_psd_feb_upstream", "esc_psd_feb_downstream",
             "esc_energy_eff", "esc_num_eff", "weight_coupled", "spectra_coupled", "energy_transfer_pool",
             "spectra_sf", "spectra_pf", "therm_d2N_sf", "therm_d2N_pf", "dNdp_cr_sf")


def reduce_tallies_host(t: abi.Tallies, comm) -> abi.Tallies:
    """Sum per-rank tallies on the host (used with host-comm engines; the CUDA engine reduces with NCCL)."""
    for nm in _REDUCE_F:
        a = getattr(t, nm)
        if a is not None:
            comm.allreduce_sum_(a)
    comm.allreduce_sum_(t.num_crossings)
    keys = sorted(t.scalars)
    v = np.array([t.scalars[k] for k in keys])
    comm.allreduce_sum_(v)
    t.scalars = dict(zip(keys, v.tolist()))
    skeys = [k for k in sorted(t.stats) if k != "n_fate"]
    s = np.array([t.stats[k] for k in skeys] + list(t.stats["n_fate"]), dtype=np.int64)
    comm.allreduce_sum_(s)
    t.stats = dict(zip(skeys, s[: len(skeys)].tolist()))
    t.stats["n_fate"] = s[len(skeys):].tolist()
    return t


def main_loops(run: problem.Run, engine: abi.Engine, *, n_iters: int | None = None, comm=None,
               device_comm: bool = False, want_psd: bool = True, want_log: bool = True, profile_update=None,
               host_pcut_loop: bool = False, shuffle_population: bool = False, generate_in_library: bool = False,
               only_ions=None, pop_seed_offset: int = 0, thermo: bool = False):
    """loop_itr / loop_ion / loop_pcut of main_loops.jl:52-341.

    Returns a list (per iteration) of lists (per ion) of dicts with the per-ion tallies (pure sums),
    the flux arrays as the reference holds them (fast-push prefill + sums + 1e-99 floor) and the pcut
    bookkeeping.  `comm` shards the population over ranks; `device_comm` means the engine already
    reduces inside the library (NCCL) so the host must not reduce again.  `generate_in_library` hands init_pop to the
    library in run-length form (mcs_begin_ion_generate): nothing per-particle is drawn or copied by the host.
    `thermo` (engine built with bin_thermal) adds what ion_finalize.jl:38-47 gets from thermo_calcs — P_psd_par, P_psd_perp,
    energy_density_psd, d2N_pop per zone — computed by mcs_thermo on the tallies where they lie.
    """
    inp = run.inp
    prof = run.profile
    n_iters = inp.num_iterations if n_iters is None else n_iters
    rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
    results = []
    for i_iter in range(1, n_iters + 1):
        eps_target = problem.populate_eps_target(run, prof)  # main_loops.jl:80-81
        pool = np.zeros(run.n_grid)                            # zero!(energy_transfer_pool)  :83
        per_ion = []
        for i_ion in range(1, run.n_ions + 1):
            sp = run.species[i_ion - 1]
            if (sp.n0 == 0 and inp.skip_zero_density_species) or (only_ions is not None and i_ion not in only_ions):  # SURVEY B-11
                per_ion.append(None)
                continue
            engine.set_profile(prof, eps_target, pool.copy())  # energy_recv_pool .= energy_transfer_pool  :164
            if generate_in_library:
                ip = problem.injection_spec(run, prof, i_ion)
                n = ip.n
                lo, hi = shard_bounds(n, rank, world)
                engine.begin_ion_generate(i_iter, i_ion, species_struct(run, i_ion), ip, first_global=lo, n_local=hi - lo,
                                          shuffle=shuffle_population)
            else:
                rng = np.random.default_rng((i_iter - 1) * run.n_ions + (i_ion - 1) + pop_seed_offset)  # stands in for :120-121
                ip = problem.init_pop(run, prof, i_ion, rng, shuffle=shuffle_population)
                n = len(ip.pop["weight"])
                lo, hi = shard_bounds(n, rank, world)
                pop = {k: v[lo:hi] for k, v in ip.pop.items()}
                engine.begin_ion(i_iter, i_ion, species_struct(run, i_ion), pop, first_global=lo)
            p_hi = problem.pcut_hi(inp.en_pcut_hi, sp.mass)
            if host_pcut_loop or (comm is not None and not device_comm):
                n_run, n_used, n_saved = run_ion_host_comm(engine, run.pcuts, p_hi, inp.n_pts_pcut,
                                                           inp.n_pts_pcut_hi, comm)
            else:
                n_run, n_used, n_saved = engine.run_ion(run.pcuts, p_hi, inp.n_pts_pcut, inp.n_pts_pcut_hi)
            t = engine.end_ion(want_psd=want_psd, want_log=want_log)
            if comm is not None and not device_comm:
                reduce_tallies_host(t, comm)
            pool = pool + t.energy_transfer_pool
            th = {}
            if thermo:
                cosc, ptc, zone_pop = problem.thermo_inputs(run, prof, i_ion - 1)
                host = comm is not None and not device_comm      # the sums over ranks exist on the host only
                res = engine.thermo(cosc, ptc, zone_pop, sp.T, **(dict(psd=t.psd, therm_d2N_pf=t.therm_d2N_pf,
                                                                        num_crossings=t.num_crossings) if host else {}))
                th = dict(zip(("P_psd_par", "P_psd_perp", "energy_density_psd", "d2N_pop"), res), zone_pop=zone_pop)
            per_ion.append(dict(**th,
                tallies=t, n_pcuts_run=n_run, n_used=n_used, n_saved=n_saved, n_pts_inj=n,
                pxx_flux=ip.pxx_flux + t.pxx_flux + 1.0e-99, pxz_flux=ip.pxz_flux + t.pxz_flux + 1.0e-99,
                energy_flux=ip.energy_flux + t.energy_flux + 1.0e-99, weight_running=ip.weight_running,
            ))
        results.append(per_ion)
        if profile_update is not None:
            prof = profile_update(run, prof, per_ion)
    return results
