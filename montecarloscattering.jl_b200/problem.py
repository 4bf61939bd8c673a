"""Host-side producer of the hot path's inputs (numpy; runs once per run / per ion).

The reference builds these in Julia from ./mc_in.toml; no Julia exists in this image, so the few
routines whose outputs the transport kernel ingests are restated here so that tests, smoke() and
bench.py can make the five BASELINE.json configs.  Raw cgs doubles throughout.

  ShockInput            <- keys of mc_in.toml                      (/root/reference mc_in.toml:1-239)
  setup_run()           <- MonteCarloScattering.jl:70-338,414-493  (scalars, PSD bin parameters, grid, profile)
  setup_grid()          <- initializers.jl:403-476                 (as written, incl. the non-monotonic upstream
                                                                    block; `fixed_grid` opts out)
  setup_profile()       <- initializers.jl:774-850, 879-945
  upstream_fluxes()     <- initializers.jl:513-626
  set_inj_dist()        <- initializers.jl:1251-1453
  init_pop()            <- initializers.jl:977-1134 + ion_init.jl:29-53
  populate_eps_target() <- iter_init.jl:1-15
  get_pmax_cutoff(), pcut_hi() <- ion_init.jl:55-82
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# constants (Unitful 1.x / CODATA 2018)
MP = 1.67262192369e-24
ME = 9.1093837015e-28
CL = 2.99792458e10
QCGS = 4.80320471257e-10
KB = 1.380649e-16
KEV = 1.602176634e-9
SIGMA_T = 6.6524587321e-25
B_CMB0 = 3.27e-6
PC_CM = 3.0856775814913674e18
RAD_LOSS_FAC = 4.0 / 3.0 * CL * SIGMA_T / (CL**3 * ME**2 * 8 * math.pi)  # constants.jl:30
# parameters.jl:9-32
NA_PARTICLES = 100_000
NUM_THERM_BINS = 150
BETA_REL_FL = 0.02
E_REL_PT = 0.005

DEFAULT_PCUTS = [
    0.01, 0.6, 1.6, 2.0, 4.5, 9.0, 30.0, 50.0, 200.0, 300.0, 500.0, 1000.0, 2000.0, 5000.0, 1.000e4, 3.162e4,
    1.000e5, 3.162e5, 1.000e6, 3.162e6, 1.000e7, 1.778e7, 3.162e7, 5.623e7, 1.000e8, 1.778e8, 3.162e8, 5.623e8,
    1.000e9, 1.778e9, 3.162e9, 5.623e9, 1.00e10, 1.778e10, 3.162e10, 5.623e10, 1.000e11, 1.778e11, 3.162e11,
    5.623e11, 1.000e12, 1.778e12, 3.162e12, 5.623e12, 1.000e13,
]  # mc_in.toml:84-130

# Cut-offs for the non-relativistic configs [m_p c]. The bundled list was written for gamma0 = 5: at
# u0 = 1e4 km/s it leaves ONE wide pcut (0.01 -> 0.6 m_p c) below the FEB-limited maximum (~0.6 m_p c), i.e. no
# splitting at all; this ladder splits every factor ~1.5 so each pcut holds its target population.
NONREL_PCUTS = [0.01, 0.04, 0.06, 0.09, 0.13, 0.2, 0.3, 0.45, 0.6, 1.0, 1.6, 2.5]


@dataclass
class ShockInput:
    """The keys of mc_in.toml that reach the hot path (defaults = the bundled file)."""
    shock_speed: float = 5.0
    shock_speed_unit: str = "gamma"
    num_iterations: int = 20
    coarse_scattering_Ng: float = 100.0
    fine_scattering_Ng: float = 2000.0
    aa_ion: list = field(default_factory=lambda: [1.0, float("nan")])  # nan = electron
    zz_ion: list = field(default_factory=lambda: [1.0, -1.0])
    tz_ion: list = field(default_factory=lambda: [1e6, 1e6])
    denz_ion: list = field(default_factory=lambda: [1.0, 0.0])
    input_distribution: int = 1
    injection_energy_keV: float = 1e3
    injection_weights: bool = True
    maximum_energy: tuple = (0.0, 0.0, 1e10)
    gyrofactor: float = 1.0
    b_mag_upstream: float = 1e-5
    theta_b0: float = 0.0
    x_grid_limits: tuple = (-1e7, 1e1)
    feb_upstream: tuple = (-1e2, 0.0)
    feb_downstream: tuple = (0.0, 0.0)
    x_spec: list = field(default_factory=list)
    use_custom_frg: bool = False
    n_pts_inj: int = 100
    n_pts_pcut: int = 400
    n_pts_pcut_hi: int = 2000
    en_pcut_hi: float = 1_000_000.0
    momentum_cutoffs: list = field(default_factory=lambda: list(DEFAULT_PCUTS))
    no_scatter: bool = True
    no_dsa: bool = True
    smooth_shocks: bool = False
    target_compression_ratio: float = -1.0
    maximum_age: float = 3.15e11
    tcuts: list | None = field(default_factory=lambda: [1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 3e13])
    use_retro: bool | None = True
    fast_upstream_transport: bool = True
    proton_fast_transport_stop: float = -1.0
    electron_energy_mfp_threshold: float = 1e4
    radiation_losses: bool = True
    jet_distance: float = 1e3
    redshift: float = 0.0
    energy_transfer_frac: float = 0.1
    b_field_turbulence: float = 1.0
    b_field_amplify: float = 1.0
    use_custom_epsB: bool = True
    num_psd_bins_per_decade: tuple = (10, 10)
    psd_linear_cosine_bins: int = 119
    psd_log_theta_decs: int = 4
    emin_therm_fac: float = 0.01  # EMNFC
    inj_fracs: list | None = None  # INJFR
    # --- build-specific switches (not in mc_in.toml) --------------------------------------------
    fixed_grid: bool = False        # True: monotonic upstream log zones + downstream block ending at x_grid_stop
    compat_zero_first_particle: bool = False  # SURVEY B-7: keep the as-written ptot=0 first particle
    abs_charge: bool = True         # electrons: the reference passes zz = -qcgs into gyro_denom; use |zz|
    skip_zero_density_species: bool = True  # SURVEY B-11
    allow_large_n: bool = True      # lift parameters.jl na_particles = 100_000 (SURVEY B-16)


def bundled_input() -> ShockInput:
    """config 0: mc_in.toml verbatim."""
    return ShockInput()


def planar_test_particle_input(n_per_pcut: int = 1_000_000, **kw) -> ShockInput:
    """config 1: unmodified planar nonrelativistic shock, protons, no smoothing."""
    d = dict(
        shock_speed=1e4, shock_speed_unit="km/s", num_iterations=1, aa_ion=[1.0], zz_ion=[1.0], tz_ion=[1e6],
        denz_ion=[1.0], n_pts_inj=n_per_pcut, n_pts_pcut=n_per_pcut, n_pts_pcut_hi=n_per_pcut, no_scatter=False,
        no_dsa=False, use_retro=True, tcuts=None, maximum_age=-1.0, energy_transfer_frac=0.0,
        b_field_turbulence=0.0, use_custom_epsB=False, radiation_losses=False, smooth_shocks=False,
        momentum_cutoffs=list(NONREL_PCUTS),
    )
    d.update(kw)
    return ShockInput(**d)


def nonlinear_input(n_per_pcut: int = 10_000_000, **kw) -> ShockInput:
    """config 2: as planar, 10 iterations, forced compression ratio > r_RH (modified precursor)."""
    d = dict(num_iterations=10, smooth_shocks=True, target_compression_ratio=8.0)
    d.update(kw)
    return planar_test_particle_input(n_per_pcut, **d)


def relativistic_input(n_per_pcut: int = 1_000_000, **kw) -> ShockInput:
    """config 3: upstream Lorentz factor 10."""
    d = dict(shock_speed=10.0, shock_speed_unit="gamma", momentum_cutoffs=list(DEFAULT_PCUTS))
    d.update(kw)
    return planar_test_particle_input(n_per_pcut, **d)


def multi_species_input(n_per_pcut: int = 10_000_000, **kw) -> ShockInput:
    """config 4: p + He + e-, gamma0 = 1.5, radiative losses and energy transfer on."""
    d = dict(
        shock_speed=1.5, shock_speed_unit="gamma", aa_ion=[1.0, 4.0, float("nan")], zz_ion=[1.0, 2.0, -1.0],
        tz_ion=[1e6, 1e6, 1e6], denz_ion=[1.0, 0.1, 1.2], radiation_losses=True, energy_transfer_frac=0.1,
        fast_upstream_transport=True, momentum_cutoffs=list(DEFAULT_PCUTS),
    )
    d.update(kw)
    return planar_test_particle_input(n_per_pcut, **d)


@dataclass
class Profile:
    x_grid_rg: np.ndarray
    x_grid_cm: np.ndarray
    ux_sk: np.ndarray
    uz_sk: np.ndarray
    utot: np.ndarray
    gam_sf: np.ndarray
    gam_ef: np.ndarray
    beta_ef: np.ndarray
    btot: np.ndarray
    theta: np.ndarray
    epsB: np.ndarray


@dataclass
class Species:
    aa: float
    mass: float
    charge: float  # esu, signed as in the reference
    T: float
    n0: float
    is_electron: bool


@dataclass
class Run:
    inp: ShockInput
    species: list
    u0: float
    beta0: float
    gam0: float
    u2: float
    beta2: float
    gam2: float
    bmag0: float
    bmag2: float
    rg0: float
    r_comp: float
    r_RH: float
    x_grid_start: float
    x_grid_stop: float
    feb_upstream: float
    feb_downstream: float
    use_prp: bool
    n_grid: int
    i_grid_feb: int
    i_shock: int
    profile: Profile
    F_px_upstream: float
    F_pz_upstream: float
    F_energy_upstream: float
    pcuts: np.ndarray
    tcuts: np.ndarray
    do_tcuts: bool
    age_max: float
    do_retro: bool
    pe_crit: float
    gam_e_crit: float
    B_CMBz: float
    psd_mom_min: float
    psd_mom_max: float
    num_psd_mom_bins: int
    num_psd_theta_bins: int
    psd_cos_fine: float
    psd_theta_min: float
    delta_cos: float
    Emax: float
    Emax_per_aa: float
    pmax: float
    n_pts_max: int
    electron_weight_fac: float
    inj_fracs: list
    x_spec_cm: list

    @property
    def n_ions(self):
        return len(self.species)


# ------------------------------------------------------------------------------------------------
def parse_shock_speed(v: float, unit: str):
    """data_input.jl:2-28"""
    if v <= 0:
        raise ValueError("Shock speed must be positive")
    if unit in ("gamma", "γ"):
        if v <= 1:
            raise ValueError("shock-speed: Lorentz factor must be > 1")
        g = v
        b = math.sqrt(1 - 1 / g**2)
        u = b * CL
    elif unit == "km/s":
        u = v * 1.0e5
        b = u / CL
        g = 1 / math.sqrt(1 - b**2)
    elif unit == "c":
        b = v
        u = b * CL
        g = 1 / math.sqrt(1 - b**2)
    else:
        raise ValueError("shock-speed: unknown units")
    return u, b, g


def get_redshift(d_cm_mpc: float) -> float:
    """cosmo_calc.jl:32-50 (flat LCDM with radiation, Planck-2013 numbers at :9-14)."""
    if d_cm_mpc < 0.443:
        return 0.0
    from scipy.integrate import quad
    from scipy.optimize import brentq
    h = 0.678
    om_r = 0.4165 / (h * 100) ** 2
    om_m = 0.317 - 0.5 * om_r
    om_k = 0.0
    om_l = 1 - om_k - om_m - om_r
    d_h = 2997.92458 / h
    E = lambda z: math.sqrt(om_r * (1 + z) ** 4 + om_m * (1 + z) ** 3 + om_k * (1 + z) ** 2 + om_l)
    dc = lambda z: d_h * quad(lambda t: 1 / E(t), 0, z)[0]
    return brentq(lambda z: dc(z) - d_cm_mpc, 0.0, 50.0, xtol=1e-14)


def calc_rRH(beta0, species):
    """initializers.jl:72-118 as written: every beta0 >= beta_rel_fl takes the non-relativistic Ellison (1985)
    formula (the relativistic test is inverted and its branch is a MethodError, SURVEY B-12)."""
    P0 = sum(s.n0 * s.T for s in species) * KB
    rho0 = sum(s.n0 * s.mass for s in species)
    cs = math.sqrt(5.0 / 3.0 * P0 / rho0)
    mach = beta0 * CL / cs
    return 8 / (2 + 6 / mach**2), 5.0 / 3.0


FIRST_ZONE = [-9.0, -8.0, -7.0, -6.0, -5.0, -4.5, -4.0, -3.5, -3.0, -2.5, -2.0, -1.8, -1.6, -1.4, -1.2, -1.0, -0.9,
              -0.8, -0.7, -0.6, -0.5, -0.4, -0.3, -0.2, -0.15, -0.1, -0.07, -0.05, -0.04, -0.03, -0.02, -0.015, -0.01,
              -3.0e-3, -1.0e-3]
EXTREMELY_FINE_SPACING = [-1.0e-4, -1.0e-7, 0.0, 1.0e-7, 1.0e-4]
DOWNSTREAM_SPACING = [1.0e-3, 1.0e-2, 2.0e-2, 3.0e-2, 5.0e-2, 7.0e-2, 0.1, 0.15, 0.2, 0.25, 0.3, 0.4, 0.5, 0.6, 0.8,
                      1.0]


def setup_grid(x_start_rg, x_stop_rg, use_prp, feb_downstream, rg0, fixed=False):
    """initializers.jl:436-476. Returns (x_grid_rg[0:n_grid+1], x_grid_start, x_grid_stop)."""
    x_grid_start = x_start_rg * rg0
    x_grid_stop = x_stop_rg * rg0 if use_prp else feb_downstream
    n_up, n_dn = 27, 16
    xs = [-1.0e30]
    if not fixed:
        dlog = (math.log10(-x_start_rg) - 1) / n_up - 1  # precedence slip kept (B-3)
        xs += [-(10.0 ** (math.log10(-x_start_rg) + k * (-dlog))) for k in range(n_up)]
    else:
        dlog = (math.log10(-x_start_rg) - 1) / n_up
        xs += [-(10.0 ** (math.log10(-x_start_rg) - k * dlog)) for k in range(n_up)]
    xs += FIRST_ZONE + EXTREMELY_FINE_SPACING + DOWNSTREAM_SPACING
    x_end_man = xs[-1]
    dlog = (math.log10(x_grid_stop / rg0) - math.log10(x_end_man)) / n_dn
    if not fixed:
        xs += [10.0 ** (math.log10(x_end_man) + k * dlog) for k in range(n_dn)]  # repeats 1.0, stops short
    else:
        xs += [10.0 ** (math.log10(x_end_man) + (k + 1) * dlog) for k in range(n_dn)]
    xs.append(1.0e30)
    return np.array(xs, dtype=np.float64), x_grid_start, x_grid_stop


def upstream_fluxes(species, B0, theta_B0, u0, beta0, gam0):
    """initializers.jl:513-626 (parallel shock: B_z = 0)."""
    P0 = sum(s.n0 * s.T for s in species) * KB
    rho0 = sum(s.n0 * s.mass for s in species)
    G = 5.0 / 3.0
    e0 = rho0 * CL**2 + 1 / (G - 1) * P0
    Bx = B0 * math.cos(math.radians(theta_B0))
    Bz = B0 * math.sin(math.radians(theta_B0))
    if beta0 >= BETA_REL_FL:
        F_px = (gam0 * beta0) ** 2 * (e0 + P0) + P0 + gam0**2 * ((beta0 * B0) ** 2 + Bz**2 - Bx**2) / (8 * math.pi)
        F_pz = -gam0 * Bx * Bz / (4 * math.pi)
        F_en = CL * (gam0**2 * beta0 * (e0 + P0) + gam0**2 * beta0 * Bz**2 / (4 * math.pi)) - gam0 * u0 * rho0 * CL**2
    else:
        Xi = G / (G - 1)
        F_px = rho0 * u0**2 * (1 + beta0**2) + P0 * (1 + Xi * beta0**2) + Bz**2 / (8 * math.pi)
        F_pz = -Bx * Bz / (4 * math.pi)
        F_en = rho0 * u0**3 * (1 + 1.25 * beta0**2) / 2 + P0 * u0 * Xi * (1 + beta0**2) + u0 * Bz**2 / (4 * math.pi)
    return F_px, F_pz, F_en


def setup_profile(u0, beta0, gam0, B0, theta_B0, r_comp, bturb_comp_frac, bfield_amp, use_custom_epsB, species,
                  F_px_up, F_en_up, x_grid_cm, x_grid_rg) -> Profile:
    """initializers.jl:774-850 with set_custom_epsB! :879-945."""
    n = len(x_grid_rg)
    ux = np.empty(n); gsf = np.empty(n); bef = np.empty(n); gef = np.empty(n); bt = np.empty(n)
    comp_fac = 0.0
    for i in range(n):
        if x_grid_cm[i] < 0:
            ux[i], gsf[i], bef[i], gef[i], bt[i] = u0, gam0, 0.0, 1.0, B0
        else:
            u = u0 / r_comp
            b = u / CL
            ux[i] = u
            gsf[i] = 1 / math.sqrt(1 - b**2)
            bef[i] = (beta0 - b) / (1 - beta0 * b)
            gef[i] = 1 / math.sqrt(1 - bef[i] ** 2)
            z_comp = (gam0 * u0) / (gsf[i] * u)
            aux = math.sqrt((1 + 2 * z_comp**2) / 3)
            # NB `local comp_fac` inside the Julia loop: the outer comp_fac handed to set_custom_epsB! stays 0.0
            cf = 1 + (aux - 1) * bturb_comp_frac
            amp = 1 + (cf - 1) * bfield_amp
            bt[i] = B0 * amp
    epsB = np.full(n, 1.0e-99)
    if use_custom_epsB:
        n0 = sum(s.n0 * s.mass for s in species) / MP
        e0 = n0 * MP * CL**2
        epsB0 = B0**2 / (8 * math.pi * e0)
        n0_e = species[-1].n0
        sigma = 2 * epsB0 / gam0
        with np.errstate(divide="ignore"):
            rg2sd = beta0 / math.sqrt(sigma * n0 / n0_e) if n0_e > 0 else 0.0  # n_e = 0: sqrt(Inf) -> rg2sd = 0
        ed2 = (F_en_up + gam0 * u0 * e0) / ux[-1] - F_px_up
        epsB2 = (B0 * comp_fac) ** 2 / (8 * math.pi * ed2)
        with np.errstate(divide="ignore", invalid="ignore"):
            end_decay = (5.0e-3 / epsB2) / rg2sd if (epsB2 != 0 and rg2sd != 0) else float("inf")
        for i in range(n):
            xsd = x_grid_rg[i] * rg2sd
            if xsd < -50:
                epsB[i] = max(1.04e-5 / abs(xsd) ** 0.6, epsB0)
            elif xsd < 50:
                epsB[i] = 1.0e-4
            elif x_grid_rg[i] < end_decay:
                epsB[i] = 5.0e-3 / xsd
            else:
                epsB[i] = epsB2
        for i in range(n):
            ed = (F_en_up + gam0 * u0 * e0) / ux[i] - F_px_up
            bt[i] = math.sqrt(abs(8 * math.pi * epsB[i] * ed))
    return Profile(x_grid_rg=np.asarray(x_grid_rg, float), x_grid_cm=np.asarray(x_grid_cm, float), ux_sk=ux,
                   uz_sk=np.zeros(n), utot=ux.copy(), gam_sf=gsf, gam_ef=gef, beta_ef=bef, btot=bt,
                   theta=np.full(n, math.radians(theta_B0)), epsB=epsB)


def setup_run(inp: ShockInput) -> Run:
    """MonteCarloScattering.jl:70-493 — everything the transport loop needs that is fixed for the run."""
    u0, beta0, gam0 = parse_shock_speed(inp.shock_speed, inp.shock_speed_unit)
    species = []
    for aa, zz, T, n in zip(inp.aa_ion, inp.zz_ion, inp.tz_ion, inp.denz_ion):
        ele = isinstance(aa, float) and math.isnan(aa)
        if ele:
            aa, zz = ME / MP, -1.0
        species.append(Species(aa=aa, mass=aa * MP, charge=zz * QCGS, T=T, n0=n, is_electron=ele))
    e = inp.maximum_energy
    Emax = Emax_per_aa = pmax = 0.0
    if e[0] > 0:
        Emax = e[0] * KEV
    elif e[1] > 0:
        Emax_per_aa = e[1] * KEV
    elif e[2] > 0:
        pmax = e[2] * MP * CL
    else:
        raise ValueError("ENMAX: at least one choice must be non-zero.")
    B0 = inp.b_mag_upstream
    rg0 = (gam0 * MP * CL**2 * beta0) / (QCGS * B0)
    if inp.theta_b0 != 0:
        raise ValueError("program cannot currently handle oblique shocks")
    xs_rg, xe_rg = inp.x_grid_limits
    # get_feb, data_input.jl:136-166
    fu = inp.feb_upstream
    feb_up = fu[0] * rg0 if fu[0] < 0 else fu[1] * PC_CM
    fd = inp.feb_downstream
    use_prp = False
    if fd[0] > 0:
        feb_dn = fd[0] * rg0
    elif fd[1] > 0:
        feb_dn = fd[1] * PC_CM
    else:
        feb_dn, use_prp = 0.0, True
    if not inp.allow_large_n and max(inp.n_pts_inj, inp.n_pts_pcut, inp.n_pts_pcut_hi) > NA_PARTICLES:
        raise ValueError("Array size na_particles too small.")
    pcuts = np.array(inp.momentum_cutoffs, float) * (MP * CL)
    r_RH, _ = calc_rRH(beta0, species)
    r_comp = r_RH if inp.target_compression_ratio == -1 else inp.target_compression_ratio
    beta2 = beta0 / r_comp
    gam2 = 1 / math.sqrt(1 - beta2**2)
    u2 = beta2 * CL
    age_max = inp.maximum_age if inp.maximum_age >= 0 else -1.0
    do_retro = inp.use_retro if inp.use_retro is not None else age_max > 0
    # parse_electron_critical_energy, data_input.jl:51-69
    Ec = inp.electron_energy_mfp_threshold
    if Ec is None or Ec <= 0:
        pe_crit, ge_crit = -ME * CL, -1.0
    else:
        rm = Ec * KEV / (ME * CL**2)
        if rm < 1.0e-2:
            pe_crit, ge_crit = math.sqrt(2 * ME * Ec * KEV), 1.0
        else:
            pe_crit, ge_crit = ME * CL * math.sqrt((rm + 1) ** 2 - 1), rm + 1
    do_tcuts = inp.tcuts is not None
    tcuts = np.array(inp.tcuts if do_tcuts else [], float)
    if do_tcuts:
        if age_max < 0:
            raise ValueError("tcut tracking must be used with an accel time limit")
        if tcuts[-1] <= 10 * age_max:
            raise ValueError("TCUTS: final tcut must be much (10x) larger than age_max.")
    x_grid_rg, x_grid_start, x_grid_stop = setup_grid(xs_rg, xe_rg, use_prp, feb_dn, rg0, fixed=inp.fixed_grid)
    n_grid = len(x_grid_rg) - 2
    x_grid_cm = x_grid_rg * rg0
    # PSD parameters, MonteCarloScattering.jl:275-338
    lin = inp.psd_linear_cosine_bins
    psd_cos_fine = 1 - 2 / (lin + 1)
    th_fine = math.acos(psd_cos_fine)
    psd_theta_min = th_fine / 10.0**inp.psd_log_theta_decs
    if inp.input_distribution == 1:
        Emin = KB * min(s.T for s in species) * inp.emin_therm_fac
    else:
        Emin = inp.injection_energy_keV * KEV / 5
    m_min = min(s.mass for s in species)
    if Emin < m_min * CL**2 / 1000:
        psd_mom_min = math.sqrt(2 * m_min * Emin)
    else:
        psd_mom_min = m_min * CL * math.sqrt((1 + Emin / (m_min * CL**2)) ** 2 - 1)
    m_max = max(s.mass for s in species)
    if Emax > 0:
        psd_mom_max = m_max * CL * math.sqrt((1 + Emax / (m_max * CL**2)) ** 2 - 1)
    elif Emax_per_aa > 0:
        psd_mom_max = m_max * CL * math.sqrt((1 + Emax_per_aa / (MP * CL**2)) ** 2 - 1)
    else:
        psd_mom_max = pmax
    psd_mom_max *= 2 * gam0
    bpd_p, bpd_t = inp.num_psd_bins_per_decade
    num_psd_mom_bins = int(math.log10(psd_mom_max / psd_mom_min) * bpd_p) + 2  # initializers.jl:217-218
    log_t_bins = int(math.log10(th_fine / psd_theta_min) * bpd_t)  # :269
    delta_cos = (psd_cos_fine + 1) / lin
    num_psd_theta_bins = log_t_bins + lin
    i_grid_feb = int(np.argmax(x_grid_cm > feb_up)) - 1  # findfirst(>(feb_upstream)) - 1
    z = inp.redshift
    if inp.jet_distance > 0 and inp.redshift > 0:
        raise ValueError("At most one of jet-distance and redshift may be non-zero")
    if inp.jet_distance > 0:
        z = get_redshift(inp.jet_distance)
    B_CMBz = B_CMB0 * (1 + z) ** 2
    F_px, F_pz, F_en = upstream_fluxes(species, B0, inp.theta_b0, u0, beta0, gam0)
    prof = setup_profile(u0, beta0, gam0, B0, inp.theta_b0, r_comp, inp.b_field_turbulence, inp.b_field_amplify,
                         inp.use_custom_epsB, species, F_px, F_en, x_grid_cm, x_grid_rg)
    bmag2 = float(prof.btot[-1])
    i_shock = int(np.nonzero(x_grid_rg <= 0)[0][-1])  # findlast(<=(0))
    n_pts_max = max(inp.n_pts_pcut, inp.n_pts_pcut_hi)
    with np.errstate(divide="ignore"):
        ewf = float(np.float64(1.0) / np.float64(species[-1].n0))
    inj_fracs = list(inp.inj_fracs) if inp.inj_fracs is not None else [1.0] * len(species)
    return Run(inp=inp, species=species, u0=u0, beta0=beta0, gam0=gam0, u2=u2, beta2=beta2, gam2=gam2, bmag0=B0,
               bmag2=bmag2, rg0=rg0, r_comp=r_comp, r_RH=r_RH, x_grid_start=x_grid_start, x_grid_stop=x_grid_stop,
               feb_upstream=feb_up, feb_downstream=feb_dn, use_prp=use_prp, n_grid=n_grid, i_grid_feb=i_grid_feb,
               i_shock=i_shock, profile=prof, F_px_upstream=F_px, F_pz_upstream=F_pz, F_energy_upstream=F_en,
               pcuts=pcuts, tcuts=tcuts, do_tcuts=do_tcuts, age_max=age_max, do_retro=bool(do_retro),
               pe_crit=pe_crit, gam_e_crit=ge_crit, B_CMBz=B_CMBz, psd_mom_min=psd_mom_min, psd_mom_max=psd_mom_max,
               num_psd_mom_bins=num_psd_mom_bins, num_psd_theta_bins=num_psd_theta_bins, psd_cos_fine=psd_cos_fine,
               psd_theta_min=psd_theta_min, delta_cos=delta_cos, Emax=Emax, Emax_per_aa=Emax_per_aa, pmax=pmax,
               n_pts_max=n_pts_max, electron_weight_fac=ewf, inj_fracs=inj_fracs,
               x_spec_cm=[x * rg0 for x in inp.x_spec])


# ------------------------------------------------------------------------------------------------
def populate_eps_target(run: Run, prof: Profile) -> np.ndarray:
    """iter_init.jl:1-15 with z_max of main_loops.jl:80; eps_target has n_grid entries, Julia index i -> [i-1]."""
    eps = np.zeros(run.n_grid)
    z_max = run.gam0 * run.beta0 / (run.gam2 * run.beta2)
    prefac = run.inp.energy_transfer_frac / (z_max - 1)
    for i in range(1, run.n_grid + 1):
        if prof.ux_sk[i] != run.u0:
            z_curr = run.gam0 * run.u0 / (prof.gam_sf[i] * prof.ux_sk[i])
            eps[i - 1] = prefac * (z_curr - 1)
    return eps


def get_pmax_cutoff(run: Run, aa: float) -> float:
    """ion_init.jl:55-72"""
    m = aa * MP
    E0 = m * CL**2
    if run.Emax > 0:
        return m * CL * math.sqrt((1 + run.Emax / E0) ** 2 - 1)
    if run.Emax_per_aa > 0:
        return m * CL * math.sqrt((1 + run.Emax_per_aa / E0) ** 2 - 1)
    if run.pmax > 0:
        return run.pmax
    raise ValueError("Max CR energy not set")


def pcut_hi(energy_pcut_hi_keV: float, m: float) -> float:
    """ion_init.jl:74-82. The non-relativistic branch returns a pure number in the reference (unit slip,
    SURVEY B-8); the momentum evidently meant is m_p c sqrt(2 E/m_p c^2)."""
    e = energy_pcut_hi_keV * KEV / (MP * CL**2)
    if e < E_REL_PT:
        return MP * CL * math.sqrt(2 * e)
    return m * CL * math.sqrt((e + 1) ** 2 - 1)


def _mb_bin_area(p1, p2, E1, E2):
    f1 = math.exp(2 * math.log(p1) - E1)
    f2 = math.exp(2 * math.log(p2) - E2)
    return (p2 - p1) * (f1 + f2) / 2


def inj_bins(inj_weight: bool, n_pts_inj: int, inp_distr: int, T_or_E: float, m: float, n0: float,
             compat_zero_first=False):
    """initializers.jl:1251-1328 with its helpers :1330-1514 in run-length form: the injected particles are
    `count[b]` copies of (ptot[b], weight[b]), bins in order of momentum.  Returns (ptot[nb], weight[nb], count[nb])."""
    if not 0 < inp_distr < 3:
        raise ValueError("Code can only do inp_distr = 1 or 2.")
    nb = NUM_THERM_BINS
    E0 = m * CL**2
    kT = KB * T_or_E
    kT_min, kT_max = 2.0e-3 * kT, 10 * kT
    if kT / E0 < E_REL_PT:
        p_min, p_max = math.sqrt(2 * m * kT_min), math.sqrt(2 * m * kT_max)
    else:
        p_min = math.sqrt((kT_min + E0) ** 2 - E0**2) / CL
        p_max = math.sqrt((kT_max + E0) ** 2 - E0**2) / CL
    dp = (p_max - p_min) / nb
    p_range = [p_min + k * dp for k in range(nb + 1)]
    if kT / E0 < E_REL_PT:
        E_range = [p * p / (2 * m * kT) for p in p_range]
    else:
        E_range = [math.hypot(p * CL, E0) / kT for p in p_range]
    areas = [_mb_bin_area(p_range[i], p_range[i + 1], E_range[i], E_range[i + 1]) for i in range(nb)]
    area_tot = 0.0
    for a in areas:
        area_tot += a
    centres = np.array([math.sqrt(p_range[i] * p_range[i + 1]) for i in range(nb)])
    if inj_weight:
        area_per_pt = area_tot / n_pts_inj
        counts = np.array([int(np.round(a / area_per_pt)) for a in areas], np.int64)  # Julia round(Int, x): ties to even
        ptot = centres
        if compat_zero_first:  # `n_pts_tot = 1` at :1425 leaves slot 1 at ptot = 0 (SURVEY B-7)
            ptot = np.concatenate(([0.0], ptot))
            counts = np.concatenate(([1], counts)).astype(np.int64)
        n = int(counts.sum())
        weight = np.full(len(ptot), n0 / n)
    else:
        n_per_bin = n_pts_inj // nb
        if n_per_bin < 5:
            raise ValueError("too few particles per bin; increase n_pts_inj")
        ptot = centres
        counts = np.full(nb, n_per_bin, np.int64)
        weight = np.array([a / area_tot / n_per_bin * n0 for a in areas])
    n_tot = int(counts.sum())
    if inp_distr == 2:
        E_inj = T_or_E * KEV
        pm = math.sqrt(2 * m * E_inj) if E_inj / E0 < E_REL_PT else math.sqrt(E_inj**2 - E0**2) / CL
        ptot, weight, counts = np.array([pm]), np.array([n0 / n_tot]), np.array([n_pts_inj], np.int64)
    return np.asarray(ptot, float), np.asarray(weight, float), counts


def set_inj_dist(inj_weight: bool, n_pts_inj: int, inp_distr: int, T_or_E: float, m: float, n0: float,
                 compat_zero_first=False):
    """initializers.jl:1251-1328, one entry per particle. Returns (ptot[n], weight[n])."""
    ptot, weight, counts = inj_bins(inj_weight, n_pts_inj, inp_distr, T_or_E, m, n0, compat_zero_first)
    return np.repeat(ptot, counts), np.repeat(weight, counts)


# McsInjection.mode (include/mcs.h)
INJ_UPSTREAM, INJ_FASTPUSH_NONREL, INJ_FASTPUSH_REL = 0, 1, 2
INJ_PERM_STRIDE = 64


@dataclass
class InjectionSpec:
    """init_pop in run-length form: everything a generator (host mirror, oracle, CUDA) needs to emit the population.
    Per bin: momentum, weight, count and the two ends (lo, hi) of the range the x-velocity of a fast-pushed particle is
    drawn from (TriangularDist(lo, hi, hi), initializers.jl:1100-1125), `gfac` = gamma_pf * m.  All per-bin values are
    computed here, on the host, so a generator only does sqrt / add / mul / div per particle (bit-reproducible)."""
    mode: int
    bin_ptot: np.ndarray
    bin_weight: np.ndarray
    bin_count: np.ndarray
    bin_lo: np.ndarray
    bin_hi: np.ndarray
    bin_gfac: np.ndarray
    x_cm: float
    grid: int
    u_stop: float
    pxx_flux: np.ndarray
    pxz_flux: np.ndarray
    energy_flux: np.ndarray

    @property
    def n(self) -> int:
        return int(self.bin_count.sum())

    @property
    def weight_running(self) -> float:
        nz = np.nonzero(self.bin_count)[0]
        return float(self.bin_weight[nz[0]]) if len(nz) else 0.0


@dataclass
class InitPop:
    pop: dict
    pxx_flux: np.ndarray
    pxz_flux: np.ndarray
    energy_flux: np.ndarray
    weight_running: float


def injection_spec(run: Run, prof: Profile, i_ion: int) -> InjectionSpec:
    """The deterministic part of init_pop (initializers.jl:977-1134) and F_update! (:1157-1223)."""
    inp = run.inp
    sp = run.species[i_ion - 1]
    m, ng = sp.mass, run.n_grid
    pxx, pxz, efl = np.zeros(ng), np.zeros(ng), np.zeros(ng)
    if not inp.fast_upstream_transport:
        T_or_E = sp.T if inp.input_distribution == 1 else inp.injection_energy_keV
        ptot, w, cnt = inj_bins(inp.injection_weights, inp.n_pts_inj, inp.input_distribution, T_or_E, m, sp.n0,
                                inp.compat_zero_first_particle)
        z = np.zeros(len(ptot))
        return InjectionSpec(INJ_UPSTREAM, ptot, w, cnt, z, z.copy(), np.full(len(ptot), m),
                             run.x_grid_start - 10 * run.rg0 * inp.gyrofactor, 0, 0.0, pxx, pxz, efl)
    if inp.input_distribution > 1:
        raise ValueError("fast push will only work with thermal input distr.")
    x_stop_rg = inp.proton_fast_transport_stop
    i_stop = int(np.argmax(prof.x_grid_rg > x_stop_rg)) - 1
    rel = run.beta0 >= BETA_REL_FL
    dr = run.u0 / prof.ux_sk[i_stop]
    if rel:
        dr *= run.gam0 / prof.gam_sf[i_stop]
    G = 5.0 / 3.0
    temp_ratio = dr**G / dr
    if KB * sp.T * temp_ratio > 4 * m * CL**2 * E_REL_PT:
        raise ValueError("Fast push cannot work: thermal particles become mildly relativistic.")
    if i_ion == 1:  # F_update!
        P0 = sum(s.n0 * s.T for s in run.species) * KB
        rho0 = sum(s.n0 * s.mass for s in run.species)
        Xi = G / (G - 1)
        for i in range(1, i_stop + 1):
            uc, gc = prof.ux_sk[i], prof.gam_sf[i]
            bc = uc / CL
            gb = gc * bc
            d = (run.gam0 * run.u0) / (gc * uc)
            rho, P = rho0 * d, P0 * d**G
            if not rel:
                Fp = rho * uc**2 * (1 + bc**2) + P * (1 + Xi * bc**2)
                Fe = rho / 2 * uc**3 * (1 + 1.25 * bc**2) + P * uc * Xi * (1 + bc**2)
            else:
                e = rho * CL**2
                Fp = P + gb**2 * (e + Xi * P)
                Fe = gb * gc * CL * (e + Xi * P) - gb * CL * e
            pxx[i - 1], pxz[i - 1], efl[i - 1] = Fp, 0.0, Fe
    ptot, w, cnt = inj_bins(inp.injection_weights, inp.n_pts_inj, inp.input_distribution, sp.T * temp_ratio, m,
                            sp.n0, inp.compat_zero_first_particle)
    u = prof.ux_sk[i_stop]
    bu = u / CL
    if rel:
        gpf = np.hypot(1.0, ptot / (m * CL))
        bpf = np.sqrt(1 - 1 / gpf**2)
        lo = np.abs((bu - bpf) / (1 - bu * bpf))
        hi = np.abs((bu + bpf) / (1 + bu * bpf))
        gfac = gpf * m
    else:
        vt = ptot / m
        lo, hi = np.abs(u - vt), np.abs(u + vt)
        gfac = np.full(len(ptot), 1.0 * m)
    return InjectionSpec(INJ_FASTPUSH_REL if rel else INJ_FASTPUSH_NONREL, ptot, w, cnt, lo, hi, gfac,
                         x_stop_rg * run.rg0, i_stop, float(u), pxx, pxz, efl)


def injection_permutation(n: int, stride: int = INJ_PERM_STRIDE) -> np.ndarray:
    """Slot -> origin index.  The reference orders the injected particles by momentum bin.  Sharding contiguous index
    blocks over GPUs then gives rank 0 the slow half and the last rank the fast half of the Maxwellian, i.e.
    unequal numbers of survivors per rank.  A fixed permutation (independent of the rank count) removes that; the order
    only decides which RNG counter a particle gets.  The permutation is `stride` strided sub-sequences laid end to end
    (indices 0,64,128,... then 1,65,...): any block of n/W particles (W = 1,2,4,8,...,64 ranks) is a fair sample of the
    Maxwellian AND stays momentum-sorted inside, which keeps the lanes of a warp on similar trajectories (a random
    shuffle costs 3 % on one GPU)."""
    return np.concatenate([np.arange(k, n, stride) for k in range(stride)]) if n else np.zeros(0, np.int64)


def expand_injection(spec: InjectionSpec, rng, shuffle: bool = False) -> InitPop:
    """Host generator: one uniform block for pb (or one triangular draw per particle), then one uniform block for phi
    (ion_init.jl:51).  `rng.random(n)` is numpy's Generator or PhiloxInjectionRng (the stream the oracle and the CUDA
    generator use)."""
    cnt = spec.bin_count
    n = spec.n
    ptot, w = np.repeat(spec.bin_ptot, cnt), np.repeat(spec.bin_weight, cnt)
    r = rng.random(n)
    if spec.mode == INJ_UPSTREAM:
        pb = ptot * 2 * (r - 0.5)
    else:
        lo, hi, gfac = np.repeat(spec.bin_lo, cnt), np.repeat(spec.bin_hi, cnt), np.repeat(spec.bin_gfac, cnt)
        # TriangularDist(a, b, b) sample = a + (b - a) sqrt(U)   
        vx = lo + (hi - lo) * np.sqrt(r)
        if spec.mode == INJ_FASTPUSH_REL:
            bu = spec.u_stop / CL
            vx_pf = (vx - bu) / (1 - vx * bu) * CL
            pb = gfac * vx_pf
        else:
            pb = gfac * (vx - spec.u_stop)
    x = np.full(n, spec.x_cm)
    grid = np.full(n, spec.grid, np.int64)
    phi = 2 * math.pi * rng.random(n)
    if shuffle:
        perm = injection_permutation(n)
        w, ptot, pb, x, grid, phi = w[perm], ptot[perm], pb[perm], x[perm], grid[perm], phi[perm]
    pop = dict(weight=w, ptot_pf=ptot, pb_pf=pb, x_cm=x, grid=grid, phi_rad=phi)
    return InitPop(pop=pop, pxx_flux=spec.pxx_flux, pxz_flux=spec.pxz_flux, energy_flux=spec.energy_flux,
                   weight_running=spec.weight_running)


def init_pop(run: Run, prof: Profile, i_ion: int, rng, shuffle: bool = False) -> InitPop:
    """init_pop (initializers.jl:977-1134), F_update! (:1157-1223) and the phase draw of
    assign_particle_properties_to_population! (ion_init.jl:51). `rng` replaces Random.Xoshiro of
    main_loops.jl:120-121 (host side, outside the replaced region)."""
    return expand_injection(injection_spec(run, prof, i_ion), rng, shuffle)


_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32-10 (Salmon et al. 2011) on numpy uint32 arrays; the generator of the library (include/mcs.h)."""
    c = [np.asarray(v, np.uint64) & np.uint64(0xFFFFFFFF) for v in np.broadcast_arrays(c0, c1, c2, c3)]
    m32 = np.uint64(0xFFFFFFFF)
    for r in range(10):
        p0, p1 = _PHILOX_M0 * c[0], _PHILOX_M1 * c[2]
        ka, kb = np.uint64((k0 + r * _PHILOX_W0) & 0xFFFFFFFF), np.uint64((k1 + r * _PHILOX_W1) & 0xFFFFFFFF)
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ ka, p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ kb, p0 & m32]
    return c


class PhiloxInjectionRng:
    """The uniforms the library's generators use for particle j of (iteration, ion): Philox counter
    (0, j, i_ion << 16, i_iter) — pcut field 0, which the transport never uses — key = seed; first `random(n)` call returns
    the first 53-bit uniform of each block (pitch angle), the second call the second (phase)."""

    def __init__(self, seed: int, i_iter: int, i_ion: int):
        self.k0, self.k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        self.c2, self.c3 = (i_ion << 16) & 0xFFFFFFFF, i_iter & 0xFFFFFFFF
        self._calls = 0
        self._blk = None

    def random(self, n: int) -> np.ndarray:
        if self._blk is None or len(self._blk[0]) != n:
            j = np.arange(n, dtype=np.uint64)
            self._blk = philox4x32_10(0, j, self.c2, self.c3, self.k0, self.k1)
            self._calls = 0
        o = self._blk
        lo, hi = (o[0], o[1]) if self._calls == 0 else (o[2], o[3])
        self._calls += 1
        return (((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64) * 2.0**-53


def synthetic_precursor(run: Run, r_sub: float = 3.0, scale_rg: float = 5.0) -> Profile:
    """A smoothed (nonlinear-shock-shaped) profile for the "nonlinear" config without running smoothers.jl.

    smooth_grid_par (smoothers.jl:54-349) stays in host Julia and is out of scope; what the transport loop sees
    of it is a velocity profile that decreases monotonically through a precursor to a weak subshock, with the
    derived arrays recomputed exactly as smoothers.jl:324-346 does.  Here u(x) = u_sub + (u0 - u_sub)(1 - e^{x/L})
    for x < 0 (L = scale_rg rg0) with a subshock of compression r_sub, and u = u0/r_comp downstream, so that the
    total compression is the run's r_comp and every upstream zone has its own flow speed (one transform_p_PSP
    per zone change)."""
    p = run.profile
    n = len(p.x_grid_rg)
    u_dn = run.u0 / run.r_comp
    u_sub = min(u_dn * r_sub, run.u0)
    ux = np.where(p.x_grid_rg < 0, u_sub + (run.u0 - u_sub) * (1 - np.exp(np.minimum(p.x_grid_rg, 0.0) / scale_rg)), u_dn)
    ux[0] = run.u0
    gsf = 1 / np.sqrt(1 - (ux / CL) ** 2)
    bef = (run.u0 - ux) / (CL - run.u0 * ux / CL)
    gef = 1 / np.sqrt(1 - bef**2)
    z = (run.gam0 * run.u0) / (gsf * ux)
    comp = 1 + (np.sqrt(1 / 3 + 2 / 3 * z**2) - 1) * run.inp.b_field_turbulence
    bt = run.bmag0 * (1 + (comp - 1) * run.inp.b_field_amplify)
    if run.inp.use_custom_epsB:
        n0 = sum(s.n0 * s.mass for s in run.species) / MP
        e0 = n0 * MP * CL**2
        ed = (run.F_energy_upstream + run.gam0 * run.u0 * e0) / ux - run.F_px_upstream
        bt = np.sqrt(np.abs(8 * math.pi * p.epsB * ed))
    return Profile(x_grid_rg=p.x_grid_rg.copy(), x_grid_cm=p.x_grid_cm.copy(), ux_sk=ux, uz_sk=np.zeros(n), utot=ux.copy(),
                   gam_sf=gsf, gam_ef=gef, beta_ef=bef, btot=bt, theta=p.theta.copy(), epsB=p.epsB.copy())


# ------------------------------------------------------------------------------------------------
# Inputs of the pressure consumer: bin centres and zone populations, host side as in the reference
def psd_bounds(run: Run, as_written: bool = True):
    """(psd_mom_bounds [0..M+1], psd_theta_bounds [0..T+1]) of set_psd_mom_bins / set_psd_angle_bins
    (/root/reference/src/initializers.jl:216-285).

    as_written: the reference ends set_psd_angle_bins with `sort!(psd_theta_bounds)`, which interleaves the angle part
    (radians, ascending) with the cosine part (descending from psd_cos_fine to -1) — the docstring above it promises the
    unsorted order.  False returns that intended order."""
    bpd_p, bpd_t = run.inp.num_psd_bins_per_decade
    M = run.num_psd_mom_bins
    log_p_min = math.log10(run.psd_mom_min / (MP * CL))
    mom = np.concatenate(([-99.0], log_p_min + np.arange(M + 1) / bpd_p))
    lin = run.inp.psd_linear_cosine_bins
    n_log = run.num_psd_theta_bins - lin
    root = 10.0 ** (1.0 / bpd_t)
    th = np.concatenate(([1.0e-99], run.psd_theta_min * root ** np.arange(n_log), run.psd_cos_fine - run.delta_cos * np.arange(lin + 1)))
    if as_written:
        th = np.sort(th)
    return mom, th


def thermo_inputs(run: Run, prof: Profile, i_ion: int, jet_rad_pc: float = 0.438, jet_open_ang_deg: float = 5.0,
                  as_written: bool = False):
    """(cos_center [T+1], pt_center [M+1], zone_pop [n_grid]) for `mcs_thermo`.

    cos_center / pt_center follow thermo_calcs.jl:55-82, zone_pop follows set_grid_volumes!
    (particle_counter.jl:1463-1523; jet radius / opening angle default to the bundled mc_in.toml:192-195).
    as_written keeps two quirks of the reference: the sorted theta bounds (see psd_bounds) and
    `pt_center = exp10(log_p_mid) g cm/s`, although psd_mom_bounds holds log10(p / m_p c) — i.e. momenta too large by
    1/(m_p c).  The default (False) uses the unsorted bounds and multiplies by m_p c, which is what the re-binning against
    psd_mom_min [g cm/s] needs to land inside the grid."""
    mom, th = psd_bounds(run, as_written=as_written)
    T, M, lin = run.num_psd_theta_bins, run.num_psd_mom_bins, run.inp.psd_linear_cosine_bins
    cosc = np.zeros(T + 1)
    for j in range(T + 1):
        if j > T - lin:
            hi, lo = th[j], th[j + 1]
        elif j == T - lin:
            hi, lo = math.cos(th[j]), th[j + 1]
        else:
            hi, lo = math.cos(th[j]), math.cos(th[j + 1])
        cosc[j] = -(lo + hi) / 2
    ptc = 10.0 ** ((mom[:-1] + mom[1:]) / 2)
    if not as_written:
        ptc = ptc * (MP * CL)
    # set_grid_volumes!
    ng, ish, g0 = run.n_grid, run.i_shock, run.gam0
    xg = np.asarray(prof.x_grid_cm, float)
    dx = np.diff(xg)                       # dx[i] = x[i+1] - x[i], zone i = 1..n_grid
    sph = (1 - math.cos(math.radians(jet_open_ang_deg))) / 2   # parse_jet_frac, data_input.jl:158-166
    rad_cm = jet_rad_pc * PC_CM
    area = np.zeros(ng + 1)
    r_min = rad_cm - xg[ish]
    for i in range(ish - 1, 0, -1):
        r_max = r_min + dx[i] / g0
        area[i] = math.pi * (r_max + r_min) ** 2 * sph
        r_min = r_max
    r_max = rad_cm - xg[ish]
    for i in range(ish, ng + 1):
        r_min = r_max - dx[i] / g0
        area[i] = math.pi * (r_max + r_min) ** 2 * sph
        r_max = r_min
    F_up = g0 * run.species[i_ion].n0 * run.beta0 * CL
    zone_pop = np.array([F_up * area[i] * dx[i] / float(prof.ux_sk[i]) for i in range(1, ng + 1)])
    return cosc, ptc, zone_pop
