"""ctypes mirror of include/mcs.h and a thin object wrapper around one McsHandle.

The wrapper is backend-agnostic on purpose: it binds whatever shared library it is given that
exports the mcs_* C-ABI.  The product (`engine.load_cuda_engine`) only ever gives it the CUDA
library; tests hand it the CPU oracle to compare against.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

MCS_ABI_VERSION = 1
NA_C = 100
PSD_MAX = 200
MAX_IONS = 8
MAX_XSPEC = 16

RNG_PHILOX = 0
RNG_REPLAY = 1
COMPAT_RETRO_KEEP_NEW_PITCH = 1
COMPAT_DEFAULT = COMPAT_RETRO_KEEP_NEW_PITCH

FATE_SAVED, FATE_DOWNSTREAM, FATE_FEB_PMAX, FATE_AGE, FATE_ZERO_ENERGY, FATE_ERROR = range(6)

_d, _i32, _i64, _u64, _u32 = C.c_double, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32
_pd, _pi64, _pu8, _pi32 = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)


class McsConfig(C.Structure):
    _fields_ = [
        ("abi_version", _i32), ("device", _i32),
        ("mp_g", _d), ("c_cms", _d), ("qcgs_esu", _d), ("E_rel_pt", _d), ("rad_loss_fac", _d),
        ("gam0", _d), ("beta0", _d), ("u0", _d), ("u2", _d), ("bmag2", _d),
        ("pe_crit", _d), ("gam_e_crit", _d), ("eta_mfp", _d),
        ("psd_mom_min", _d), ("psd_cos_fine", _d), ("delta_cos", _d), ("psd_theta_min", _d),
        ("psd_bins_per_dec_mom", _i32), ("psd_bins_per_dec_theta", _i32),
        ("num_psd_mom_bins", _i32), ("num_psd_theta_bins", _i32),
        ("energy_transfer_frac", _d),
        ("feb_upstream", _d), ("feb_downstream", _d), ("x_grid_stop", _d),
        ("B_CMBz", _d),
        ("xn_per_fine", _d), ("xn_per_coarse", _d),
        ("age_max", _d),
        ("n_grid", _i32), ("i_grid_feb", _i32), ("i_shock", _i32), ("n_ions", _i32),
        ("n_pts_max", _i64), ("na_cr", _i64),
        ("n_xspec", _i32), ("x_spec", _d * MAX_XSPEC),
        ("n_tcuts", _i32), ("tcuts", _d * NA_C),
        ("inj_fracs", _d * MAX_IONS),
        ("do_rad_losses", _i32), ("do_retro", _i32), ("do_tcuts", _i32), ("dont_DSA", _i32),
        ("dont_scatter", _i32), ("use_custom_frg", _i32), ("use_custom_epsB", _i32),
        ("helix_cap", _i32), ("retro_cap", _i64), ("seed", _u64), ("compat", _u32),
        ("rng_mode", _i32), ("threads", _i32), ("bin_thermal", _i32), ("dynamic_queue", _i32),
        ("det_tallies", _i32),
    ]


class McsSpecies(C.Structure):
    _fields_ = [("aa", _d), ("zz_esu", _d), ("n0", _d), ("pmax_cutoff", _d), ("electron_weight_fac", _d)]


class McsTallies(C.Structure):
    _fields_ = [
        ("pxx_flux", _pd), ("pxz_flux", _pd), ("energy_flux", _pd), ("psd", _pd), ("num_crossings", _pi64),
        ("n_cr_count", _i64), ("n_cr_overflow", _i64),
        ("therm_grid", _pi64), ("therm_px_sk", _pd), ("therm_ptot_sk", _pd), ("therm_weight", _pd),
        ("esc_psd_feb_upstream", _pd), ("esc_psd_feb_downstream", _pd),
        ("esc_energy_eff", _pd), ("esc_num_eff", _pd), ("weight_coupled", _pd), ("spectra_coupled", _pd),
        ("energy_transfer_pool", _pd), ("spectra_sf", _pd), ("spectra_pf", _pd),
        ("therm_d2N_sf", _pd), ("therm_d2N_pf", _pd), ("dNdp_cr_sf", _pd),
        ("esc_flux", _d), ("px_esc_feb", _d), ("energy_esc_feb", _d),
        ("sum_P_downstream", _d), ("sum_KE_downstream", _d),
        ("px_esc_upstream", _d), ("energy_esc_upstream", _d),
        ("n_helix_steps", _i64), ("n_retro_steps", _i64),
        ("n_warn_pperp", _i64), ("n_warn_psd_mom", _i64), ("n_neg_sqrt", _i64), ("n_retro_capped", _i64),
        ("n_errors", _i64), ("n_fate", _i64 * 6),
    ]


class McsPopulation(C.Structure):
    _fields_ = [
        ("weight", _pd), ("ptot_pf", _pd), ("pb_pf", _pd), ("x_cm", _pd), ("xn_per", _pd), ("prp_x_cm", _pd),
        ("acctime_sec", _pd), ("phi_rad", _pd), ("grid", _pi64), ("tcut", _pi64), ("downstream", _pu8), ("inj", _pu8),
    ]


class McsInjection(C.Structure):
    _fields_ = [
        ("n_bins", _i32), ("mode", _i32), ("bin_ptot", _pd), ("bin_weight", _pd), ("bin_start", _pi64),
        ("bin_lo", _pd), ("bin_hi", _pd), ("bin_gfac", _pd), ("x_cm", _d), ("u_stop", _d), ("grid", _i64),
        ("perm_stride", _i32), ("reserved", _i32),
    ]


class McsTraceRec(C.Structure):
    _fields_ = [
        ("x_cm", _d), ("ptot_pf", _d), ("pb_pf", _d), ("phi_rad", _d), ("acctime_sec", _d), ("prp_x_cm", _d),
        ("i_grid", _i32), ("helix_count", _i32), ("flags", _i32), ("n_draws", _i32),
    ]


class McsTiming(C.Structure):
    _fields_ = [
        ("transport_ms", _d), ("split_ms", _d), ("reduce_ms", _d), ("h2d_ms", _d), ("d2h_ms", _d), ("comm_ms", _d), ("ion_loop_ms", _d),
        ("transport_launches", _i64), ("other_launches", _i64), ("local_steps", _i64), ("local_particles", _i64),
        ("local_reds", _i64),
    ]


TRACE_DTYPE = np.dtype(
    [("x_cm", "f8"), ("ptot_pf", "f8"), ("pb_pf", "f8"), ("phi_rad", "f8"), ("acctime_sec", "f8"), ("prp_x_cm", "f8"),
     ("i_grid", "i4"), ("helix_count", "i4"), ("flags", "i4"), ("n_draws", "i4")]
)

POP_F64 = ("weight", "ptot_pf", "pb_pf", "x_cm", "xn_per", "prp_x_cm", "acctime_sec", "phi_rad")
POP_I64 = ("grid", "tcut")
POP_U8 = ("downstream", "inj")

# every symbol include/mcs.h declares
ABI_SYMBOLS = (
    "mcs_last_error", "mcs_backend", "mcs_abi_sizes", "mcs_default_config", "mcs_create", "mcs_destroy",
    "mcs_comm_unique_id", "mcs_comm_init", "mcs_set_profile", "mcs_begin_ion", "mcs_begin_ion_generate", "mcs_run_pcut", "mcs_split", "mcs_split_explicit",
    "mcs_run_ion", "mcs_end_ion", "mcs_get_population", "mcs_population_size", "mcs_get_fates",
    "mcs_replay_set_stream", "mcs_trace_enable", "mcs_trace_get", "mcs_get_timing", "mcs_measure_fp64_peak",
    "mcs_measure_atomic_peak", "mcs_measure_scatter_peak", "mcs_selftest_math", "mcs_thermo",
)


class McsThermoIn(C.Structure):
    _fields_ = [("cos_center", _pd), ("pt_center", _pd), ("zone_pop", _pd), ("temperature_K", _d),
                ("psd", _pd), ("therm_d2N_pf", _pd), ("num_crossings", _pi64)]


class McsError(RuntimeError):
    pass


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach argtypes/restypes; raises AttributeError if a declared symbol is missing."""
    for s in ABI_SYMBOLS:
        getattr(lib, s)
    H = C.c_void_p
    lib.mcs_last_error.restype = C.c_char_p
    lib.mcs_backend.restype = C.c_char_p
    lib.mcs_abi_sizes.argtypes = [C.POINTER(_i32 * 6)]
    lib.mcs_default_config.argtypes = [C.POINTER(McsConfig)]
    lib.mcs_default_config.restype = None
    lib.mcs_create.argtypes = [C.POINTER(McsConfig), C.POINTER(H)]
    lib.mcs_destroy.argtypes = [H]
    lib.mcs_comm_unique_id.argtypes = [C.c_void_p]
    lib.mcs_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.mcs_set_profile.argtypes = [H, _i32] + [_pd] * 11
    lib.mcs_begin_ion.argtypes = [H, _i32, _i32, C.POINTER(McsSpecies), _i64, _i64, C.POINTER(McsPopulation)]
    lib.mcs_begin_ion_generate.argtypes = [H, _i32, _i32, C.POINTER(McsSpecies), _i64, _i64, C.POINTER(McsInjection)]
    lib.mcs_run_pcut.argtypes = [H, _i32, _d, _d, _pi64, _pi64]
    lib.mcs_split.argtypes = [H, _i64, _pi64, _pi64, _pi64]
    lib.mcs_split_explicit.argtypes = [H, _i64, _i64, _pi64]
    lib.mcs_run_ion.argtypes = [H, _pd, _i32, _d, _i64, _i64, _pi32, _pi64, _pi64]
    lib.mcs_end_ion.argtypes = [H, C.POINTER(McsTallies)]
    lib.mcs_get_population.argtypes = [H, _i32, _i64, C.POINTER(McsPopulation), _pu8]
    lib.mcs_population_size.argtypes = [H]
    lib.mcs_population_size.restype = _i64
    lib.mcs_get_fates.argtypes = [H, _i64, _pi32, _pi32, _pi64, _pi64]
    lib.mcs_replay_set_stream.argtypes = [H, _pd, _pi64, _i64]
    lib.mcs_trace_enable.argtypes = [H, _pi64, _i32, _i32]
    lib.mcs_trace_get.argtypes = [H, C.c_void_p, _pi32]
    lib.mcs_get_timing.argtypes = [H, C.POINTER(McsTiming), _i32]
    lib.mcs_measure_fp64_peak.argtypes = [H, _pd]
    lib.mcs_measure_atomic_peak.argtypes = [H, _i64, _pd]
    lib.mcs_measure_scatter_peak.argtypes = [H, _pd]
    lib.mcs_selftest_math.argtypes = [H, _i64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.mcs_thermo.argtypes = [H, C.POINTER(McsThermoIn), _pd, _pd, _pd, _pd]
    sizes = (_i32 * 6)()
    lib.mcs_abi_sizes(C.byref(sizes))
    want = [C.sizeof(t) for t in (McsConfig, McsSpecies, McsTallies, McsPopulation, McsTraceRec, McsTiming)]
    if list(sizes) != want:
        raise McsError(f"ABI struct size mismatch: library {list(sizes)} vs ctypes mirror {want}")
    return lib


@dataclass
class Tallies:
    """Per-ion pure sums with the reference's shapes (column-major: first axis fastest)."""
    pxx_flux: np.ndarray
    pxz_flux: np.ndarray
    energy_flux: np.ndarray
    psd: np.ndarray  # [n_grid, T+2, M+2] C-order view of the (M+2, T+2, n_grid) column-major array
    num_crossings: np.ndarray
    therm_grid: np.ndarray
    therm_px_sk: np.ndarray
    therm_ptot_sk: np.ndarray
    therm_weight: np.ndarray
    n_cr_overflow: int
    esc_psd_feb_upstream: np.ndarray  # [jt, ip]
    esc_psd_feb_downstream: np.ndarray
    esc_energy_eff: np.ndarray
    esc_num_eff: np.ndarray
    weight_coupled: np.ndarray
    spectra_coupled: np.ndarray  # [tcut, ip]
    energy_transfer_pool: np.ndarray
    spectra_sf: np.ndarray
    spectra_pf: np.ndarray
    therm_d2N_sf: np.ndarray = None  # [n_grid, M+2, T+2] (C order of the reference's [jth, k, i])
    therm_d2N_pf: np.ndarray = None
    dNdp_cr_sf: np.ndarray = None    # [n_grid, M+2]
    scalars: dict = field(default_factory=dict)
    stats: dict = field(default_factory=dict)


class Engine:
    """One McsHandle. Method names follow include/mcs.h."""

    def __init__(self, lib: C.CDLL, cfg: McsConfig):
        self.lib = lib
        self.cfg = cfg
        self._h = C.c_void_p()
        self._check(lib.mcs_create(C.byref(cfg), C.byref(self._h)))
        self.n_grid = cfg.n_grid
        self.M = cfg.num_psd_mom_bins
        self.T = cfg.num_psd_theta_bins

    # -- plumbing -------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise McsError(f"mcs error {rc}: {self.lib.mcs_last_error().decode()}")

    def close(self):
        if self._h:
            self.lib.mcs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def backend(self) -> str:
        return self.lib.mcs_backend().decode()

    # -- ABI ------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.lib.mcs_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, nranks: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._check(self.lib.mcs_comm_init(self._h, rank, nranks, buf))

    def set_profile(self, prof, eps_target=None, energy_recv_pool=None):
        """prof: object with x_grid_cm, ux_sk, uz_sk, utot, gam_sf, gam_ef, beta_ef, btot, theta (n_grid+2 each)."""
        names = ("x_grid_cm", "ux_sk", "uz_sk", "utot", "gam_sf", "gam_ef", "beta_ef", "btot", "theta")
        arrs = [np.ascontiguousarray(getattr(prof, n), dtype=np.float64) for n in names]
        for a in arrs:
            if a.shape != (self.n_grid + 2,):
                raise McsError("profile arrays must have n_grid+2 nodes")
        opt = []
        for a in (eps_target, energy_recv_pool):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                if a.shape != (self.n_grid,):
                    raise McsError("eps_target / energy_recv_pool must have n_grid entries")
            opt.append(a)
        self._keep = arrs + opt
        self._check(self.lib.mcs_set_profile(self._h, self.n_grid, *[_ptr(a, _pd) for a in arrs],
                                             *[_ptr(a, _pd) for a in opt]))

    @staticmethod
    def _pop_struct(pop: dict, n: int, writable=False):
        st = McsPopulation()
        keep = {}
        for names, dt, typ in ((POP_F64, np.float64, _pd), (POP_I64, np.int64, _pi64), (POP_U8, np.uint8, _pu8)):
            for nm in names:
                a = pop.get(nm)
                if a is None:
                    continue
                if not writable:
                    a = np.ascontiguousarray(a, dtype=dt)
                if a.dtype != dt or a.size < n or not a.flags.c_contiguous:
                    raise McsError(f"population field {nm}: need contiguous {dt} with >= {n} entries")
                keep[nm] = a
                setattr(st, nm, a.ctypes.data_as(typ))
        return st, keep

    def begin_ion(self, i_iter: int, i_ion: int, species: McsSpecies, pop: dict, first_global: int = 0):
        n = len(pop["weight"])
        st, keep = self._pop_struct(pop, n)
        self._check(self.lib.mcs_begin_ion(self._h, i_iter, i_ion, C.byref(species), n, first_global, C.byref(st)))

    def begin_ion_generate(self, i_iter: int, i_ion: int, species: McsSpecies, spec, first_global: int = 0,
                           n_local: int | None = None, shuffle: bool = False):
        """mcs_begin_ion_generate from a problem.InjectionSpec: the population is produced inside the library."""
        keep = [np.ascontiguousarray(a, np.float64) for a in (spec.bin_ptot, spec.bin_weight, spec.bin_lo, spec.bin_hi,
                                                              spec.bin_gfac)]
        start = np.concatenate(([0], np.cumsum(spec.bin_count))).astype(np.int64)
        n_total = int(start[-1])
        inj = McsInjection(n_bins=len(keep[0]), mode=spec.mode, bin_ptot=_ptr(keep[0], _pd), bin_weight=_ptr(keep[1], _pd),
                           bin_start=_ptr(start, _pi64), bin_lo=_ptr(keep[2], _pd), bin_hi=_ptr(keep[3], _pd),
                           bin_gfac=_ptr(keep[4], _pd), x_cm=spec.x_cm, u_stop=spec.u_stop, grid=spec.grid,
                           perm_stride=64 if shuffle else 0, reserved=0)
        n_local = n_total - first_global if n_local is None else n_local
        self._check(self.lib.mcs_begin_ion_generate(self._h, i_iter, i_ion, C.byref(species), first_global, n_local,
                                                    C.byref(inj)))
        return n_total

    def run_pcut(self, i_pcut: int, pcut: float, pcut_prev: float):
        ns, nst = _i64(), _i64()
        self._check(self.lib.mcs_run_pcut(self._h, i_pcut, pcut, pcut_prev, C.byref(ns), C.byref(nst)))
        return ns.value, nst.value

    def split(self, n_pts_target: int):
        a, b, m = _i64(), _i64(), _i64()
        self._check(self.lib.mcs_split(self._h, n_pts_target, C.byref(a), C.byref(b), C.byref(m)))
        return a.value, b.value, m.value

    def split_explicit(self, i_mult: int, first_global_child: int) -> int:
        a = _i64()
        self._check(self.lib.mcs_split_explicit(self._h, i_mult, first_global_child, C.byref(a)))
        return a.value

    def run_ion(self, pcuts, p_pcut_hi: float, n_pts_pcut: int, n_pts_pcut_hi: int):
        pc = np.ascontiguousarray(pcuts, dtype=np.float64)
        n_run = _i32()
        used = np.zeros(len(pc), np.int64)
        saved = np.zeros(len(pc), np.int64)
        self._check(self.lib.mcs_run_ion(self._h, _ptr(pc, _pd), len(pc), p_pcut_hi, n_pts_pcut, n_pts_pcut_hi,
                                         C.byref(n_run), _ptr(used, _pi64), _ptr(saved, _pi64)))
        return n_run.value, used[: n_run.value], saved[: n_run.value]

    def end_ion(self, want_psd=True, want_log=True) -> Tallies:
        ng, M2, T2, e1 = self.n_grid, self.M + 2, self.T + 2, PSD_MAX + 1
        nx = max(self.cfg.n_xspec, 0)
        L = max(int(self.cfg.na_cr), 1) if want_log else 0
        z = np.zeros
        t = Tallies(
            pxx_flux=z(ng), pxz_flux=z(ng), energy_flux=z(ng),
            psd=z((ng, T2, M2)) if want_psd else None, num_crossings=z(ng, np.int64),
            therm_grid=z(L, np.int64) if want_log else None, therm_px_sk=z(L) if want_log else None,
            therm_ptot_sk=z(L) if want_log else None, therm_weight=z(L) if want_log else None, n_cr_overflow=0,
            esc_psd_feb_upstream=z((e1, e1)), esc_psd_feb_downstream=z((e1, e1)),
            esc_energy_eff=z(e1), esc_num_eff=z(e1), weight_coupled=z(NA_C), spectra_coupled=z((NA_C, e1)),
            energy_transfer_pool=z(ng), spectra_sf=z((max(nx, 1), e1)), spectra_pf=z((max(nx, 1), e1)),
        )
        if self.cfg.bin_thermal:
            t.therm_d2N_sf, t.therm_d2N_pf, t.dNdp_cr_sf = z((ng, M2, T2)), z((ng, M2, T2)), z((ng, M2))
        st = McsTallies()
        for nm, typ in (("pxx_flux", _pd), ("pxz_flux", _pd), ("energy_flux", _pd), ("psd", _pd),
                        ("num_crossings", _pi64), ("therm_grid", _pi64), ("therm_px_sk", _pd),
                        ("therm_ptot_sk", _pd), ("therm_weight", _pd), ("esc_psd_feb_upstream", _pd),
                        ("esc_psd_feb_downstream", _pd), ("esc_energy_eff", _pd), ("esc_num_eff", _pd),
                        ("weight_coupled", _pd), ("spectra_coupled", _pd), ("energy_transfer_pool", _pd),
                        ("spectra_sf", _pd), ("spectra_pf", _pd), ("therm_d2N_sf", _pd), ("therm_d2N_pf", _pd),
                        ("dNdp_cr_sf", _pd)):
            setattr(st, nm, _ptr(getattr(t, nm), typ))
        self._check(self.lib.mcs_end_ion(self._h, C.byref(st)))
        if want_log:
            n = int(st.n_cr_count)
            t.therm_grid, t.therm_px_sk = t.therm_grid[:n], t.therm_px_sk[:n]
            t.therm_ptot_sk, t.therm_weight = t.therm_ptot_sk[:n], t.therm_weight[:n]
        t.n_cr_overflow = int(st.n_cr_overflow)
        t.scalars = {k: getattr(st, k) for k in ("esc_flux", "px_esc_feb", "energy_esc_feb", "sum_P_downstream",
                                                 "sum_KE_downstream", "px_esc_upstream", "energy_esc_upstream")}
        t.stats = {k: int(getattr(st, k)) for k in ("n_cr_count", "n_cr_overflow", "n_helix_steps", "n_retro_steps",
                                                    "n_warn_pperp", "n_warn_psd_mom", "n_neg_sqrt", "n_retro_capped",
                                                    "n_errors")}
        t.stats["n_fate"] = [int(v) for v in st.n_fate]
        return t

    def thermo(self, cos_center, pt_center, zone_pop, temperature_K: float, psd=None, therm_d2N_pf=None, num_crossings=None):
        """thermo_calcs on the tallies of the ion just ended (or on the three host arrays, given together):
        returns (P_psd_par, P_psd_perp, energy_density_psd, d2N_pop), [n_grid] each."""
        ng = self.n_grid
        f = lambda a, n: np.ascontiguousarray(a, dtype=np.float64).reshape(-1)[:n] if a is not None else None
        keep = [f(cos_center, self.T + 1), f(pt_center, self.M + 1), f(zone_pop, ng), f(psd, None), f(therm_d2N_pf, None),
                np.ascontiguousarray(num_crossings, dtype=np.int64) if num_crossings is not None else None]
        if len(keep[0]) != self.T + 1 or len(keep[1]) != self.M + 1 or len(keep[2]) != ng:
            raise McsError("thermo: cos_center [T+1], pt_center [M+1], zone_pop [n_grid]")
        n_psd = ng * (self.M + 2) * (self.T + 2)
        for a in keep[3:5]:
            if a is not None and a.size != n_psd:
                raise McsError("thermo: psd / therm_d2N_pf must hold (M+2)(T+2)n_grid cells")
        if keep[5] is not None and keep[5].size != ng:
            raise McsError("thermo: num_crossings [n_grid]")
        tin = McsThermoIn(_ptr(keep[0], _pd), _ptr(keep[1], _pd), _ptr(keep[2], _pd), float(temperature_K),
                          _ptr(keep[3], _pd), _ptr(keep[4], _pd), _ptr(keep[5], _pi64))
        out = [np.zeros(ng) for _ in range(4)]
        self._check(self.lib.mcs_thermo(self._h, C.byref(tin), *[_ptr(o, _pd) for o in out]))
        return tuple(out)

    def population_size(self) -> int:
        return int(self.lib.mcs_population_size(self._h))

    def get_population(self, which: int = 0, n: int | None = None):
        n = self.population_size() if n is None else n
        pop = {nm: np.zeros(n, np.float64) for nm in POP_F64}
        pop.update({nm: np.zeros(n, np.int64) for nm in POP_I64})
        pop.update({nm: np.zeros(n, np.uint8) for nm in POP_U8})
        l_save = np.zeros(n, np.uint8)
        st, keep = self._pop_struct(pop, n, writable=True)
        self._check(self.lib.mcs_get_population(self._h, which, n, C.byref(st), _ptr(l_save, _pu8)))
        pop["l_save"] = l_save
        return pop

    def get_fates(self, n: int):
        fate, helix = np.zeros(n, np.int32), np.zeros(n, np.int32)
        retro, draws = np.zeros(n, np.int64), np.zeros(n, np.int64)
        self._check(self.lib.mcs_get_fates(self._h, n, _ptr(fate, _pi32), _ptr(helix, _pi32), _ptr(retro, _pi64),
                                           _ptr(draws, _pi64)))
        return dict(fate=fate, helix_count=helix, retro_steps=retro, n_draws=draws)

    def replay_set_stream(self, u, offsets):
        u = np.ascontiguousarray(u, dtype=np.float64)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._check(self.lib.mcs_replay_set_stream(self._h, _ptr(u, _pd), _ptr(off, _pi64), len(off) - 1))

    def trace_enable(self, idx, max_steps: int):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        self._trace = (len(idx), max_steps)
        self._check(self.lib.mcs_trace_enable(self._h, _ptr(idx, _pi64), len(idx), max_steps))

    def trace_get(self):
        n, ms = self._trace
        recs = np.zeros((n, ms), TRACE_DTYPE)
        cnt = np.zeros(n, np.int32)
        self._check(self.lib.mcs_trace_get(self._h, recs.ctypes.data_as(C.c_void_p), _ptr(cnt, _pi32)))
        return [recs[i, : cnt[i]] for i in range(n)]

    def timing(self, reset=False) -> dict:
        t = McsTiming()
        self._check(self.lib.mcs_get_timing(self._h, C.byref(t), 1 if reset else 0))
        return {k: getattr(t, k) for k, _ in McsTiming._fields_}

    def measure_fp64_peak(self) -> float:
        v = _d()
        self._check(self.lib.mcs_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def measure_scatter_peak(self) -> float:
        v = _d()
        self._check(self.lib.mcs_measure_scatter_peak(self._h, C.byref(v)))
        return v.value

    def selftest_math(self, n: int) -> tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.mcs_selftest_math(self._h, n, C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_atomic_peak(self, n_cells: int) -> float:
        v = _d()
        self._check(self.lib.mcs_measure_atomic_peak(self._h, n_cells, C.byref(v)))
        return v.value


def default_config(lib: C.CDLL) -> McsConfig:
    cfg = McsConfig()
    lib.mcs_default_config(C.byref(cfg))
    return cfg
