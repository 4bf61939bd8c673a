"""Host-side consumers of the per-ion tallies: the grid table `mc_grid.dat`.

`mc_grid_table` restates the diagnostic half of `smooth_grid_par` (/root/reference/src/smoothers.jl:76-272): from the
shock profile and the flux tallies handed back by `mcs_end_ion` it forms, per grid zone, the 2 + 33 columns the reference
writes to `mc_grid.dat` — positions (linear / log, r_g0 / cm), momentum and energy flux normalised to the far-upstream
values (with the electromagnetic parts of Double et al. 2004 Eqs. 27-28), flow speed, field, compression, and the
pressures from the flux equations and from the PSD.  The profile UPDATE of the same function (`new_velocity_profile`,
smoothers.jl:284-346) stays with the caller: north_star keeps smoothing in host Julia.

The reference builds each row with `string(i_iter, i, x, ...)`, i.e. it concatenates the numbers WITHOUT separators
(smoothers.jl:234-272) — the file cannot be read back as written.  `write_mc_grid` writes the same values separated by
blanks, one row per zone, followed by the reference's blank plot-separator line (`print_plot_vals`).
"""
from __future__ import annotations

import math

import numpy as np

from . import problem

MC_GRID_COLUMNS = (
    "i_iter", "i", "x_rg", "x_log_rg", "x_cm", "x_log_cm", "pxx_norm", "log_pxx_norm", "pxz_norm", "log_pxz_norm", "en_norm",
    "log_en_norm", "ux_norm", "log_ux_norm", "uz_norm", "log_uz_norm", "B_G", "log_B", "theta_deg", "gam_sf", "inv_density_ratio",
    "density_ratio", "log_P_px", "log_P_en", "log_P_psd_par", "log_P_psd_perp", "log_P_tot_MC", "P_aniso", "log_P_px_tp",
    "log_P_en_tp", "log_P0", "log_1_minus_qesc_px", "log_1_minus_qesc_en", "eps_B", "log_eps_B",
)


def _log10(x):
    return math.log10(x) if x > 0 else float("nan")


def mc_grid_table(run: problem.Run, prof: problem.Profile, pxx_flux, energy_flux, *, i_iter: int = 1, gamma_grid=None,
                  P_psd_par=None, P_psd_perp=None, q_esc_cal_px: float = 0.0, q_esc_cal_energy: float = 0.0,
                  gamma2: float | None = None) -> np.ndarray:
    """Rows of mc_grid.dat, one per zone i = 1..n_grid: array [n_grid, 35] in the order of MC_GRID_COLUMNS.

    pxx_flux / energy_flux are the arrays the reference holds at the end of the ion loop (fast-push prefill + tallies +
    1e-99, i.e. `driver.main_loops(...)[it][ion]["pxx_flux"]`).  gamma_grid[i, 0:2] are the adiabatic indices before /
    after the iteration (5/3 when omitted: a cold test-particle flow); P_psd_par / P_psd_perp come from thermo_calcs
    (1e-99 when omitted, the value main_loops.jl:62 initialises them to)."""
    ng = run.n_grid
    c = problem.CL
    sp = run.species
    n0 = sum(s.n0 * s.aa for s in sp)                      # smoothers.jl:84-86
    P0 = sum(s.n0 * s.T for s in sp) * problem.KB
    e0 = n0 * problem.MP * c * c
    G2 = (5.0 / 3.0) if gamma2 is None else gamma2
    gam_grid = np.full((ng + 2, 2), 5.0 / 3.0) if gamma_grid is None else np.asarray(gamma_grid, float)
    Ppar = np.full(ng + 1, 1.0e-99) if P_psd_par is None else np.concatenate(([0.0], np.asarray(P_psd_par, float)))
    Pperp = np.full(ng + 1, 1.0e-99) if P_psd_perp is None else np.concatenate(([0.0], np.asarray(P_psd_perp, float)))
    eps_b = getattr(prof, "epsB", None)
    rows = np.zeros((ng, len(MC_GRID_COLUMNS)))
    u0, b0, g0, u2, b2, g2 = run.u0, run.beta0, run.gam0, run.u2, run.beta2, run.gam2
    # test-particle pressures (smoothers.jl:215-226), evaluated once
    P_px_tp = (run.F_px_upstream - g2 * b2 * g0 * e0) / (1 + (g2 * b2) ** 2 * G2 / (G2 - 1))
    P_en_tp = (run.F_energy_upstream + g0 * u0 * e0 * (1 - g2)) / (g2**2 * u2 * G2 / (G2 - 1))
    for i in range(1, ng + 1):
        xr = float(prof.x_grid_rg[i])
        x_log = -math.log10(-xr) if xr < -1 else (math.log10(xr) if xr > 1 else 0.0)                  # :108-114
        x_log_cm = -math.log10(-xr * run.rg0) if xr < 0 else (math.log10(xr * run.rg0) if xr > 0 else 0.0)
        ux, gsf = float(prof.ux_sk[i]), float(prof.gam_sf[i])
        b_ux, b_uz = ux / c, float(prof.uz_sk[i]) / c
        B, th = float(prof.btot[i]), float(prof.theta[i])
        Xi_pre = gam_grid[i, 0] / (gam_grid[i, 0] - 1)
        ux_norm, uz_norm = ux / float(prof.ux_sk[1]), 1.0e-99                                            # :140-141
        g2f, gb = gsf * gsf, gsf * b_ux
        dens = g0 * b0 / gb
        Bx, Bz = B * math.cos(th), B * math.sin(th)
        pxx_EM = gb**2 / (8 * math.pi) * B**2 + g2f / (8 * math.pi) * (Bz**2 - Bx**2) - (g2f - gsf) / (2 * math.pi) * (b_uz / b_ux) * Bx * Bz
        en_EM = g2f / (4 * math.pi) * b_ux * c * Bz**2 - (2 * g2f - gsf) / (4 * math.pi) * b_uz * c * Bx * Bz
        pxx_norm = (float(pxx_flux[i - 1]) + pxx_EM) / run.F_px_upstream                                 # :160-164
        en_norm = (float(energy_flux[i - 1]) + en_EM) / run.F_energy_upstream
        pxx_log = max(_log10(abs(pxx_norm)), -99.0) if pxx_norm != 0 else -99.0
        en_log = max(_log10(en_norm), -99.0) if en_norm > 0 else -99.0
        P_px = (run.F_px_upstream * (1.0 - q_esc_cal_px) - gb**2 * dens * e0) / (1 + gb**2 * Xi_pre)    # :183-185
        P_en = (run.F_energy_upstream * (1 - q_esc_cal_energy) + g0 * b0 * c * e0 - g2f * ux * dens * e0) / (g2f * ux * Xi_pre)
        P_px, P_en = max(P_px, 1.0e-99), max(P_en, 1.0e-99)
        P_tot = Ppar[i] + Pperp[i]
        eb = float(eps_b[i]) if eps_b is not None else float("nan")
        rows[i - 1] = (
            i_iter, i, xr, x_log, float(prof.x_grid_cm[i]), x_log_cm, pxx_norm, pxx_log, 1.0e-99, -99.0, en_norm, en_log,
            ux_norm, _log10(ux_norm), uz_norm, _log10(uz_norm), B, _log10(B), math.degrees(th), gsf, 1 / dens, dens,
            _log10(P_px), _log10(P_en), _log10(Ppar[i]), _log10(Pperp[i]), _log10(P_tot), 2 * Ppar[i] / Pperp[i],
            _log10(P_px_tp) if P_px_tp > 0 else float("nan"), _log10(P_en_tp) if P_en_tp > 0 else float("nan"), _log10(P0),
            _log10(1 - q_esc_cal_px), _log10(1 - q_esc_cal_energy), eb, _log10(eb) if eb == eb and eb > 0 else float("nan"),
        )
    return rows


def write_mc_grid(path: str, rows: np.ndarray) -> None:
    """mc_grid.dat: one row per zone (iteration, zone, 33 columns), blank-separated, then the plot separator line."""
    with open(path, "w") as f:
        for r in rows:
            f.write(f"{int(r[0]):4d} {int(r[1]):4d} " + " ".join(f"{v: .10e}" for v in r[2:]) + "\n")
        f.write("\n")


def read_mc_grid(path: str) -> np.ndarray:
    """Inverse of write_mc_grid (the column numbers are the ones `read_old_prof` of the reference relies on)."""
    return np.loadtxt(path, ndmin=2)
