"""Statistical parity of full runs (north_star: "production runs must match the reference's PSD spectra, escape fluxes
and returned-flux profiles within statistical error, per-bin chi-square and KS at a stated p").

One high-statistics run of the engine under test against R independent low-statistics replicas of the checker (different
seeds and injection draws).  Per bin: z = (mean_R - value) / (sd_R / sqrt(R)) follows Student-t with R-1 degrees of freedom
when both sample the same distribution and the high-statistics run's own noise is negligible (it has >= 30x the particles of
all replicas together; its variance is added from the replica variance scaled by the particle ratio anyway).  Reported:
chi-square of the z scores (t^2 rescaled by (R-3)/(R-1) to unit mean) with its p-value, and a KS test of the z scores
against t_{R-1}.  Used by tests/test_stat_parity_gpu.py and by bench.py's cpu_baseline leg."""
import numpy as np
from scipy import stats


def observables(t, run):
    """The spectra and profiles compared: angle- and zone-summed dN/dp, momentum and energy flux profiles, downstream
    escape spectrum, upstream escape spectrum."""
    return {
        "dNdp": t.psd.sum(axis=(0, 1)),
        "pxx_flux": np.asarray(t.pxx_flux, float),
        "energy_flux": np.asarray(t.energy_flux, float),
        "esc_spectrum_downstream": t.esc_psd_feb_downstream.sum(axis=0),
        "esc_spectrum_upstream": t.esc_psd_feb_upstream.sum(axis=0),
    }


def compare(big: dict, replicas: list, n_ratio: float, min_bins: int = 3) -> dict:
    """big: observables of the high-statistics run; replicas: list of observables; n_ratio = particles(big) / particles(one
    replica).  Returns {name: {bins, chi2_per_bin, p_chi2, p_ks, max_abs_z}}."""
    R = len(replicas)
    out = {}
    for key in big:
        G = np.array([r[key] for r in replicas], float)
        mean, sd = G.mean(axis=0), G.std(axis=0, ddof=1)
        ok = (sd > 0) & (np.count_nonzero(G, axis=0) == R) & (big[key] != 0)   # populated in every replica
        if ok.sum() < min_bins:
            out[key] = {"bins": int(ok.sum()), "skipped": "too few populated bins"}
            continue
        var = sd[ok] ** 2 * (1.0 / R + 1.0 / n_ratio)
        z = (mean[ok] - big[key][ok]) / np.sqrt(var)
        chi2 = float((z**2).sum() * (R - 3) / (R - 1))
        out[key] = {"bins": int(ok.sum()), "chi2_per_bin": chi2 / int(ok.sum()), "p_chi2": float(stats.chi2.sf(chi2, int(ok.sum()))),
                    "p_ks": float(stats.kstest(z, stats.t(df=R - 1).cdf).pvalue), "max_abs_z": float(np.abs(z).max())}
    return out
