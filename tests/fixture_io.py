"""Reference fixtures: the layout tools/record_reference.jl writes from the REAL Julia reference (manifest.json + raw
little-endian arrays), a writer of the same layout from any engine of this repository (so that the consuming side is
tested without Julia), and the replay runner that checks an engine against a fixture.

Layout (format "mcs-reference-fixture-1"):
  ion_records[]:  i_iter, i_ion, scalars{...names as in main_loops.jl...}, arrays at the START of the ion (profile, pcuts,
                  tcuts, x_spec, inj_fracs, energy_transfer_pool, pxx/pxz/energy_flux right after init_pop) and, prefixed
                  end_, the tallies at the END of the ion as the reference holds them (floors and earlier ions included).
  pcut_records[]: i_iter, i_ion, i_pcut, n_pts_use, new_<field> (population entering the pcut), saved_<field> + saved_l_save,
                  draws / draws_off (uniforms each particle consumed, in order), trace / trace_off (8 x n rows: x, ptot, pb,
                  phi, acctime, prp_x, i_grid, n_draws at the end of each of the first passes of the first particles).
"""
from __future__ import annotations

import json
import os

import numpy as np

from mcs_b200 import abi, driver, problem

POP_FIELDS = ("weight", "ptot_pf", "pb_pf", "x_cm", "xn_per", "prp_x_cm", "acctime_sec", "phi_rad", "grid", "tcut",
              "downstream", "inj")
_DT = {"f8": np.float64, "i8": np.int64, "u1": np.uint8}

# scalar names of the recorder (main_loops.jl identifiers) -> McsConfig fields
CFG_KEYS = {"γ₀": "gam0", "β₀": "beta0", "u₀": "u0", "u₂": "u2", "bmag₂": "bmag2", "pₑ_crit": "pe_crit", "γₑ_crit": "gam_e_crit",
            "η_mfp": "eta_mfp", "psd_mom_min": "psd_mom_min", "psd_cos_fine": "psd_cos_fine", "Δcos": "delta_cos",
            "psd_θ_min": "psd_theta_min", "psd_bins_per_dec_mom": "psd_bins_per_dec_mom",
            "psd_bins_per_dec_θ": "psd_bins_per_dec_theta", "num_psd_mom_bins": "num_psd_mom_bins",
            "num_psd_θ_bins": "num_psd_theta_bins", "energy_transfer_frac": "energy_transfer_frac",
            "feb_upstream": "feb_upstream", "feb_downstream": "feb_downstream", "x_grid_stop": "x_grid_stop", "B_CMBz": "B_CMBz",
            "xn_per_fine": "xn_per_fine", "xn_per_coarse": "xn_per_coarse", "age_max": "age_max", "n_grid": "n_grid",
            "i_grid_feb": "i_grid_feb", "i_shock": "i_shock", "n_ions": "n_ions", "do_rad_losses": "do_rad_losses",
            "do_retro": "do_retro", "do_tcuts": "do_tcuts", "dont_DSA": "dont_DSA", "dont_scatter": "dont_scatter",
            "use_custom_frg": "use_custom_frg", "use_custom_εB": "use_custom_epsB"}
PROFILE_KEYS = {"x_grid_cm": "x_grid_cm", "uₓ_sk_grid": "ux_sk", "uz_sk_grid": "uz_sk", "utot_grid": "utot", "γ_sf_grid": "gam_sf",
                "γ_ef_grid": "gam_ef", "β_ef_grid": "beta_ef", "btot_grid": "btot", "θ_grid": "theta"}


class Fixture:
    def __init__(self, path):
        self.path = path
        self.m = json.load(open(os.path.join(path, "manifest.json"), encoding="utf-8"))
        if self.m.get("format") != "mcs-reference-fixture-1":
            raise ValueError(f"{path}: unknown fixture format")

    def arr(self, rec, key):
        d = rec[key]
        a = np.fromfile(os.path.join(self.path, d["file"]), dtype=_DT[d["dtype"]])
        shape = d["shape"]
        return a.reshape(shape, order="F") if len(shape) > 1 else a   # Julia arrays are column-major


class FixtureWriter:
    def __init__(self, path):
        self.path = path
        os.makedirs(path, exist_ok=True)
        self.m = {"format": "mcs-reference-fixture-1", "pcut_records": [], "ion_records": []}

    def put(self, name, a):
        a = np.asarray(a)
        dt = "u1" if a.dtype == np.uint8 or a.dtype == np.bool_ else ("i8" if a.dtype.kind in "iu" else "f8")
        np.asfortranarray(a.astype(_DT[dt])).ravel(order="F").tofile(os.path.join(self.path, name + ".bin"))
        return {"file": name + ".bin", "dtype": dt, "shape": list(a.shape)}

    def close(self):
        json.dump(self.m, open(os.path.join(self.path, "manifest.json"), "w", encoding="utf-8"), ensure_ascii=False)


def write_fixture_from_engine(path, lib, run, philox, *, i_ion=1, n_trace=8, max_passes=60, seed=210):
    """Run `lib` (Philox mode) through every pcut of one ion and write what the Julia recorder would have written.
    `philox(seed, global_idx, ctr2, ctr3, k)` returns the first k uniforms of a particle's stream."""
    w = FixtureWriter(path)
    inp, prof = run.inp, run.profile
    cfg = driver.make_config(lib, run, seed=seed, na_cr=1_000_000)
    eng = abi.Engine(lib, cfg)
    eps = problem.populate_eps_target(run, prof)
    ip = problem.init_pop(run, prof, i_ion, np.random.default_rng(i_ion - 1))
    sp = driver.species_struct(run, i_ion)
    spc = run.species[i_ion - 1]
    p_hi = problem.pcut_hi(inp.en_pcut_hi, spc.mass)
    inv = {v: k for k, v in CFG_KEYS.items()}
    scal = {inv[f]: getattr(cfg, f) for f in inv}
    scal.update({"n_pts_max": int(cfg.n_pts_max), "n_pts_pcut": inp.n_pts_pcut, "n_pts_pcut_hi": inp.n_pts_pcut_hi,
                 "n_xspec": int(cfg.n_xspec), "n_tcuts": int(cfg.n_tcuts), "electron_weight_fac": sp.electron_weight_fac,
                 "aa": sp.aa, "zz": sp.zz_esu, "m": spc.mass, "pmax_cutoff": sp.pmax_cutoff, "n0": sp.n0,
                 "n_pts_use": len(ip.pop["weight"]), "p_pcut_hi": p_hi})
    rec = {"i_iter": 1, "i_ion": i_ion, "scalars": scal}
    key = f"it1_ion{i_ion}"
    pinv = {v: k for k, v in PROFILE_KEYS.items()}
    for f, jn in pinv.items():
        rec[jn] = w.put(f"{key}_begin_{f}", getattr(prof, f))
    rec["ε_target"] = w.put(f"{key}_begin_eps", eps)
    rec["energy_transfer_pool"] = w.put(f"{key}_begin_pool", np.zeros(run.n_grid))
    rec["pcuts"] = w.put(f"{key}_begin_pcuts", np.asarray(run.pcuts, float))
    rec["tcuts"] = w.put(f"{key}_begin_tcuts", np.asarray(run.tcuts, float))
    rec["x_spec"] = w.put(f"{key}_begin_xspec", np.asarray(run.x_spec_cm, float))
    rec["inj_fracs"] = w.put(f"{key}_begin_injfracs", np.asarray(run.inj_fracs, float))
    for nm, a in (("pxx_flux", ip.pxx_flux), ("pxz_flux", ip.pxz_flux), ("energy_flux", ip.energy_flux)):
        rec[nm] = w.put(f"{key}_begin_{nm}", a + 1.0e-99)
    w.m["ion_records"].append(rec)
    eng.set_profile(prof, eps, np.zeros(run.n_grid))
    eng.begin_ion(1, i_ion, sp, ip.pop)
    for k, pcut in enumerate(run.pcuts, start=1):
        n = eng.population_size()
        pk = f"{key}_pcut{k}"
        pr = {"i_iter": 1, "i_ion": i_ion, "i_pcut": k, "n_pts_use": n}
        new = eng.get_population(0, n)
        for f in POP_FIELDS:
            pr["new_" + f] = w.put(f"{pk}_new_{f}", new[f])
        idx = np.arange(min(n, n_trace))
        eng.trace_enable(idx, max_passes)
        ns, _ = eng.run_pcut(k, float(pcut), float(run.pcuts[k - 2]) if k > 1 else 0.0)
        fates = eng.get_fates(n)
        sv = eng.get_population(1, n)
        for f in POP_FIELDS:
            pr["saved_" + f] = w.put(f"{pk}_saved_{f}", sv[f])
        pr["saved_l_save"] = w.put(f"{pk}_saved_l_save", sv["l_save"])
        off = np.concatenate(([0], np.cumsum(fates["n_draws"]))).astype(np.int64)
        u = np.zeros(off[-1])
        for i in range(n):
            u[off[i]:off[i + 1]] = philox(seed, i, (k & 0xFFFF) | (i_ion << 16), 1, int(fates["n_draws"][i]))
        pr["draws_off"] = w.put(f"{pk}_draws_off", off)
        pr["draws"] = w.put(f"{pk}_draws", u)
        tr = eng.trace_get()
        toff = np.concatenate(([0], np.cumsum([len(t) for t in tr]))).astype(np.int64)
        rows = np.zeros((8, toff[-1]))
        for i, t in enumerate(tr):
            sl = slice(toff[i], toff[i + 1])
            for r, f in enumerate(("x_cm", "ptot_pf", "pb_pf", "phi_rad", "acctime_sec", "prp_x_cm", "i_grid", "n_draws")):
                rows[r, sl] = t[f]
        pr["trace_off"] = w.put(f"{pk}_trace_off", toff)
        pr["trace"] = w.put(f"{pk}_trace", rows)
        w.m["pcut_records"].append(pr)
        if ns == 0:
            break
        eng.split(inp.n_pts_pcut if pcut < p_hi else inp.n_pts_pcut_hi)
    t = eng.end_ion()
    rec["end_scalars"] = {"∑P_downstream": 1e-99 + t.scalars["sum_P_downstream"], "∑KEdensity_downstream": 1e-99 + t.scalars["sum_KE_downstream"]}
    M2, T2, ng = run.num_psd_mom_bins + 2, run.num_psd_theta_bins + 2, run.n_grid
    for nm, a in (("pxx_flux", ip.pxx_flux + t.pxx_flux + 1e-99), ("pxz_flux", ip.pxz_flux + t.pxz_flux + 1e-99),
                  ("energy_flux", ip.energy_flux + t.energy_flux + 1e-99),
                  ("psd", (t.psd + 1e-99).transpose(2, 1, 0)), ("num_crossings", t.num_crossings),
                  ("esc_psd_feb_upstream", (t.esc_psd_feb_upstream + 1e-99).T), ("esc_psd_feb_downstream", (t.esc_psd_feb_downstream + 1e-99).T),
                  ("energy_transfer_pool", t.energy_transfer_pool)):
        rec["end_" + nm] = w.put(f"{key}_end_{nm}", a)
    w.close()
    return path


def _rel(a, b, scale):
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def run_fixture(fx: Fixture, lib, *, tol_state=1e-10, tol_trace=1e-12, tol_phi=1e-8, tol_tally=1e-10, log=print):
    """Replay every ion of the fixture on `lib` (oracle or CUDA).  Raises AssertionError on the first mismatch."""
    m = fx.m
    for ion in m["ion_records"]:
        s = ion["scalars"]
        cfg = abi.default_config(lib)
        for jn, f in CFG_KEYS.items():
            v = s[jn]
            setattr(cfg, f, int(v) if isinstance(getattr(cfg, f), int) else float(v))
        cfg.n_pts_max = int(max(int(s["n_pts_max"]), int(s["n_pts_pcut"]), int(s["n_pts_pcut_hi"])) + 8)
        cfg.na_cr = 4_000_000
        cfg.rng_mode = abi.RNG_REPLAY
        tc, xs, fr = fx.arr(ion, "tcuts"), fx.arr(ion, "x_spec"), fx.arr(ion, "inj_fracs")
        cfg.n_tcuts, cfg.n_xspec = min(int(s["n_tcuts"]), len(tc)), min(int(s["n_xspec"]), len(xs))
        for i in range(cfg.n_tcuts):
            cfg.tcuts[i] = tc[i]
        for i in range(cfg.n_xspec):
            cfg.x_spec[i] = xs[i]
        for i in range(min(len(fr), abi.MAX_IONS)):
            cfg.inj_fracs[i] = fr[i]
        eng = abi.Engine(lib, cfg)

        class Prof:
            pass
        prof = Prof()
        for jn, f in PROFILE_KEYS.items():
            setattr(prof, f, fx.arr(ion, jn))
        eng.set_profile(prof, fx.arr(ion, "ε_target"), fx.arr(ion, "energy_transfer_pool"))  # recv pool = pool so far (main_loops.jl:164)
        sp = abi.McsSpecies(aa=float(s["aa"]), zz_esu=abs(float(s["zz"])), n0=float(s["n0"]), pmax_cutoff=float(s["pmax_cutoff"]),
                            electron_weight_fac=float(s["electron_weight_fac"]) if np.isfinite(float(s["electron_weight_fac"])) else 0.0)
        pcuts = fx.arr(ion, "pcuts")
        recs = [r for r in m["pcut_records"] if r["i_iter"] == ion["i_iter"] and r["i_ion"] == ion["i_ion"]]
        recs.sort(key=lambda r: r["i_pcut"])
        mass = float(s["m"])
        qabs = abs(float(s["zz"]))
        bmag0 = float(prof.btot[0])
        for j, r in enumerate(recs):
            n, k = int(r["n_pts_use"]), int(r["i_pcut"])
            new = {f: fx.arr(r, "new_" + f)[:n] for f in POP_FIELDS}
            if j == 0:
                eng.begin_ion(int(ion["i_iter"]), int(ion["i_ion"]), sp, new)
            else:  # the library's own new_pcut must have produced the reference's next population (cuts.jl:34-98)
                assert eng.population_size() == n, f"pcut {k}: population {eng.population_size()} != reference {n}"
                cur = eng.get_population(0, n)
                for f in POP_FIELDS:
                    if f in ("grid", "tcut", "downstream", "inj", "weight"):
                        assert np.array_equal(cur[f], new[f]), f"pcut {k}: split field {f} differs"
            off = fx.arr(r, "draws_off")
            eng.replay_set_stream(fx.arr(r, "draws") if off[-1] > 0 else np.zeros(1), off)
            toff = fx.arr(r, "trace_off")
            nt = len(toff) - 1
            maxp = int(np.max(np.diff(toff))) if nt > 0 else 0
            if nt > 0 and maxp > 0:
                eng.trace_enable(np.arange(nt), maxp)
            ns, _ = eng.run_pcut(k, float(pcuts[k - 1]), float(pcuts[k - 2]) if k > 1 else 0.0)
            sv = eng.get_population(1, n)
            ls = fx.arr(r, "saved_l_save")[:n]
            assert np.array_equal(sv["l_save"], ls), f"pcut {k}: l_save differs in {(sv['l_save'] != ls).sum()} of {n} particles"
            sel = ls.astype(bool)
            ref = {f: fx.arr(r, "saved_" + f)[:n] for f in POP_FIELDS}
            for f in ("grid", "tcut", "downstream", "inj"):
                assert np.array_equal(sv[f][sel], ref[f][sel]), f"pcut {k}: saved {f} differs"
            ptot = np.abs(ref["ptot_pf"][sel])
            rg = ptot * problem.CL / (qabs * bmag0)
            scales = {"weight": np.abs(ref["weight"][sel]), "ptot_pf": ptot, "pb_pf": ptot, "x_cm": np.maximum(np.abs(ref["x_cm"][sel]), rg),
                      "xn_per": np.abs(ref["xn_per"][sel]), "prp_x_cm": np.maximum(np.abs(ref["prp_x_cm"][sel]), rg),
                      "acctime_sec": np.maximum(np.abs(ref["acctime_sec"][sel]), 1e-300), "phi_rad": np.full(ptot.shape, 2 * np.pi)}
            for f, sc in scales.items():
                e = _rel(sv[f][sel], ref[f][sel], sc)
                assert e <= (tol_phi if f == "phi_rad" else tol_state), f"pcut {k}: saved {f} differs by {e:.3e}"
            if nt > 0 and maxp > 0:
                tr_ref = fx.arr(r, "trace")
                for i, t in enumerate(eng.trace_get()):
                    a = tr_ref[:, toff[i]:toff[i + 1]]
                    assert len(t) == a.shape[1], f"pcut {k} particle {i}: {len(t)} traced passes vs {a.shape[1]}"
                    if len(t) == 0:
                        continue
                    assert np.array_equal(t["i_grid"], a[6].astype(np.int32)), f"pcut {k} particle {i}: zone sequence differs"
                    assert np.array_equal(t["n_draws"], a[7].astype(np.int32)), f"pcut {k} particle {i}: draw counts differ"
                    pt = np.abs(a[1])
                    rgt = pt * problem.CL / (qabs * bmag0)
                    for row, f, sc in ((0, "x_cm", np.maximum(np.abs(a[0]), rgt)), (1, "ptot_pf", pt), (2, "pb_pf", pt),
                                       (3, "phi_rad", np.full(pt.shape, 2 * np.pi)), (5, "prp_x_cm", np.maximum(np.abs(a[5]), rgt))):
                        e = _rel(t[f], a[row], sc)
                        assert e <= (tol_phi if f == "phi_rad" else tol_trace), f"pcut {k} particle {i}: {f} differs by {e:.3e} along the trace"
            log(f"  ion {ion['i_ion']} pcut {k}: {n} particles, {int(sel.sum())} saved, {nt} traced: ok")
            if ns == 0 or j + 1 == len(recs):
                break
            eng.split(int(s["n_pts_pcut"]) if pcuts[k - 1] < float(s["p_pcut_hi"]) else int(s["n_pts_pcut_hi"]))
        t = eng.end_ion()
        if "end_psd" in ion:
            for nm in ("pxx_flux", "pxz_flux", "energy_flux"):
                want = fx.arr(ion, "end_" + nm) - fx.arr(ion, nm)
                got = getattr(t, nm)
                sc = np.abs(want).max() if want.size else 1.0
                assert np.max(np.abs(want - got)) <= tol_tally * 1e2 * max(sc, 1e-300), nm   # difference of two sums: absolute scale
            want = fx.arr(ion, "end_psd")
            got = t.psd.transpose(2, 1, 0)
            big = want > 1e-90
            assert np.array_equal(big, got > 0), "PSD occupancy pattern differs"
            assert _rel(got[big], want[big], np.abs(want[big])) <= tol_tally, "psd"
            assert np.array_equal(fx.arr(ion, "end_num_crossings"), t.num_crossings), "num_crossings"
        log(f"  ion {ion['i_ion']}: tallies ok")
        eng.close()
