"""Quick CUDA-vs-oracle comparison on small cases (development aid; the real tests live in tests/)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import mcs_b200
from mcs_b200 import problem, driver, abi
import oracle_engine

def compare(inp, label, n_pcuts=None, seed=210):
    run = problem.setup_run(inp)
    if n_pcuts: run.pcuts = run.pcuts[:n_pcuts]
    olib, clib = oracle_engine.load_oracle_library(), mcs_b200.load_cuda_library()
    engs = []
    for lib in (olib, clib):
        cfg = driver.make_config(lib, run, na_cr=4_000_000, seed=seed)
        engs.append(abi.Engine(lib, cfg))
    prof = run.profile
    eps = problem.populate_eps_target(run, prof)
    ip = problem.init_pop(run, prof, 1, np.random.default_rng(0))
    sp = driver.species_struct(run, 1)
    for e in engs:
        e.set_profile(prof, eps, np.zeros(run.n_grid)); e.begin_ion(1, 1, sp, ip.pop)
    p_hi = problem.pcut_hi(inp.en_pcut_hi, run.species[0].mass)
    ok = True
    for k, pcut in enumerate(run.pcuts, start=1):
        n = engs[0].population_size()
        assert n == engs[1].population_size()
        out = []
        for e in engs:
            t0 = time.time(); ns, nst = e.run_pcut(k, pcut, run.pcuts[k-2] if k > 1 else 0.0); dt = time.time() - t0
            out.append((ns, nst, dt, e.get_fates(n), e.get_population(1, n)))
        (ns0, st0, dt0, f0, s0), (ns1, st1, dt1, f1, s1) = out
        bad_f = int((f0['fate'] != f1['fate']).sum()); bad_h = int((f0['helix_count'] != f1['helix_count']).sum())
        bad_d = int((f0['n_draws'] != f1['n_draws']).sum()); bad_r = int((f0['retro_steps'] != f1['retro_steps']).sum())
        mx = 0.0
        m = s0['l_save'].astype(bool) & s1['l_save'].astype(bool)
        for nm in abi.POP_F64:
            a, b = s0[nm][m], s1[nm][m]
            scale = np.maximum(np.abs(a), 1e-300)
            if nm == 'pb_pf': scale = np.abs(s0['ptot_pf'][m])
            if nm == 'phi_rad': scale = np.full_like(a, 2*np.pi)
            if a.size: mx = max(mx, float(np.max(np.abs(a-b)/scale)))
        bad_i = sum(int((s0[nm][m] != s1[nm][m]).sum()) for nm in abi.POP_I64 + abi.POP_U8)
        print(f"[{label}] pcut {k}: n={n} saved {ns0}/{ns1} steps {st0}/{st1} oracle {dt0:.2f}s gpu {dt1:.3f}s | mismatches fate={bad_f} helix={bad_h} draws={bad_d} retro={bad_r} ints={bad_i} max_rel={mx:.2e}")
        ok &= (ns0 == ns1 and bad_f == 0 and bad_h == 0 and bad_i == 0)
        if ns0 == 0: break
        target = inp.n_pts_pcut if pcut < p_hi else inp.n_pts_pcut_hi
        for e in engs: e.split(target)
    t0, t1 = engs[0].end_ion(), engs[1].end_ion()
    def rel(a, b):
        d = np.abs(a-b); s = np.maximum(np.abs(a), np.abs(b)); s[s == 0] = 1
        return float((d/s).max()) if a.size else 0.0
    for nm in ("pxx_flux","pxz_flux","energy_flux","psd","esc_psd_feb_upstream","esc_psd_feb_downstream","esc_energy_eff","esc_num_eff","weight_coupled","spectra_coupled","energy_transfer_pool"):
        ok &= rel(getattr(t0,nm), getattr(t1,nm)) < 1e-8
        print(f"   {nm:26s} rel diff {rel(getattr(t0,nm), getattr(t1,nm)):.2e}  sum {getattr(t0,nm).sum():.6e} / {getattr(t1,nm).sum():.6e}")
    print("   num_crossings equal:", bool((t0.num_crossings == t1.num_crossings).all()), " log", len(t0.therm_grid), len(t1.therm_grid))
    k0 = np.lexsort((t0.therm_weight, t0.therm_ptot_sk, t0.therm_px_sk, t0.therm_grid)); k1 = np.lexsort((t1.therm_weight, t1.therm_ptot_sk, t1.therm_px_sk, t1.therm_grid))
    if len(k0) == len(k1) and len(k0):
        print("   log sorted: grid equal", bool((t0.therm_grid[k0] == t1.therm_grid[k1]).all()), "px rel", rel(t0.therm_px_sk[k0], t1.therm_px_sk[k1]), "w rel", rel(t0.therm_weight[k0], t1.therm_weight[k1]))
    for k in t0.scalars: ok &= abs(t0.scalars[k]-t1.scalars[k]) <= 1e-9*abs(t0.scalars[k])
    ok &= t0.stats == t1.stats and bool((t0.num_crossings == t1.num_crossings).all())
    print("   scalars", {k: (t0.scalars[k], t1.scalars[k]) for k in t0.scalars})
    print("   stats", t0.stats, "\n        ", t1.stats)
    print("   timing", engs[1].timing())
    return ok

if __name__ == "__main__":
    print("backend", mcs_b200.load_cuda_library().mcs_backend())
    ok = compare(problem.bundled_input(), "bundled", n_pcuts=2)
    ok &= compare(problem.planar_test_particle_input(3000, momentum_cutoffs=[0.01,0.04,0.06,0.09,0.13,0.2,0.3]), "planar")
    ok &= compare(problem.relativistic_input(1000), "rel10", n_pcuts=8)
    print("ALL OK" if ok else "MISMATCH")
