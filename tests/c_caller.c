/* A third-party C caller of include/mcs.h: compiled as C (not C++) by tests/test_abi.py and linked against the CPU oracle,
 * which exports the same ABI as libmcs_b200.so.  Proves the header is valid C with C linkage and that the call sequence a
 * binder makes (default config -> create -> profile -> begin ion -> run pcut -> end ion) works from C.
 * Replaces, from the caller's point of view, the loop at /root/reference/src/main_loops.jl:228-292. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcs.h"

#define NG 6
#define NP 64

int main(void) {
    int32_t sizes[6];
    McsConfig cfg;
    McsHandle* h = NULL;
    McsSpecies sp;
    McsPopulation pop;
    McsTallies t;
    double xg[NG + 2] = {-1e30, -3e12, -1e12, 0.0, 1e12, 3e12, 1e13, 1e30};
    double ux[NG + 2], uz[NG + 2], ut[NG + 2], gsf[NG + 2], gef[NG + 2], bef[NG + 2], bt[NG + 2], th[NG + 2];
    double eps[NG], pool[NG], w[NP], p[NP], pb[NP], x[NP], phi[NP], pxx[NG];
    int64_t grid[NP], ncross[NG], n_saved = -1, n_steps = -1;
    int i, rc;

    if (mcs_abi_sizes(sizes) != MCS_OK || sizes[0] != (int32_t)sizeof(McsConfig) || sizes[2] != (int32_t)sizeof(McsTallies)) {
        fprintf(stderr, "struct sizes differ between this translation unit and the library\n");
        return 2;
    }
    mcs_default_config(&cfg);
    cfg.gam0 = 1.0005; cfg.beta0 = 0.0334; cfg.u0 = 1.0e9; cfg.u2 = 2.5e8; cfg.bmag2 = 1.0e-5;
    cfg.psd_mom_min = 1.0e-20; cfg.psd_cos_fine = 0.9; cfg.delta_cos = 0.02; cfg.psd_theta_min = 1.0e-3;
    cfg.psd_bins_per_dec_mom = 10; cfg.psd_bins_per_dec_theta = 10; cfg.num_psd_mom_bins = 40; cfg.num_psd_theta_bins = 40;
    cfg.feb_upstream = -2.5e12; cfg.feb_downstream = -1.0; cfg.x_grid_stop = 1.0e13;
    cfg.n_grid = NG; cfg.i_grid_feb = 1; cfg.i_shock = 3; cfg.n_ions = 1; cfg.n_pts_max = NP; cfg.na_cr = 100000;
    for (i = 0; i < NG + 2; i++) {
        ux[i] = xg[i] < 0.0 ? cfg.u0 : cfg.u2; uz[i] = 0.0; ut[i] = ux[i];
        gsf[i] = 1.0 / sqrt(1.0 - (ux[i] / cfg.c_cms) * (ux[i] / cfg.c_cms)); gef[i] = 1.0; bef[i] = 0.0; bt[i] = 1.0e-5; th[i] = 0.0;
    }
    memset(eps, 0, sizeof eps); memset(pool, 0, sizeof pool);
    rc = mcs_create(&cfg, &h);
    if (rc != MCS_OK) { fprintf(stderr, "mcs_create: %s\n", mcs_last_error()); return 3; }
    rc = mcs_set_profile(h, NG, xg, ux, uz, ut, gsf, gef, bef, bt, th, eps, pool);
    if (rc != MCS_OK) { fprintf(stderr, "mcs_set_profile: %s\n", mcs_last_error()); return 4; }
    sp.aa = 1.0; sp.zz_esu = cfg.qcgs_esu; sp.n0 = 1.0; sp.pmax_cutoff = 1.0e-10; sp.electron_weight_fac = 0.0;
    for (i = 0; i < NP; i++) {
        w[i] = 1.0 / NP; p[i] = 1.0e-17 * (1.0 + i); pb[i] = 0.5 * p[i]; x[i] = -0.5e12; phi[i] = 0.1 * i; grid[i] = 2;
    }
    memset(&pop, 0, sizeof pop);
    pop.weight = w; pop.ptot_pf = p; pop.pb_pf = pb; pop.x_cm = x; pop.phi_rad = phi; pop.grid = grid;
    rc = mcs_begin_ion(h, 1, 1, &sp, NP, 0, &pop);
    if (rc != MCS_OK) { fprintf(stderr, "mcs_begin_ion: %s\n", mcs_last_error()); return 5; }
    rc = mcs_run_pcut(h, 1, 1.0e-15, 0.0, &n_saved, &n_steps);
    if (rc != MCS_OK) { fprintf(stderr, "mcs_run_pcut: %s\n", mcs_last_error()); return 6; }
    memset(&t, 0, sizeof t);
    t.pxx_flux = pxx; t.num_crossings = ncross;
    rc = mcs_end_ion(h, &t);
    if (rc != MCS_OK) { fprintf(stderr, "mcs_end_ion: %s\n", mcs_last_error()); return 7; }
    if (n_steps <= 0 || t.n_helix_steps + t.n_retro_steps != n_steps) { fprintf(stderr, "step counts inconsistent\n"); return 8; }
    {
        int64_t tot = 0;
        for (i = 0; i < 6; i++) tot += t.n_fate[i];
        if (tot != NP) { fprintf(stderr, "fates do not add up: %lld of %d\n", (long long)tot, NP); return 9; }
    }
    printf("c_caller ok: backend %s, %lld steps, %lld saved\n", mcs_backend(), (long long)n_steps, (long long)n_saved);
    mcs_destroy(h);
    return 0;
}
