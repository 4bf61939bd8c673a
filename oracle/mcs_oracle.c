/*
 * mcs_oracle.c — CPU ORACLE for the per-particle transport loop.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path may import, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the reference (abhro/MonteCarloScattering.jl) ships no golden vectors or
 * functional tests (test/runtests.jl:1-4 is Aqua only) and no Julia runtime exists in this image,
 * so this restatement cannot be checked against reference outputs.  It is pinned instead by the
 * known answers the source itself documents (tests/test_oracle_known_answers.py).
 *
 * What it is: a scalar FP64, serial, line-by-line restatement in plain C of
 *   src/particle_loop.jl (all), src/scattering.jl (all), src/transformers.jl:440-607,
 *   src/all_flux.jl (all), src/get_psd_bins.jl (all), src/prob_return.jl:36-344,
 *   src/cuts.jl (all), src/particle_finish.jl (all)
 * written from the Julia text (raw cgs doubles instead of Unitful quantities), exposing the same
 * C-ABI as the CUDA library (include/mcs.h) so the same host driver can run either.
 * Compile with -O2 -ffp-contract=off: Julia never contracts a*b+c.
 *
 * Where the reference as written throws or hangs (class F) the evidently intended
 * arithmetic is used and the site is marked "F-n"; class K quirks are kept and marked "K-n".
 */
#include "../include/mcs.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.141592653589793
#define TWO_PI 6.283185307179586
/* 2*pi split for the extended-precision reduction in mod2pi (Julia Base.mod2pi keeps a hi/lo pair) */
#define TWO_PI_LO 2.4492935982947064e-16
#define SIN_UPPER_LIMIT 0.99999999999999989 /* prevfloat(1.0), scattering.jl:3 */
#define ALL_FLUX_SPIKE_AWAY 1000.0          /* all_flux.jl:4 */
#define PF_SPIKE_AWAY 1000.0                /* particle_finish.jl:5 */

static __thread char g_err[512];
static int fail(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
const char* mcs_last_error(void) { return g_err; }
const char* mcs_backend(void) { return "cpu-oracle"; }

/* ------------------------------------------------------------------------------------------ */
/* population storage                                                                          */
typedef struct {
    double *weight, *ptot, *pb, *x, *xn_per, *prp_x, *acctime, *phi;
    int64_t *grid, *tcut;
    uint8_t *down, *inj;
} Pop;

static int pop_alloc(Pop* p, int64_t n) {
    size_t nd = (size_t)(n > 0 ? n : 1);
    p->weight = calloc(nd, 8); p->ptot = calloc(nd, 8); p->pb = calloc(nd, 8); p->x = calloc(nd, 8);
    p->xn_per = calloc(nd, 8); p->prp_x = calloc(nd, 8); p->acctime = calloc(nd, 8); p->phi = calloc(nd, 8);
    p->grid = calloc(nd, 8); p->tcut = calloc(nd, 8); p->down = calloc(nd, 1); p->inj = calloc(nd, 1);
    return (p->weight && p->ptot && p->pb && p->x && p->xn_per && p->prp_x && p->acctime && p->phi &&
            p->grid && p->tcut && p->down && p->inj) ? 0 : -1;
}
static void pop_free(Pop* p) {
    free(p->weight); free(p->ptot); free(p->pb); free(p->x); free(p->xn_per); free(p->prp_x);
    free(p->acctime); free(p->phi); free(p->grid); free(p->tcut); free(p->down); free(p->inj);
    memset(p, 0, sizeof *p);
}
static void pop_zero(Pop* p, int64_t n) {
    size_t nd = (size_t)n;
    memset(p->weight, 0, nd * 8); memset(p->ptot, 0, nd * 8); memset(p->pb, 0, nd * 8); memset(p->x, 0, nd * 8);
    memset(p->xn_per, 0, nd * 8); memset(p->prp_x, 0, nd * 8); memset(p->acctime, 0, nd * 8);
    memset(p->phi, 0, nd * 8); memset(p->grid, 0, nd * 8); memset(p->tcut, 0, nd * 8);
    memset(p->down, 0, nd); memset(p->inj, 0, nd);
}

struct McsHandle {
    McsConfig cfg;
    McsSpecies sp;
    int have_profile, have_ion, ended;
    int32_t i_iter, i_ion, i_pcut;
    double pcut, pcut_prev;
    int n_grid, M, T; /* zones, num_psd_mom_bins, num_psd_theta_bins */
    /* grid arrays, n_grid+2 nodes, index == Julia offset index */
    double *xg, *ux, *uz, *ut, *gsf, *gef, *bef, *bt, *th;
    double *eps_target, *recv_pool; /* [n_grid], Julia index i -> [i-1] */
    /* populations */
    Pop cur, saved, next;
    uint8_t* l_save;
    int64_t n_use, first_global, n_saved_last;
    /* per-particle outcomes of the last pcut */
    int32_t *fate, *helix;
    int64_t *retro, *draws;
    /* tallies (per ion) */
    double *pxx, *pxz, *efl, *psd;
    int64_t* ncross;
    int64_t n_cr, n_cr_over;
    int64_t* tg;
    double *tpx, *tpt, *tw;
    double *esc_up, *esc_dn, *esc_en_eff, *esc_num_eff, *w_coupled, *s_coupled, *pool, *spec_sf, *spec_pf;
    double *th_sf, *th_pf; /* thermal crossings binned on the fly, [jth + (T+2)*(k + (M+2)*(i-1))] */
    double esc_flux, px_esc_feb, en_esc_feb, sumP, sumKE, px_esc_up, en_esc_up;
    int64_t n_helix, n_retro, w_pperp, w_psdmom, n_negsqrt, n_retro_cap, n_err, n_fate[6];
    /* replay + trace */
    double* replay_u;
    int64_t* replay_off;
    int64_t replay_n;
    int64_t* trace_idx;
    int32_t n_trace, trace_max;
    McsTraceRec* trace_recs;
    int32_t* trace_cnt;
    int par; /* >1 threads: tallies use omp atomic */
};

/* ------------------------------------------------------------------------------------------ */
/* RNG: Philox4x32-10 counter stream, or a recorded uniform stream (replay).                   */
/* Stands in for Random.Xoshiro(iseed_mod) of particle_loop.jl:34-41 (one private stream per    */
/* (iter, ion, pcut, particle)); Random.rand is a 53-bit uniform in [0,1).                      */
typedef struct {
    int mode;
    uint32_t key[2], ctr[4], out[4];
    int64_t n;
    const double* ru;
    int64_t rn;
    int exhausted;
} Rng;

static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u53(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53; }

static double rng_uniform(Rng* g) {
    if (g->mode == MCS_RNG_REPLAY) {
        if (g->n >= g->rn) { g->exhausted = 1; g->n++; return 0.5; }
        return g->ru[g->n++];
    }
    int64_t k = g->n++;
    if ((k & 1) == 0) {
        g->ctr[0] = (uint32_t)(k >> 1);
        philox4x32_10(g->ctr, g->key, g->out);
        return (double)((((uint64_t)g->out[1] << 32) | g->out[0]) >> 11) * 0x1.0p-53;
    }
    return (double)((((uint64_t)g->out[3] << 32) | g->out[2]) >> 11) * 0x1.0p-53;
}

/* ------------------------------------------------------------------------------------------ */
/* small numerics that mirror Julia Base                                        */
static double mod2pi(double x) {
    /* Base.mod2pi: identity on [0,2pi); otherwise reduce with an extended-precision 2pi */
    if (x >= 0.0 && x < TWO_PI) return x;
    double k = floor(x / TWO_PI);
    double r = fma(-k, TWO_PI, x);
    r = r - k * TWO_PI_LO;
    if (r < 0.0) r += TWO_PI;
    if (r >= TWO_PI) r -= TWO_PI;
    return r;
}
static double norm3(double x, double y, double z) {
    /* LinearAlgebra.norm(::SVector{3}) : sqrt(sum(abs2)) with a scaled fallback at 0/Inf */
    double s = x * x + y * y + z * z;
    if (s == 0.0 || isinf(s)) {
        double m = fmax(fabs(x), fmax(fabs(y), fabs(z)));
        if (m == 0.0 || isinf(m)) return m;
        double a = x / m, b = y / m, c = z / m;
        return m * sqrt(a * a + b * b + c * c);
    }
    return sqrt(s);
}
#define TADD(h, lv, v)                                          \
    do {                                                        \
        if ((h)->par) {                                         \
            _Pragma("omp atomic") lv += (v);                    \
        } else                                                  \
            lv += (v);                                          \
    } while (0)

static double sqrt_guard(struct McsHandle* h, double a) {
    /* Julia's sqrt throws DomainError for a<0 (App. E); clamp at 0 and count */
    if (a < 0.0) { TADD(h, h->n_negsqrt, 1); return 0.0; }
    return sqrt(a);
}

/* ------------------------------------------------------------------------------------------ */
/* get_psd_bins.jl:16-39 */
static int get_psd_bin_momentum(struct McsHandle* h, double ptot_sk) {
    const McsConfig* c = &h->cfg;
    int bin;
    if (ptot_sk < c->psd_mom_min) bin = 0;
    else bin = (int)trunc(log10(ptot_sk / c->psd_mom_min) * c->psd_bins_per_dec_mom) + 1;
    if (bin > c->num_psd_mom_bins) { TADD(h, h->w_psdmom, 1); bin = c->num_psd_mom_bins; }
    return bin;
}
/* get_psd_bins.jl:73-97 */
static int get_psd_bin_angle(struct McsHandle* h, double px_sk, double ptot_sk) {
    const McsConfig* c = &h->cfg;
    if (ptot_sk == 0.0) return 0;
    double p_cos = -px_sk / ptot_sk;
    int bin;
    if (p_cos < c->psd_cos_fine) {
        bin = c->num_psd_theta_bins - (int)trunc((p_cos + 1) / c->delta_cos);
    } else {
        double th = acos(p_cos);
        bin = th < c->psd_theta_min ? 0 : (int)trunc(log10(th / c->psd_theta_min) * c->psd_bins_per_dec_theta) + 1;
    }
    if (bin > c->num_psd_theta_bins) bin = c->num_psd_theta_bins;
    return bin;
}

/* transformers.jl:440-476 */
static void transform_p_PS(struct McsHandle* h, double aa, double pb_pf, double p_perp, double gam_pf, double phi,
                           double ux, double uz, double utot, double gam_sf, double bcos, double bsin,
                           double* ptot_sk, double psk[3], double* gam_sk) {
    (void)uz; (void)utot;
    double m = aa * h->cfg.mp_g, mc = m * h->cfg.c_cms;
    double phi_p = phi + PI / 2;
    double p_p_cos = p_perp * cos(phi_p);
    double pfx = pb_pf * bcos - p_p_cos * bsin;
    double pfy = p_perp * sin(phi_p);
    double pfz = pb_pf * bsin + p_p_cos * bcos;
    double dpx = (gam_sf - 1) * pfx + gam_sf * gam_pf * m * ux;
    psk[0] = pfx + dpx; psk[1] = pfy; psk[2] = pfz;
    *ptot_sk = norm3(psk[0], psk[1], psk[2]);
    *gam_sk = hypot(*ptot_sk / mc, 1);
}

/* transformers.jl:523-607 */
static void transform_p_PSP(struct McsHandle* h, double aa, double* pb_pf, double* p_perp, double* gam_pf,
                            double* phi, double ux_o, double uz_o, double ut_o, double gsf_o, double bcos_o,
                            double bsin_o, double ux, double uz, double ut, double gsf, double bcos, double bsin,
                            double* ptot_pf) {
    double m = aa * h->cfg.mp_g, mc = m * h->cfg.c_cms;
    double phi_p = *phi + PI / 2;
    double p_p_cos = *p_perp * cos(phi_p);
    double fx = *pb_pf * bcos_o - p_p_cos * bsin_o;
    double fy = *p_perp * sin(phi_p);
    double fz = *pb_pf * bsin_o + p_p_cos * bcos_o;
    /* old plasma -> shock */
    double rxo = ux_o / ut_o, rzo = uz_o / ut_o;
    double sx = ((gsf_o - 1) * (rxo * rxo) + 1) * fx + (gsf_o - 1) * (ux_o * uz_o / (ut_o * ut_o)) * fz +
                gsf_o * *gam_pf * m * ux_o;
    double sy = fy;
    double sz = (gsf_o - 1) * (ux_o * uz_o / (ut_o * ut_o)) * fx + ((gsf_o - 1) * (rzo * rzo) + 1) * fz +
                gsf_o * *gam_pf * m * uz_o;
    double ptot_sk = norm3(sx, sy, sz);
    double pb_sk = sx * bcos + sz * bsin;
    if (ptot_sk < fabs(pb_sk)) TADD(h, h->w_pperp, 1); /* values unused downstream (:565) */
    double gam_sk = hypot(ptot_sk / mc, 1);
    /* shock -> new plasma */
    double rx = ux / ut, rz = uz / ut;
    double nx = ((gsf - 1) * (rx * rx) + 1) * sx + (gsf - 1) * (ux * uz / (ut * ut)) * sz - gsf * gam_sk * m * ux;
    double ny = sy;
    double nz = (gsf - 1) * (ux * uz / (ut * ut)) * sx + ((gsf - 1) * (rz * rz) + 1) * sz - gsf * gam_sk * m * uz;
    double pt = norm3(nx, ny, nz);
    double pb = nx * bcos + nz * bsin;
    double pp;
    if (pt < fabs(pb)) {
        pp = 1.0e-6 * pt;
        pb = copysign(sqrt(pt * pt - pp * pp), pb);
        TADD(h, h->w_pperp, 1);
    } else {
        pp = sqrt(pt * pt - pb * pb);
    }
    *ptot_pf = pt; *pb_pf = pb; *p_perp = pp;
    *gam_pf = hypot(pt / mc, 1);
    *phi = atan2(ny, -nx * bsin + nz * bcos) - PI / 2;
}

/* particle_loop.jl:639-650 (only p_perp is returned; the adjusted pb is dropped as in the source) */
static double perpendicular_momentum(struct McsHandle* h, double ptot, double pb) {
    if (ptot < fabs(pb)) { TADD(h, h->w_pperp, 1); return 1.0e-6 * ptot; }
    return sqrt(ptot * ptot - pb * pb);
}

/* particle_loop.jl:578-592 */
static double radiation_loss(struct McsHandle* h, double B2, double p, double dt) {
    double d = h->cfg.rad_loss_fac * B2 * p * dt;
    if (d > 1.0e-2) p /= 1 + d; else p *= 1 - d;
    return p;
}

/* scattering.jl:29-101 */
static void scattering(struct McsHandle* h, Rng* rng, double aa, double gyro_denom, double ptot_pf, double gam_pf,
                       double xn_per, double* gyro_period, double* pb_pf, double* p_perp, double* phi) {
    const McsConfig* c = &h->cfg;
    double mc = aa * c->mp_g * c->c_cms;
    double gyro_rad_tot;
    if (aa < 1 && ptot_pf < c->pe_crit) {
        gyro_rad_tot = c->pe_crit * c->c_cms * gyro_denom;
        *gyro_period = TWO_PI * c->gam_e_crit * mc * gyro_denom;
    } else {
        gyro_rad_tot = ptot_pf * c->c_cms * gyro_denom;
        *gyro_period = TWO_PI * gam_pf * mc * gyro_denom;
    }
    double vp_tg = TWO_PI * gyro_rad_tot;
    double lambda = c->eta_mfp * gyro_rad_tot; /* use_custom_frg is an error() in the reference (:53) */
    double cos_max = cos(sqrt(6 * vp_tg / (xn_per * lambda)));
    double cos_old = *pb_pf / ptot_pf, sin_old = *p_perp / ptot_pf;
    double cos_d = 1 - rng_uniform(rng) * (1 - cos_max);
    double sin_d = sqrt_guard(h, 1 - cos_d * cos_d);
    double phi_s = rng_uniform(rng) * TWO_PI - PI;
    double cos_new = cos_old * cos_d + sin_old * sin_d * cos(phi_s);
    double sin_new = sqrt_guard(h, 1 - cos_new * cos_new);
    *pb_pf = ptot_pf * cos_new;
    *p_perp = ptot_pf * sin_new;
    double phi_p_new = *phi + PI / 2;
    if (sin_new != 0) {
        double s = sin(phi_s) * sin_d / sin_new; /* get_sine_adjustment :93-101 */
        if (fabs(s) > SIN_UPPER_LIMIT) s = copysign(SIN_UPPER_LIMIT, s);
        phi_p_new += asin(s);
    }
    *phi = phi_p_new - PI / 2;
}

/* cuts.jl:149-162 */
static void tcut_track(struct McsHandle* h, int tcut_curr, double weight, double ptot_pf) {
    TADD(h, h->w_coupled[tcut_curr - 1], weight);
    int ip = get_psd_bin_momentum(h, ptot_pf);
    TADD(h, h->s_coupled[ip + (MCS_PSD_MAX + 1) * (tcut_curr - 1)], weight);
}

/* all_flux.jl:45-259: returns 0, or -1 where the reference calls error() (:73-75) */
static int all_flux(struct McsHandle* h, double aa, double pb_pf, double p_perp, double ptot_pf, double gam_pf,
                    double phi, double weight, int* i_grid_io, int* i_grid_old_out, double ux, double uz,
                    double utot, double gam_sf, double bcos, double bsin, double x, double x_old, int inj) {
    const McsConfig* c = &h->cfg;
    int ng = h->n_grid, i_grid = *i_grid_io, i_grid_old = i_grid;
    /* findnext / findprev are linear scans from the current zone (K-3) */
    if (x > x_old) {
        int k = i_grid + 1;
        while (k <= ng + 1 && !(h->xg[k] > x)) k++;
        if (k > ng + 1) return -1;
        i_grid = k - 1;
    } else {
        int k = i_grid;
        while (k >= 0 && !(h->xg[k] <= x)) k--;
        if (k < 0) return -1;
        i_grid = k;
    }
    *i_grid_io = i_grid; *i_grid_old_out = i_grid_old;
    if (i_grid == i_grid_old && i_grid > c->i_grid_feb && c->n_xspec == 0) return 0;

    double ptot_sk, psk[3], gam_sk;
    transform_p_PS(h, aa, pb_pf, p_perp, gam_pf, phi, ux, uz, utot, gam_sf, bcos, bsin, &ptot_sk, psk, &gam_sk);
    double m = aa * c->mp_g;
    double pt_o_px_sk, abs_inv_vx;
    if (ptot_sk > fabs(psk[0] * ALL_FLUX_SPIKE_AWAY)) {
        pt_o_px_sk = ALL_FLUX_SPIKE_AWAY;
        abs_inv_vx = fabs(ALL_FLUX_SPIKE_AWAY / ux);
    } else {
        pt_o_px_sk = ptot_sk / psk[0];
        abs_inv_vx = fabs(gam_sk * aa * c->mp_g / psk[0]);
    }
    double pt_o_px_pf = fmin(fabs(ptot_pf / pb_pf), ALL_FLUX_SPIKE_AWAY);
    double en_add;
    if ((gam_sk - 1) > c->E_rel_pt) en_add = (gam_sk - 1) * m * (c->c_cms * c->c_cms) * weight;
    else en_add = ptot_sk * ptot_sk / (2 * m) * weight;

    if (c->n_xspec > 0) { /* calculate_x_spec_spectra! :164-190 */
        int ipt = get_psd_bin_momentum(h, ptot_sk), ipf = get_psd_bin_momentum(h, ptot_pf);
        for (int i = 0; i < c->n_xspec; i++) {
            double xs = c->x_spec[i];
            if ((x_old < xs && x >= xs) || (x <= xs && x_old > xs)) {
                TADD(h, h->spec_sf[ipt + (MCS_PSD_MAX + 1) * i], weight * pt_o_px_sk);
                double F = fabs(pb_pf / psk[0]) * (gam_sk / gam_pf);
                TADD(h, h->spec_pf[ipf + (MCS_PSD_MAX + 1) * i], weight * pt_o_px_pf * F);
            }
        }
    }

    /* F_stream! :197-259 */
    int lo, hi, step, inj_check;
    double sign_fac;
    if (x > x_old) { lo = i_grid_old + 1; hi = i_grid; step = 1; inj_check = 0; sign_fac = 1; }
    else { lo = i_grid_old; hi = i_grid + 1; step = -1; inj_check = 1; sign_fac = -1; }
    int ipt = 0, jth = 0;
    if (inj) { ipt = get_psd_bin_momentum(h, ptot_sk); jth = get_psd_bin_angle(h, psk[0], ptot_sk); }
    double g0u0 = c->gam0 * c->u0;
    for (int i = lo; step > 0 ? i <= hi : i >= hi; i += step) {
        if (inj_check && inj && i <= c->i_grid_feb) continue;
        TADD(h, h->pxx[i - 1], sign_fac * psk[0] * weight * c->gam0 * c->u0);
        TADD(h, h->pxz[i - 1], fabs(psk[2]) * weight * c->gam0 * c->u0);
        TADD(h, h->efl[i - 1], sign_fac * en_add * c->gam0 * c->u0);
        if (inj) {
            size_t idx = (size_t)ipt + (size_t)(h->M + 2) * ((size_t)jth + (size_t)(h->T + 2) * (size_t)(i - 1));
            TADD(h, h->psd[idx], weight * abs_inv_vx);
        } else {
            int64_t slot = -1;
            if (h->par) {
#pragma omp critical(mcs_log)
                { if (h->n_cr < c->na_cr) slot = h->n_cr++; else h->n_cr_over++; }
            } else {
                if (h->n_cr < c->na_cr) slot = h->n_cr++; else h->n_cr_over++;
            }
            if (slot >= 0) {
                h->tg[slot] = i; h->tpx[slot] = psk[0]; h->tpt[slot] = ptot_sk; h->tw[slot] = weight * abs_inv_vx;
            }
            if (c->bin_thermal) {
                /* particle_counter.jl:426-445: shock-frame bins of the crossing */
                int k = get_psd_bin_momentum(h, ptot_sk), jt = get_psd_bin_angle(h, psk[0], ptot_sk);
                TADD(h, h->th_sf[(size_t)jt + (size_t)(h->T + 2) * ((size_t)k + (size_t)(h->M + 2) * (size_t)(i - 1))], weight * abs_inv_vx);
                /* thermo_calcs.jl:133-164: boost to the plasma frame of zone i, then bin */
                double E0 = m * (c->c_cms * c->c_cms), g = h->gsf[i], b = h->ux[i] / c->c_cms;
                double etot = hypot(ptot_sk * c->c_cms, E0);
                double pxX = g * (psk[0] - b * etot / c->c_cms);
                double ptX = sqrt((ptot_sk * ptot_sk - psk[0] * psk[0]) + pxX * pxX);
                if (fabs(pxX) > ptX) pxX = copysign(ptX, pxX);
                int kX = get_psd_bin_momentum(h, ptX), jX = get_psd_bin_angle(h, pxX, ptX);
                TADD(h, h->th_pf[(size_t)jX + (size_t)(h->T + 2) * ((size_t)kX + (size_t)(h->M + 2) * (size_t)(i - 1))], weight * abs_inv_vx);
            }
            TADD(h, h->ncross[i - 1], 1);
        }
    }
    (void)g0u0;
    /* upstream FEB escape scalars :155-158 (F-8: units ignored; K-9: tallied although never returned) */
    if (inj && x < c->feb_upstream && x_old >= c->feb_upstream) {
        TADD(h, h->en_esc_up, en_add * c->gam0 * c->u0);
        TADD(h, h->px_esc_up, -(psk[0] * weight * c->gam0 * c->u0));
    }
    return 0;
}

/* particle_loop.jl:652-723; outputs that the caller never uses are not returned */
static void do_energy_transfer(struct McsHandle* h, int i_grid, int i_grid_old, double* ptot_pf, double* pb_pf,
                               double* p_perp, double* gam_pf, double weight, double mc, double aa) {
    const McsConfig* c = &h->cfg;
    int i_start = i_grid_old, i_stop = i_grid < c->i_shock ? i_grid : c->i_shock;
    int scale = 0;
    double m = aa * c->mp_g, E0 = m * (c->c_cms * c->c_cms), gam_f = 0.0;
    /* F-11: an empty range (particle moved upstream) makes maximum() throw; treated as "nothing to do" */
    double emax = -INFINITY, rmax = 0.0;
    for (int i = i_start + 1; i <= i_stop; i++) {
        if (h->eps_target[i - 1] > emax) emax = h->eps_target[i - 1];
        if (h->recv_pool[i - 1] > rmax) rmax = h->recv_pool[i - 1];
    }
    if (aa >= 1 && i_start + 1 <= i_stop && emax > 0) {
        double gam_i = hypot(1, *ptot_pf / mc);
        /* eps_target[i_start] with i_start==0 is a BoundsError in the reference; zone 0 has eps = 0 */
        double eps_start = i_start >= 1 ? h->eps_target[i_start - 1] : 0.0;
        gam_f = 1 + (gam_i - 1) * (1 - h->eps_target[i_stop - 1]) / (1 - eps_start);
        int n_split = 0;
        for (int i = i_start + 1; i <= i_stop; i++) n_split += h->eps_target[i - 1] > 0;
        double inc = (gam_i - gam_f) * E0 * weight / n_split;
        for (int i = i_start + 1; i <= i_stop; i++)
            if (h->eps_target[i - 1] > 0) TADD(h, h->pool[i - 1], inc);
        scale = 1;
    } else if (rmax > 0) {
        double sum = 0.0;
        for (int i = i_start + 1; i <= i_stop; i++) sum += h->recv_pool[i - 1];
        double e = sum * h->sp.electron_weight_fac;
        double gam_i = hypot(1, *ptot_pf / mc);
        gam_f = gam_i + e / E0;
        scale = 1;
    }
    if (scale) {
        double pf = mc * sqrt_guard(h, gam_f * gam_f - 1);
        double s = pf / *ptot_pf;
        *pb_pf *= s; *p_perp *= s; *ptot_pf = pf; *gam_pf = gam_f;
    }
}

/* prob_return.jl:217-344 */
static void retro_time(struct McsHandle* h, Rng* rng, double aa, double zz, double* gyro_denom, double prp_x,
                       double* ptot_pf, double* pb_pf, double* p_perp, double* gam_pf, double* acctime,
                       double weight, int* tcut_curr, double mc, int* lose_pt, double* phi_out,
                       int64_t* n_steps) {
    const McsConfig* c = &h->cfg;
    int ng = h->n_grid;
    double xn_per = 10.0, phi_step = TWO_PI / xn_per;
    double t_step_fac = TWO_PI * aa * c->mp_g * c->c_cms * *gyro_denom / xn_per;
    double ux = -h->ux[ng], gsf = h->gsf[ng], gef = h->gef[ng], B = h->bt[ng];
    if (c->use_custom_epsB) B *= sqrt(c->x_grid_stop / prp_x);
    double bcos = cos(h->th[ng]), bsin = sin(h->th[ng]);
    double Bcmb = c->B_CMBz * gef, B2 = B * B + Bcmb * Bcmb;
    *lose_pt = 0;
    double x = prp_x;
    double phi = rng_uniform(rng) * TWO_PI;
    int64_t steps = 0;
    for (;;) {
        steps++;
        double x_old = x, phi_old = phi, ptot_old = *ptot_pf;
        double cos_old = *pb_pf / *ptot_pf, sin_old = *p_perp / *ptot_pf;
        if (c->use_custom_epsB) {
            B = h->bt[ng] * sqrt(c->x_grid_stop / x);
            B2 = B * B + Bcmb * Bcmb;
            *gyro_denom = 1 / (zz * B);
        }
        double gyro_rad = *p_perp * c->c_cms * *gyro_denom;
        phi = mod2pi(phi_old + phi_step);
        double t_step = t_step_fac * *gam_pf;
        double x_move = *pb_pf * t_step_fac / (aa * c->mp_g);
        x = x_old + gsf * (x_move * bcos - gyro_rad * bsin * (cos(phi) - cos(phi_old)) + ux * t_step);
        *acctime += t_step * gef;
        if (c->do_tcuts && *tcut_curr <= c->n_tcuts && *acctime >= c->tcuts[*tcut_curr - 1]) {
            tcut_track(h, *tcut_curr, weight, *ptot_pf);
            *tcut_curr += 1;
        }
        /* large-angle scattering */
        phi = TWO_PI * rng_uniform(rng);
        *pb_pf = (2 * rng_uniform(rng) - 1) * *ptot_pf;
        *p_perp = sqrt_guard(h, *ptot_pf * *ptot_pf - *pb_pf * *pb_pf);
        if (c->do_rad_losses && aa < 1) *ptot_pf = radiation_loss(h, B2, *ptot_pf, t_step);
        if (*ptot_pf <= 0) {
            *ptot_pf = 1.0e-99; *gam_pf = 1.0; *lose_pt = 1;
            break;
        }
        if (c->compat & MCS_COMPAT_RETRO_KEEP_NEW_PITCH) {
            /* F-6: keep the freshly drawn pitch, rescaled for the momentum lost this step */
            double r = *ptot_pf / ptot_old;
            *pb_pf *= r; *p_perp *= r;
        } else { /* as written (:329-330): the old pitch is restored */
            *pb_pf = *ptot_pf * cos_old; *p_perp = *ptot_pf * sin_old;
        }
        *gam_pf = hypot(1, *ptot_pf / mc);
        if (x < prp_x) break;
        if (steps >= c->retro_cap) { TADD(h, h->n_retro_cap, 1); break; }
    }
    *phi_out = phi;
    *n_steps += steps;
}

/* prob_return.jl:36-173; returns i_return, -9 where the reference calls error() (:134) */
static int prob_return(struct McsHandle* h, Rng* rng, double x_old, double aa, double zz, double* gyro_denom,
                       double* x, double* prp_x, double* ptot_pf, double* gam_pf, double* pb_pf, double* p_perp,
                       double* acctime, double* phi, int helix_count, double pcut_prev, double weight,
                       int* tcut_curr, double mc, int* lose_pt, int64_t* retro_steps, int* went_retro) {
    const McsConfig* c = &h->cfg;
    int i_return = 2;
    *lose_pt = 0;
    if (*x < c->x_grid_stop) {
        /* nothing */
    } else if (x_old < c->x_grid_stop && c->x_grid_stop <= *x) {
        double gyro_tmp = (c->use_custom_epsB && *x > c->x_grid_stop) ? sqrt(c->x_grid_stop / *x) : 1.0;
        double grt = *ptot_pf * c->c_cms * gyro_tmp / (c->qcgs_esu * c->bmag2); /* K-5: Z=1 charge */
        double L = c->eta_mfp / 3 * grt * *ptot_pf / (aa * c->mp_g * *gam_pf * c->u2);
        *prp_x = *x + 3 * L;
    } else if (x_old < *prp_x && *x >= *prp_x) {
        double vt = *ptot_pf / (*gam_pf * aa * c->mp_g);
        double r = (vt - c->u2) / (vt + c->u2), prob_ret = r * r;
        if (vt < c->u2 || rng_uniform(rng) > prob_ret) {
            i_return = 0;
        } else {
            i_return = 1;
            if (!c->do_retro) return -9;
            *went_retro = 1;
            retro_time(h, rng, aa, zz, gyro_denom, *prp_x, ptot_pf, pb_pf, p_perp, gam_pf, acctime, weight,
                       tcut_curr, mc, lose_pt, phi, retro_steps);
            if (*lose_pt) i_return = 0;
            *x = *prp_x;
        }
    } else {
        if (aa < 1 && *ptot_pf < pcut_prev && helix_count % 1000 == 0) {
            double grt = *ptot_pf * c->c_cms * *gyro_denom;
            double L = c->eta_mfp / 3 * grt * *ptot_pf / (aa * c->mp_g * *gam_pf * c->u2);
            if (*x > 2.0e3 * L) *prp_x = 0.8 * *x;
            else *prp_x = fmin(*prp_x, c->x_grid_stop + L * pow(pcut_prev / *ptot_pf, 5));
        }
    }
    return i_return;
}

/* particle_finish.jl:46-107 */
static void particle_finish(struct McsHandle* h, int i_reason, double aa, double pb_pf, double p_perp,
                            double gam_pf, double phi, double ux, double uz, double utot, double gam_sf,
                            double bcos, double bsin, double weight) {
    const McsConfig* c = &h->cfg;
    double m = aa * c->mp_g, E0 = m * (c->c_cms * c->c_cms);
    double ptot_sk, psk[3], gam_sk;
    transform_p_PS(h, aa, pb_pf, p_perp, gam_pf, phi, ux, uz, utot, gam_sf, bcos, bsin, &ptot_sk, psk, &gam_sk);
    int ip = get_psd_bin_momentum(h, ptot_sk), jt = get_psd_bin_angle(h, psk[0], ptot_sk);
    if (ip > MCS_PSD_MAX) ip = MCS_PSD_MAX; /* esc arrays are 0:psd_max; BoundsError otherwise */
    if (jt > MCS_PSD_MAX) jt = MCS_PSD_MAX;
    double wf;
    if (ptot_sk > fabs(PF_SPIKE_AWAY * psk[0])) wf = gam_sk * m * PF_SPIKE_AWAY / ptot_sk;
    else wf = gam_sk * (m / fabs(psk[0]));
    if (i_reason == 1) {
        TADD(h, h->esc_dn[ip + (MCS_PSD_MAX + 1) * jt], weight * wf);
    } else if (i_reason == 2) {
        TADD(h, h->esc_flux, weight);
        TADD(h, h->esc_up[ip + (MCS_PSD_MAX + 1) * jt], weight * wf);
        int rel = (gam_sk - 1) >= c->E_rel_pt; /* F-8: `E_rel_pt / E0` is a unit error; cf. all_flux.jl:104 */
        double Ek = rel ? (gam_sk - 1) * E0 : ptot_sk * ptot_sk / (2 * m);
        double en_add = Ek * weight;
        TADD(h, h->px_esc_feb, fabs(psk[0]) * weight);
        TADD(h, h->en_esc_feb, en_add);
        TADD(h, h->esc_en_eff[ip], en_add);
        TADD(h, h->esc_num_eff[ip], weight);
    }
    /* i_reason 3, 4: nothing (:98-103) */
}

static void trace_push(struct McsHandle* h, int slot, double x, double ptot, double pb, double phi, double acct,
                       double prp, int i_grid, int helix, int down, int inj, int retro, int i_return, int64_t nd) {
    if (slot < 0) return;
    int k = h->trace_cnt[slot];
    if (k >= h->trace_max) return;
    McsTraceRec* r = &h->trace_recs[(size_t)slot * h->trace_max + k];
    r->x_cm = x; r->ptot_pf = ptot; r->pb_pf = pb; r->phi_rad = phi; r->acctime_sec = acct; r->prp_x_cm = prp;
    r->i_grid = i_grid; r->helix_count = helix;
    r->flags = (down ? 1 : 0) | (inj ? 2 : 0) | (retro ? 4 : 0) | ((i_return + 1) << 8);
    r->n_draws = (int32_t)nd;
    h->trace_cnt[slot] = k + 1;
}

/* particle_loop.jl:1-508 followed by the caller's particle_finish! (main_loops.jl:267-279) */
static void particle_loop(struct McsHandle* h, int64_t ip) {
    const McsConfig* c = &h->cfg;
    const double aa = h->sp.aa, zz = h->sp.zz_esu, m = aa * c->mp_g, mc = m * c->c_cms, cl = c->c_cms;
    Rng rng;
    memset(&rng, 0, sizeof rng);
    rng.mode = c->rng_mode;
    if (rng.mode == MCS_RNG_REPLAY) {
        if (h->replay_u && ip < h->replay_n) {
            rng.ru = h->replay_u + h->replay_off[ip];
            rng.rn = h->replay_off[ip + 1] - h->replay_off[ip];
        }
    } else {
        rng.key[0] = (uint32_t)c->seed; rng.key[1] = (uint32_t)(c->seed >> 32);
        rng.ctr[1] = (uint32_t)(h->first_global + ip);
        rng.ctr[2] = ((uint32_t)h->i_pcut & 0xFFFFu) | ((uint32_t)h->i_ion << 16);
        rng.ctr[3] = (uint32_t)h->i_iter;
    }
    int slot = -1;
    for (int t = 0; t < h->n_trace; t++) if (h->trace_idx[t] == ip) slot = t;

    int helix_count = 0;
    Pop* P = &h->cur;
    double weight = P->weight[ip], ptot_pf = P->ptot[ip], pb_pf = P->pb[ip];
    int i_grid = (int)P->grid[ip], i_grid_old = i_grid;
    int l_down = P->down[ip], inj = P->inj[ip];
    double xn_per = P->xn_per[ip], prp_x = P->prp_x[ip], acctime = P->acctime[ip], phi = P->phi[ip];
    int tcut_curr = (int)P->tcut[ip];
    double x = P->x[ip];

    double gam_pf = hypot(1, ptot_pf / mc);
    double gyro_denom = 1 / (zz * h->bt[i_grid]);
    if (c->use_custom_epsB && x > c->x_grid_stop) gyro_denom *= sqrt(x / c->x_grid_stop);
    double gyro_rad_tot = ptot_pf * cl * gyro_denom;
    double gyro_period = TWO_PI * gam_pf * m * cl * gyro_denom;

    double ux = h->ux[i_grid], uz = h->uz[i_grid], utot = h->ut[i_grid], gsf = h->gsf[i_grid];
    double gef = h->gef[i_grid], bmag = h->bt[i_grid], bth = h->th[i_grid];
    double bsin = sin(bth), bcos = cos(bth);

    int keep = 1, i_return = -1, i_reason = 0, lose_pt = 0, saved = 0, err = 0;
    double t_step = 0.0;
    double p_perp = perpendicular_momentum(h, ptot_pf, pb_pf); /* Code Block 1 */
    double gyro_rad = p_perp * cl * gyro_denom;
    double x_old = 0.0;
    int64_t retro_steps = 0;

    while (keep) {
        helix_count++;
        if (helix_count > c->helix_cap) { i_reason = 1; break; } /* K-1 */
        if (i_return == 1) {
            /* Code Block 1 again: the particle has just come back from retro_time */
            p_perp = perpendicular_momentum(h, ptot_pf, pb_pf);
            gyro_rad = p_perp * cl * gyro_denom;
        } else {
            /* Code Block 3 */
            double ux_o = ux, uz_o = uz, ut_o = utot, gsf_o = gsf, bsin_o = bsin, bcos_o = bcos;
            ux = h->ux[i_grid]; uz = h->uz[i_grid]; utot = h->ut[i_grid]; gsf = h->gsf[i_grid];
            gef = h->gef[i_grid]; bmag = h->bt[i_grid]; bth = h->th[i_grid];
            bsin = sin(bth); bcos = cos(bth);
            if (c->use_custom_epsB && x > c->x_grid_stop) bmag = h->bt[h->n_grid] * sqrt(c->x_grid_stop / x);
            gyro_denom = 1 / (zz * bmag);
            if (ux != ux_o) {
                transform_p_PSP(h, aa, &pb_pf, &p_perp, &gam_pf, &phi, ux_o, uz_o, ut_o, gsf_o, bcos_o, bsin_o, ux,
                                uz, utot, gsf, bcos, bsin, &ptot_pf);
                gyro_rad = p_perp * cl * gyro_denom;
                gyro_rad_tot = ptot_pf * cl * gyro_denom;
            }
            if (c->energy_transfer_frac > 0 && !inj && x_old <= 0 && i_grid_old != i_grid)
                do_energy_transfer(h, i_grid, i_grid_old, &ptot_pf, &pb_pf, &p_perp, &gam_pf, weight, mc, aa);
            if (c->dont_scatter && x > 10 * gyro_rad) { i_return = 0; i_reason = 1; keep = 0; continue; }
            if (ptot_pf > h->sp.pmax_cutoff) {
                double ptot_sk, psk[3], gam_sk;
                transform_p_PS(h, aa, pb_pf, p_perp, gam_pf, phi, ux, uz, utot, gsf, bcos, bsin, &ptot_sk, psk,
                               &gam_sk);
                if (ptot_sk > h->sp.pmax_cutoff) { i_reason = 2; keep = 0; continue; }
            }
            if (inj && x < c->feb_upstream) { i_reason = 2; keep = 0; continue; }
            if (c->age_max > 0 && acctime > c->age_max) { i_reason = 3; keep = 0; continue; }
            if (c->do_rad_losses && aa < 1) {
                double p_old = ptot_pf, Bcmb = c->B_CMBz * gef;
                ptot_pf = radiation_loss(h, bmag * bmag + Bcmb * Bcmb, ptot_pf, t_step);
                if (ptot_pf <= 0) {
                    ptot_pf = 1.0e-99; pb_pf = 1.0e-99; p_perp = 1.0e-99; gam_pf = 1;
                    i_reason = 4; keep = 0; continue;
                }
                gam_pf = hypot(ptot_pf / mc, 1);
                pb_pf *= ptot_pf / p_old;
                p_perp *= ptot_pf / p_old;
                gyro_rad_tot = ptot_pf * cl * gyro_denom;
                gyro_rad = p_perp * cl * gyro_denom;
            }
            if (!c->dont_scatter)
                scattering(h, &rng, aa, gyro_denom, ptot_pf, gam_pf, xn_per, &gyro_period, &pb_pf, &p_perp, &phi);
            if (l_down) {
                acctime += t_step * gef;
                if (c->do_tcuts && tcut_curr <= c->n_tcuts && acctime >= c->tcuts[tcut_curr - 1]) {
                    tcut_track(h, tcut_curr, weight, ptot_pf);
                    tcut_curr++;
                }
                if (ptot_pf > h->pcut) { /* :361-380 */
                    Pop* S = &h->saved;
                    h->l_save[ip] = 1; saved = 1;
                    S->weight[ip] = weight; S->ptot[ip] = ptot_pf; S->pb[ip] = pb_pf; S->x[ip] = x;
                    S->grid[ip] = i_grid; S->down[ip] = (uint8_t)l_down; S->inj[ip] = (uint8_t)inj;
                    S->xn_per[ip] = xn_per;
                    S->prp_x[ip] = x < prp_x ? prp_x : x * 1.1; /* F-8: `* 1.1cm` unit slip */
                    S->acctime[ip] = acctime; S->phi[ip] = phi; S->tcut[ip] = tcut_curr;
                    keep = 0; continue;
                }
            }
            xn_per = x > gyro_rad_tot ? c->xn_per_coarse : c->xn_per_fine;
        }

        /* Code Block 2 */
        x_old = x;
        double phi_old = phi;
        t_step = gyro_period / xn_per;
        /* no_DSA_loop :510-571 */
        for (int pass = 0;; pass++) {
            phi = mod2pi(phi + TWO_PI / xn_per);
            double x_move = pb_pf * t_step / (gam_pf * m);
            double dx = gsf * (x_move * bcos - gyro_rad * bsin * (cos(phi) - cos(phi_old)) + ux * t_step);
            x = x_old + dx;
            if (x <= 0 && x_old > 0 && !inj && (c->dont_DSA || c->inj_fracs[h->i_ion - 1] < 1)) {
                if (c->dont_DSA || rng_uniform(&rng) > c->inj_fracs[h->i_ion - 1]) {
                    if (pb_pf < 0) pb_pf = -pb_pf; else phi = rng_uniform(&rng) * TWO_PI;
                } else break;
            } else break;
            if (pass > 1000) { err = 1; break; }
        }
        if (x_old < 0 && x >= 0) {
            l_down = 1;
            double L = c->eta_mfp / 3 * gyro_rad_tot * ptot_pf / (m * gam_pf * c->u2);
            prp_x = fmax(prp_x, L);
        }
        if (l_down && x < 0) inj = 1;

        if (err || all_flux(h, aa, pb_pf, p_perp, ptot_pf, gam_pf, phi, weight, &i_grid, &i_grid_old, ux, uz, utot,
                            gsf, bcos, bsin, x, x_old, inj) != 0) {
            err = 1; break; /* reference: error() all_flux.jl:73-75 (NaN position) */
        }

        /* downstream_test :595-637 */
        int do_prob_ret = 1, went_retro = 0;
        if (c->feb_downstream > 0 && x > c->feb_downstream) {
            i_return = 0; do_prob_ret = 0;
        } else if (x > 1.1 * prp_x) {
            double v_fac;
            if (aa < 1 && ptot_pf < c->pe_crit) {
                double gyro_fac = c->pe_crit * cl * gyro_denom;
                v_fac = gyro_fac * c->pe_crit / (m * c->gam_e_crit * c->u2);
            } else {
                v_fac = gyro_rad_tot * ptot_pf / (m * gam_pf * c->u2);
            }
            double L = c->eta_mfp / 3 * v_fac;
            if (x > 6.91 * L) { i_return = 0; do_prob_ret = 0; }
        }
        if (do_prob_ret) {
            i_return = prob_return(h, &rng, x_old, aa, zz, &gyro_denom, &x, &prp_x, &ptot_pf, &gam_pf, &pb_pf, &p_perp,
                                   &acctime, &phi, helix_count, h->pcut_prev, weight, &tcut_curr, mc, &lose_pt,
                                   &retro_steps, &went_retro);
            if (i_return == -9) { err = 1; break; }
        }
        trace_push(h, slot, x, ptot_pf, pb_pf, phi, acctime, prp_x, i_grid, helix_count, l_down, inj, went_retro,
                   i_return, rng.n);
        if (i_return == 0) {
            double vel = ptot_pf / m;
            if ((gam_pf - 1) >= c->E_rel_pt) vel /= gam_pf;
            TADD(h, h->sumP, ptot_pf / 3 * vel * weight * h->sp.n0);
            TADD(h, h->sumKE, (gam_pf - 1) * m * (cl * cl) * weight * h->sp.n0);
            i_reason = lose_pt ? 4 : 1;
            keep = 0; continue;
        }
        if (rng.exhausted) { err = 1; break; }
    }
    if (err || rng.exhausted) { i_reason = MCS_FATE_ERROR; saved = 0; h->l_save[ip] = 0; TADD(h, h->n_err, 1); }
    if (!saved && i_reason >= 1 && i_reason <= 4)
        particle_finish(h, i_reason, aa, pb_pf, p_perp, gam_pf, phi, ux, uz, utot, gsf, bcos, bsin, weight);
    int fate = saved ? MCS_FATE_SAVED : i_reason;
    h->fate[ip] = fate; h->helix[ip] = helix_count; h->retro[ip] = retro_steps; h->draws[ip] = rng.n;
    TADD(h, h->n_helix, helix_count);
    TADD(h, h->n_retro, retro_steps);
    TADD(h, h->n_fate[fate], 1);
}

/* ------------------------------------------------------------------------------------------ */
/* C-ABI                                                                                        */
int mcs_abi_sizes(int32_t out[6]) {
    out[0] = (int32_t)sizeof(McsConfig); out[1] = (int32_t)sizeof(McsSpecies); out[2] = (int32_t)sizeof(McsTallies);
    out[3] = (int32_t)sizeof(McsPopulation); out[4] = (int32_t)sizeof(McsTraceRec); out[5] = (int32_t)sizeof(McsTiming);
    return MCS_OK;
}

void mcs_default_config(McsConfig* c) {
    memset(c, 0, sizeof *c);
    c->abi_version = MCS_ABI_VERSION; c->device = -1;
    c->mp_g = 1.67262192369e-24; c->c_cms = 2.99792458e10; c->qcgs_esu = 4.80320471257e-10;
    c->E_rel_pt = 0.005;
    {
        double me = 9.1093837015e-28, sigT = 6.6524587321e-25, cc = c->c_cms;
        c->rad_loss_fac = 4.0 / 3.0 * cc * sigT / (cc * cc * cc * me * me * 8 * PI); /* constants.jl:30 */
    }
    c->eta_mfp = 1.0; c->xn_per_fine = 2000.0; c->xn_per_coarse = 100.0; c->age_max = -1.0;
    c->pe_crit = -1.0; c->gam_e_crit = -1.0;
    c->n_ions = 1; c->na_cr = 1000000; c->n_pts_max = 100000;
    for (int i = 0; i < MCS_MAX_IONS; i++) c->inj_fracs[i] = 1.0;
    c->do_retro = 1;
    c->helix_cap = 10000; c->retro_cap = 10000000; c->seed = 210; c->compat = MCS_COMPAT_DEFAULT;
    c->rng_mode = MCS_RNG_PHILOX; c->threads = 1; c->det_tallies = 1; /* CUDA library only; kept identical here */
}

static size_t psd_len(const struct McsHandle* h) { return (size_t)(h->M + 2) * (size_t)(h->T + 2) * (size_t)h->n_grid; }

int mcs_create(const McsConfig* cfg, McsHandle** out) {
    if (!cfg || !out) return fail(MCS_ERR_ARG, "null argument");
    if (cfg->abi_version != MCS_ABI_VERSION) return fail(MCS_ERR_ARG, "abi_version mismatch");
    if (cfg->n_grid < 1 || cfg->n_pts_max < 1 || cfg->n_ions < 1 || cfg->n_ions > MCS_MAX_IONS ||
        cfg->n_tcuts > MCS_NA_C || cfg->n_xspec > MCS_MAX_XSPEC || cfg->num_psd_mom_bins < 1 ||
        cfg->num_psd_theta_bins < 1)
        return fail(MCS_ERR_ARG, "bad sizes in McsConfig");
    if (cfg->use_custom_frg) return fail(MCS_ERR_UNSUPPORTED, "use_custom_frg: reference errors too (scattering.jl:53)");
    struct McsHandle* h = calloc(1, sizeof *h);
    if (!h) return fail(MCS_ERR_NOMEM, "calloc");
    h->cfg = *cfg;
    h->n_grid = cfg->n_grid; h->M = cfg->num_psd_mom_bins; h->T = cfg->num_psd_theta_bins;
    h->par = cfg->threads > 1;
    int ng2 = h->n_grid + 2, ng = h->n_grid;
    int64_t N = cfg->n_pts_max, L = cfg->na_cr > 0 ? cfg->na_cr : 1;
    h->xg = calloc(ng2, 8); h->ux = calloc(ng2, 8); h->uz = calloc(ng2, 8); h->ut = calloc(ng2, 8);
    h->gsf = calloc(ng2, 8); h->gef = calloc(ng2, 8); h->bef = calloc(ng2, 8); h->bt = calloc(ng2, 8);
    h->th = calloc(ng2, 8); h->eps_target = calloc(ng, 8); h->recv_pool = calloc(ng, 8);
    int bad = pop_alloc(&h->cur, N) | pop_alloc(&h->saved, N) | pop_alloc(&h->next, N);
    h->l_save = calloc(N, 1); h->fate = calloc(N, 4); h->helix = calloc(N, 4); h->retro = calloc(N, 8);
    h->draws = calloc(N, 8);
    h->pxx = calloc(ng, 8); h->pxz = calloc(ng, 8); h->efl = calloc(ng, 8); h->psd = calloc(psd_len(h), 8);
    h->ncross = calloc(ng, 8);
    h->tg = calloc(L, 8); h->tpx = calloc(L, 8); h->tpt = calloc(L, 8); h->tw = calloc(L, 8);
    size_t e2 = (size_t)(MCS_PSD_MAX + 1) * (MCS_PSD_MAX + 1), e1 = MCS_PSD_MAX + 1;
    h->esc_up = calloc(e2, 8); h->esc_dn = calloc(e2, 8); h->esc_en_eff = calloc(e1, 8); h->esc_num_eff = calloc(e1, 8);
    h->w_coupled = calloc(MCS_NA_C, 8); h->s_coupled = calloc(e1 * MCS_NA_C, 8); h->pool = calloc(ng, 8);
    h->spec_sf = calloc(e1 * MCS_MAX_XSPEC, 8); h->spec_pf = calloc(e1 * MCS_MAX_XSPEC, 8);
    h->th_sf = calloc(cfg->bin_thermal ? psd_len(h) : 1, 8); h->th_pf = calloc(cfg->bin_thermal ? psd_len(h) : 1, 8);
    if (bad || !h->xg || !h->psd || !h->tg || !h->tw || !h->l_save || !h->draws || !h->spec_pf) {
        mcs_destroy(h);
        return fail(MCS_ERR_NOMEM, "allocation failed");
    }
    *out = h;
    return MCS_OK;
}

int mcs_destroy(McsHandle* h) {
    if (!h) return MCS_OK;
    free(h->xg); free(h->ux); free(h->uz); free(h->ut); free(h->gsf); free(h->gef); free(h->bef); free(h->bt);
    free(h->th); free(h->eps_target); free(h->recv_pool);
    pop_free(&h->cur); pop_free(&h->saved); pop_free(&h->next);
    free(h->l_save); free(h->fate); free(h->helix); free(h->retro); free(h->draws);
    free(h->pxx); free(h->pxz); free(h->efl); free(h->psd); free(h->ncross);
    free(h->tg); free(h->tpx); free(h->tpt); free(h->tw);
    free(h->esc_up); free(h->esc_dn); free(h->esc_en_eff); free(h->esc_num_eff); free(h->w_coupled);
    free(h->s_coupled); free(h->pool); free(h->spec_sf); free(h->spec_pf); free(h->th_sf); free(h->th_pf);
    free(h->replay_u); free(h->replay_off); free(h->trace_idx); free(h->trace_recs); free(h->trace_cnt);
    free(h);
    return MCS_OK;
}

int mcs_comm_unique_id(void* id128) { memset(id128, 0, 128); return MCS_OK; }
int mcs_comm_init(McsHandle* h, int rank, int nranks, const void* id) {
    (void)h; (void)rank; (void)id;
    return nranks == 1 ? MCS_OK : fail(MCS_ERR_UNSUPPORTED, "oracle is single-rank; the host driver reduces");
}

int mcs_set_profile(McsHandle* h, int32_t n_grid, const double* xg, const double* ux, const double* uz,
                    const double* ut, const double* gsf, const double* gef, const double* bef, const double* bt,
                    const double* th, const double* eps_target, const double* recv_pool) {
    if (!h || n_grid != h->n_grid) return fail(MCS_ERR_ARG, "n_grid mismatch");
    size_t b2 = (size_t)(n_grid + 2) * 8, b = (size_t)n_grid * 8;
    memcpy(h->xg, xg, b2); memcpy(h->ux, ux, b2); memcpy(h->uz, uz, b2); memcpy(h->ut, ut, b2);
    memcpy(h->gsf, gsf, b2); memcpy(h->gef, gef, b2); memcpy(h->bef, bef, b2); memcpy(h->bt, bt, b2);
    memcpy(h->th, th, b2);
    if (eps_target) memcpy(h->eps_target, eps_target, b); else memset(h->eps_target, 0, b);
    if (recv_pool) memcpy(h->recv_pool, recv_pool, b); else memset(h->recv_pool, 0, b);
    h->have_profile = 1;
    return MCS_OK;
}

static void zero_ion_tallies(struct McsHandle* h) {
    size_t ng = (size_t)h->n_grid, e1 = MCS_PSD_MAX + 1, e2 = e1 * e1;
    memset(h->pxx, 0, ng * 8); memset(h->pxz, 0, ng * 8); memset(h->efl, 0, ng * 8);
    memset(h->psd, 0, psd_len(h) * 8); memset(h->ncross, 0, ng * 8);
    h->n_cr = 0; h->n_cr_over = 0;
    memset(h->esc_up, 0, e2 * 8); memset(h->esc_dn, 0, e2 * 8); memset(h->esc_en_eff, 0, e1 * 8);
    memset(h->esc_num_eff, 0, e1 * 8); memset(h->w_coupled, 0, MCS_NA_C * 8);
    memset(h->s_coupled, 0, e1 * MCS_NA_C * 8); memset(h->pool, 0, ng * 8);
    memset(h->spec_sf, 0, e1 * MCS_MAX_XSPEC * 8); memset(h->spec_pf, 0, e1 * MCS_MAX_XSPEC * 8);
    if (h->cfg.bin_thermal) { memset(h->th_sf, 0, psd_len(h) * 8); memset(h->th_pf, 0, psd_len(h) * 8); }
    h->esc_flux = h->px_esc_feb = h->en_esc_feb = h->sumP = h->sumKE = h->px_esc_up = h->en_esc_up = 0.0;
    h->n_helix = h->n_retro = h->w_pperp = h->w_psdmom = h->n_negsqrt = h->n_retro_cap = h->n_err = 0;
    memset(h->n_fate, 0, sizeof h->n_fate);
}

int mcs_begin_ion(McsHandle* h, int32_t i_iter, int32_t i_ion, const McsSpecies* sp, int64_t n, int64_t first_global,
                  const McsPopulation* pop) {
    if (!h || !sp || !pop) return fail(MCS_ERR_ARG, "null argument");
    if (!h->have_profile) return fail(MCS_ERR_STATE, "mcs_set_profile first");
    if (n < 0 || n > h->cfg.n_pts_max) return fail(MCS_ERR_ARG, "n_pts exceeds n_pts_max");
    if (i_ion < 1 || i_ion > h->cfg.n_ions) return fail(MCS_ERR_ARG, "i_ion out of range");
    /* the widths of the Philox counter fields (same checks, same messages as the CUDA library) */
    if (first_global < 0 || first_global + n > 0x100000000ll) return fail(MCS_ERR_ARG, "global particle index exceeds 2^32 (RNG counter width)");
    if (n > 0 && (!pop->weight || !pop->ptot_pf || !pop->pb_pf || !pop->x_cm || !pop->grid || !pop->phi_rad))
        return fail(MCS_ERR_ARG, "weight/ptot_pf/pb_pf/x_cm/grid/phi_rad are required");
    h->sp = *sp; h->i_iter = i_iter; h->i_ion = i_ion; h->i_pcut = 0;
    h->n_use = n; h->first_global = first_global; h->n_saved_last = 0;
    zero_ion_tallies(h);
    Pop* P = &h->cur;
    for (int64_t i = 0; i < n; i++) {
        P->weight[i] = pop->weight[i]; P->ptot[i] = pop->ptot_pf[i]; P->pb[i] = pop->pb_pf[i];
        P->x[i] = pop->x_cm[i]; P->grid[i] = pop->grid[i]; P->phi[i] = pop->phi_rad[i];
        P->down[i] = pop->downstream ? pop->downstream[i] : 0;
        P->inj[i] = pop->inj ? pop->inj[i] : 0;
        P->xn_per[i] = pop->xn_per ? pop->xn_per[i] : h->cfg.xn_per_fine;
        P->prp_x[i] = pop->prp_x_cm ? pop->prp_x_cm[i] : h->cfg.x_grid_stop;
        P->acctime[i] = pop->acctime_sec ? pop->acctime_sec[i] : 0.0;
        P->tcut[i] = pop->tcut ? pop->tcut[i] : 1;
        if (P->grid[i] < 0 || P->grid[i] > h->n_grid + 1) return fail(MCS_ERR_ARG, "grid index out of range");
    }
    h->have_ion = 1; h->ended = 0;
    return MCS_OK;
}

/* init_pop in run-length form (include/mcs.h McsInjection; initializers.jl:977-1134, ion_init.jl:29-53) */
static int64_t inj_origin(int64_t s, int64_t n, int64_t K) {
    if (K <= 0) return s;
    int64_t a = n / K, b = n % K; /* the first b sub-sequences hold a+1 particles, the rest a */
    if (s < b * (a + 1)) return s / (a + 1) + K * (s % (a + 1));
    s -= b * (a + 1);
    return (b + s / a) + K * (s % a);
}
int mcs_begin_ion_generate(McsHandle* h, int32_t i_iter, int32_t i_ion, const McsSpecies* sp, int64_t first_global,
                           int64_t n_local, const McsInjection* inj) {
    if (!h || !sp || !inj) return fail(MCS_ERR_ARG, "null argument");
    if (!h->have_profile) return fail(MCS_ERR_STATE, "mcs_set_profile first");
    if (i_ion < 1 || i_ion > h->cfg.n_ions) return fail(MCS_ERR_ARG, "i_ion out of range");
    if (inj->n_bins < 1 || !inj->bin_ptot || !inj->bin_weight || !inj->bin_start) return fail(MCS_ERR_ARG, "injection bins missing");
    if (inj->mode < 0 || inj->mode > 2) return fail(MCS_ERR_ARG, "injection mode");
    if (inj->mode != MCS_INJ_UPSTREAM && (!inj->bin_lo || !inj->bin_hi || !inj->bin_gfac)) return fail(MCS_ERR_ARG, "fast-push bins missing");
    int64_t n_total = inj->bin_start[inj->n_bins];
    if (n_local < 0 || first_global < 0 || first_global + n_local > n_total) return fail(MCS_ERR_ARG, "shard outside the population");
    if (n_local > h->cfg.n_pts_max) return fail(MCS_ERR_ARG, "n_pts exceeds n_pts_max");
    if (inj->grid < 0 || inj->grid > h->n_grid + 1) return fail(MCS_ERR_ARG, "grid index out of range");
    h->sp = *sp; h->i_iter = i_iter; h->i_ion = i_ion; h->i_pcut = 0;
    h->n_use = n_local; h->first_global = first_global; h->n_saved_last = 0;
    zero_ion_tallies(h);
    Pop* P = &h->cur;
    const uint32_t key[2] = {(uint32_t)h->cfg.seed, (uint32_t)(h->cfg.seed >> 32)};
    const double bu = inj->u_stop / h->cfg.c_cms;
    for (int64_t i = 0; i < n_local; i++) {
        int64_t j = inj_origin(first_global + i, n_total, inj->perm_stride);
        int lo = 0, hi = inj->n_bins; /* bin_start[lo] <= j < bin_start[hi] */
        while (hi - lo > 1) { int mid = (lo + hi) / 2; if (inj->bin_start[mid] <= j) lo = mid; else hi = mid; }
        const uint32_t ctr[4] = {0u, (uint32_t)j, (uint32_t)i_ion << 16, (uint32_t)i_iter};
        uint32_t o[4];
        philox4x32_10(ctr, key, o);
        double u1 = u53(o[1], o[0]), u2 = u53(o[3], o[2]);
        double ptot = inj->bin_ptot[lo], pb;
        if (inj->mode == MCS_INJ_UPSTREAM) pb = (ptot * 2) * (u1 - 0.5);
        else {
            double vx = inj->bin_lo[lo] + (inj->bin_hi[lo] - inj->bin_lo[lo]) * sqrt(u1);
            if (inj->mode == MCS_INJ_FASTPUSH_REL) pb = inj->bin_gfac[lo] * ((vx - bu) / (1 - vx * bu) * h->cfg.c_cms);
            else pb = inj->bin_gfac[lo] * (vx - inj->u_stop);
        }
        P->weight[i] = inj->bin_weight[lo]; P->ptot[i] = ptot; P->pb[i] = pb; P->x[i] = inj->x_cm; P->grid[i] = inj->grid;
        P->phi[i] = TWO_PI * u2;
        P->down[i] = 0; P->inj[i] = 0; P->xn_per[i] = h->cfg.xn_per_fine; P->prp_x[i] = h->cfg.x_grid_stop;
        P->acctime[i] = 0.0; P->tcut[i] = 1;
    }
    h->have_ion = 1; h->ended = 0;
    return MCS_OK;
}

int mcs_run_pcut(McsHandle* h, int32_t i_pcut, double pcut, double pcut_prev, int64_t* n_saved, int64_t* n_steps) {
    if (!h || !h->have_ion) return fail(MCS_ERR_STATE, "mcs_begin_ion first");
    if (i_pcut < 0 || i_pcut > 0xFFFF) return fail(MCS_ERR_ARG, "i_pcut outside the RNG counter's 16-bit field");
    h->i_pcut = i_pcut; h->pcut = pcut; h->pcut_prev = pcut_prev;
    int64_t n = h->n_use;
    /* main_loops.jl:184-197 */
    memset(h->l_save, 0, (size_t)h->cfg.n_pts_max);
    pop_zero(&h->saved, n);
    for (int t = 0; t < h->n_trace; t++) h->trace_cnt[t] = 0;
    int64_t s0 = h->n_helix + h->n_retro;
    if (h->par) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(h->cfg.threads)
        for (int64_t i = 0; i < n; i++) particle_loop(h, i);
    } else {
        for (int64_t i = 0; i < n; i++) particle_loop(h, i);
    }
    int64_t ns = 0;
    for (int64_t i = 0; i < n; i++) ns += h->l_save[i]; /* pcut_finalize: count(l_save), cuts.jl:105 */
    h->n_saved_last = ns;
    if (n_saved) *n_saved = ns;
    if (n_steps) *n_steps = h->n_helix + h->n_retro - s0;
    return MCS_OK;
}

/* new_pcut, cuts.jl:34-98 */
int mcs_split_explicit(McsHandle* h, int64_t i_mult, int64_t first_global_child, int64_t* n_new_local) {
    if (!h || !h->have_ion) return fail(MCS_ERR_STATE, "mcs_begin_ion first");
    if (i_mult < 1) return fail(MCS_ERR_ARG, "i_mult < 1");
    if (h->n_saved_last * i_mult > h->cfg.n_pts_max) return fail(MCS_ERR_ARG, "split population exceeds n_pts_max");
    Pop *S = &h->saved, *Nw = &h->next;
    int64_t k = 0;
    for (int64_t j = 0; j < h->n_use; j++) {
        if (!h->l_save[j]) continue;
        for (int64_t i = 0; i < i_mult; i++, k++) {
            Nw->weight[k] = S->weight[j] / (double)i_mult;
            Nw->ptot[k] = S->ptot[j]; Nw->pb[k] = S->pb[j]; Nw->x[k] = S->x[j]; Nw->grid[k] = S->grid[j];
            Nw->down[k] = S->down[j]; Nw->inj[k] = S->inj[j]; Nw->xn_per[k] = S->xn_per[j];
            Nw->prp_x[k] = S->prp_x[j]; Nw->acctime[k] = S->acctime[j]; Nw->phi[k] = S->phi[j];
            Nw->tcut[k] = S->tcut[j];
        }
    }
    Pop tmp = h->cur; h->cur = h->next; h->next = tmp;
    h->n_use = k;
    h->first_global = first_global_child;
    if (n_new_local) *n_new_local = k;
    return MCS_OK;
}

int mcs_split(McsHandle* h, int64_t n_pts_target, int64_t* n_new_local, int64_t* n_new_global, int64_t* i_mult_out) {
    if (!h || !h->have_ion) return fail(MCS_ERR_STATE, "mcs_begin_ion first");
    int64_t ns = h->n_saved_last;
    if (ns <= 0) return fail(MCS_ERR_STATE, "no saved particles to split");
    int64_t i_mult = n_pts_target / ns; /* cuts.jl:42 */
    if (i_mult < 1) i_mult = 1;
    int64_t k = 0;
    int rc = mcs_split_explicit(h, i_mult, 0, &k);
    if (rc) return rc;
    if (n_new_local) *n_new_local = k;
    if (n_new_global) *n_new_global = k;
    if (i_mult_out) *i_mult_out = i_mult;
    return MCS_OK;
}

/* loop_pcut, main_loops.jl:179-317 */
int mcs_run_ion(McsHandle* h, const double* pcuts, int32_t n_pcuts, double p_pcut_hi, int64_t n_pts_pcut,
                int64_t n_pts_pcut_hi, int32_t* n_run, int64_t* n_used, int64_t* n_saved_arr) {
    if (!h || !pcuts || n_pcuts < 1 || n_pcuts > MCS_NA_C) return fail(MCS_ERR_ARG, "bad pcuts");
    int32_t k = 0;
    for (int32_t i = 1; i <= n_pcuts; i++) {
        int64_t ns = 0, nst = 0;
        if (n_used) n_used[i - 1] = h->n_use;
        int rc = mcs_run_pcut(h, i, pcuts[i - 1], i > 1 ? pcuts[i - 2] : 0.0, &ns, &nst);
        if (rc) return rc;
        if (n_saved_arr) n_saved_arr[i - 1] = ns;
        k = i;
        if (ns == 0) break; /* pcut_finalize: break_pcut */
        int64_t target = pcuts[i - 1] < p_pcut_hi ? n_pts_pcut : n_pts_pcut_hi;
        rc = mcs_split(h, target, NULL, NULL, NULL);
        if (rc) return rc;
    }
    if (n_run) *n_run = k;
    return MCS_OK;
}

#define CPY(dst, src, n) do { if (dst) memcpy(dst, src, (size_t)(n) * sizeof *(dst)); } while (0)
int mcs_end_ion(McsHandle* h, McsTallies* t) {
    if (!h || !t) return fail(MCS_ERR_ARG, "null argument");
    size_t ng = (size_t)h->n_grid, e1 = MCS_PSD_MAX + 1, e2 = e1 * e1;
    CPY(t->pxx_flux, h->pxx, ng); CPY(t->pxz_flux, h->pxz, ng); CPY(t->energy_flux, h->efl, ng);
    CPY(t->psd, h->psd, psd_len(h)); CPY(t->num_crossings, h->ncross, ng);
    t->n_cr_count = h->n_cr; t->n_cr_overflow = h->n_cr_over;
    CPY(t->therm_grid, h->tg, h->n_cr); CPY(t->therm_px_sk, h->tpx, h->n_cr);
    CPY(t->therm_ptot_sk, h->tpt, h->n_cr); CPY(t->therm_weight, h->tw, h->n_cr);
    CPY(t->esc_psd_feb_upstream, h->esc_up, e2); CPY(t->esc_psd_feb_downstream, h->esc_dn, e2);
    CPY(t->esc_energy_eff, h->esc_en_eff, e1); CPY(t->esc_num_eff, h->esc_num_eff, e1);
    CPY(t->weight_coupled, h->w_coupled, MCS_NA_C); CPY(t->spectra_coupled, h->s_coupled, e1 * MCS_NA_C);
    CPY(t->energy_transfer_pool, h->pool, ng);
    CPY(t->spectra_sf, h->spec_sf, e1 * (size_t)h->cfg.n_xspec); CPY(t->spectra_pf, h->spec_pf, e1 * (size_t)h->cfg.n_xspec);
    if (h->cfg.bin_thermal) {
        CPY(t->therm_d2N_sf, h->th_sf, psd_len(h)); CPY(t->therm_d2N_pf, h->th_pf, psd_len(h));
        if (t->dNdp_cr_sf) { /* particle_counter.jl:81-85 */
            for (int i = 0; i < h->n_grid; i++)
                for (int k = 0; k < h->M + 2; k++) {
                    double sum = 0.0;
                    for (int j = 0; j < h->T + 2; j++) {
                        double v = h->psd[(size_t)k + (size_t)(h->M + 2) * ((size_t)j + (size_t)(h->T + 2) * (size_t)i)];
                        if (v > 0) sum += v;
                    }
                    t->dNdp_cr_sf[(size_t)k + (size_t)(h->M + 2) * (size_t)i] = sum;
                }
        }
    }
    t->esc_flux = h->esc_flux; t->px_esc_feb = h->px_esc_feb; t->energy_esc_feb = h->en_esc_feb;
    t->sum_P_downstream = h->sumP; t->sum_KE_downstream = h->sumKE;
    t->px_esc_upstream = h->px_esc_up; t->energy_esc_upstream = h->en_esc_up;
    t->n_helix_steps = h->n_helix; t->n_retro_steps = h->n_retro;
    t->n_warn_pperp = h->w_pperp; t->n_warn_psd_mom = h->w_psdmom; t->n_neg_sqrt = h->n_negsqrt;
    t->n_retro_capped = h->n_retro_cap; t->n_errors = h->n_err;
    memcpy(t->n_fate, h->n_fate, sizeof t->n_fate);
    h->ended = 1;
    return MCS_OK;
}

/* thermo_calcs.jl:31-355 on the binned tallies.  The thermal log of the reference (:96-164) is replaced by
 * th_pf, the same crossings binned as they happen (see all_flux above, which applies :143-160 per crossing). */
int mcs_thermo(McsHandle* h, const McsThermoIn* in, double* P_par, double* P_perp, double* e_dens, double* d2N_pop_out) {
    if (!h || !in || !in->cos_center || !in->pt_center || !in->zone_pop) return fail(MCS_ERR_ARG, "null argument");
    if (!h->cfg.bin_thermal) return fail(MCS_ERR_ARG, "mcs_thermo needs cfg.bin_thermal = 1");
    if (!h->have_profile) return fail(MCS_ERR_STATE, "mcs_set_profile first");
    const int resident = !(in->psd && in->therm_d2N_pf && in->num_crossings);
    if (resident && (in->psd || in->therm_d2N_pf || in->num_crossings))
        return fail(MCS_ERR_ARG, "give psd, therm_d2N_pf and num_crossings together or none of them");
    if (resident && !h->ended) return fail(MCS_ERR_STATE, "mcs_end_ion first: the tallies are not folded / summed over ranks yet");
    const McsConfig* c = &h->cfg;
    const int ng = h->n_grid, T2 = h->T + 2, M2 = h->M + 2;
    const size_t slab = (size_t)T2 * (size_t)M2;
    const double* psd = in->psd ? in->psd : h->psd;
    const double* thp = in->therm_d2N_pf ? in->therm_d2N_pf : h->th_pf;
    const int64_t* ncr = in->num_crossings ? in->num_crossings : h->ncross;
    const double cl = c->c_cms, m = h->sp.aa * c->mp_g, mc = m * cl, E0 = m * (cl * cl); /* :53 */
    const double kB = 1.380649e-16; /* Unitful k, erg/K */
    double* d2N = malloc(slab * 8);
    if (!d2N) return fail(MCS_ERR_NOMEM, "out of memory");
    for (int i = 1; i <= ng; i++) {
        const double g = h->gsf[i], b = h->ux[i] / cl;
        for (size_t q = 0; q < slab; q++) d2N[q] = 1.0e-99 + thp[q + slab * (size_t)(i - 1)]; /* :43, :133-164 */
        /* :178-207 — CR cells re-binned by the boost of their centre */
        for (int jt = 0; jt <= h->T; jt++)
            for (int k = 0; k <= h->M; k++) {
                double cell = psd[(size_t)k + (size_t)M2 * ((size_t)jt + (size_t)T2 * (size_t)(i - 1))];
                if (cell <= 1.0e-66) continue;
                double pt = in->pt_center[k], px = pt * in->cos_center[jt];
                double etot = hypot(pt * cl, E0);
                double pxX = g * (px - b * etot / cl);
                double ptX = sqrt(pt * pt - px * px + pxX * pxX);
                int kX = get_psd_bin_momentum(h, ptX), jX = get_psd_bin_angle(h, pxX, ptX);
                d2N[(size_t)jX + (size_t)T2 * (size_t)kX] += cell;
            }
        /* :209-226 — normalise to the zone population */
        double nf = 0.0;
        for (size_t q = 0; q < slab; q++) if (d2N[q] > 1.0e-66) nf += d2N[q];
        if (ncr[i - 1] == 0 && nf > 0) nf += h->sp.n0 / h->ux[i];
        if (nf > 0) nf = in->zone_pop[i - 1] / nf;
        double pop = 0.0;
        for (size_t q = 0; q < slab; q++) if (d2N[q] > 1.0e-66) { d2N[q] *= nf; }
        for (size_t q = 0; q < slab; q++) if (d2N[q] > 1.0e-66) pop += d2N[q];
        if (d2N_pop_out) d2N_pop_out[i - 1] = pop;
        /* :242-352 — the three normalisation cases, then the sums */
        double dmax = 0.0;
        for (size_t q = 0; q < slab; q++) if (d2N[q] > dmax) dmax = d2N[q];
        const double dens = c->gam0 * c->beta0 * h->sp.n0 / sqrt(g * g - 1.0);
        double par = 0.0, perp = 0.0, en = 0.0, norm = 0.0;
        int sum_cells = 1;
        if (dmax < 1.0e-66 && ncr[i - 1] == 0) { /* (1) nothing detected: cold thermal gas, Gamma = 5/3 */
            double pl = pow(dens, 5.0 / 3.0) * kB * in->temperature_K;
            par += 1.0 / 3.0 * pl; perp += 2.0 / 3.0 * pl; en += 1.5 * pl;
            sum_cells = 0;
        } else if (ncr[i - 1] == 0) { /* (2) CRs only */
            double pl = pow(dens, 5.0 / 3.0) * kB * in->temperature_K;
            pl *= 1.0 - pop / in->zone_pop[i - 1];
            par += 1.0 / 3.0 * pl; perp += 2.0 / 3.0 * pl;
            norm = dens / in->zone_pop[i - 1];
            en += 1.5 * pl;
        } else { /* (3) thermal crossings seen: d2N is the whole population */
            norm = dens / in->zone_pop[i - 1];
        }
        if (sum_cells)
            for (int k = 0; k <= h->M; k++) {
                double pt = in->pt_center[k];
                double gt = hypot(1.0, pt / mc);
                double vel = pt * cl / (mc * gt); /* :236 */
                double pf = 1.0 / 3.0 * pt * vel * norm, ef = (gt - 1.0) * E0;
                for (int jt = 0; jt <= h->T; jt++) {
                    double d = d2N[(size_t)jt + (size_t)T2 * (size_t)k];
                    if (d < 1.0e-66) continue;
                    double c2 = in->cos_center[jt] * in->cos_center[jt];
                    par += d * pf * c2; perp += d * pf * (1.0 - c2);
                    en += ef * d * norm;
                }
            }
        if (P_par) P_par[i - 1] = par;
        if (P_perp) P_perp[i - 1] = perp;
        if (e_dens) e_dens[i - 1] = en;
    }
    free(d2N);
    return MCS_OK;
}

int mcs_get_population(McsHandle* h, int32_t which, int64_t n, McsPopulation* o, uint8_t* l_save) {
    if (!h || !o) return fail(MCS_ERR_ARG, "null argument");
    if (n < 0 || n > h->cfg.n_pts_max) return fail(MCS_ERR_ARG, "n out of range");
    Pop* P = which == 0 ? &h->cur : &h->saved;
    CPY(o->weight, P->weight, n); CPY(o->ptot_pf, P->ptot, n); CPY(o->pb_pf, P->pb, n); CPY(o->x_cm, P->x, n);
    CPY(o->xn_per, P->xn_per, n); CPY(o->prp_x_cm, P->prp_x, n); CPY(o->acctime_sec, P->acctime, n);
    CPY(o->phi_rad, P->phi, n); CPY(o->grid, P->grid, n); CPY(o->tcut, P->tcut, n);
    CPY(o->downstream, P->down, n); CPY(o->inj, P->inj, n);
    CPY(l_save, h->l_save, n);
    return MCS_OK;
}
int64_t mcs_population_size(McsHandle* h) { return h ? h->n_use : -1; }

int mcs_get_fates(McsHandle* h, int64_t n, int32_t* fate, int32_t* helix, int64_t* retro, int64_t* draws) {
    if (!h || n < 0 || n > h->cfg.n_pts_max) return fail(MCS_ERR_ARG, "bad argument");
    CPY(fate, h->fate, n); CPY(helix, h->helix, n); CPY(retro, h->retro, n); CPY(draws, h->draws, n);
    return MCS_OK;
}

int mcs_replay_set_stream(McsHandle* h, const double* u, const int64_t* off, int64_t n) {
    if (!h || !u || !off || n < 0) return fail(MCS_ERR_ARG, "bad argument");
    free(h->replay_u); free(h->replay_off);
    int64_t tot = off[n];
    h->replay_u = malloc((size_t)(tot > 0 ? tot : 1) * 8);
    h->replay_off = malloc((size_t)(n + 1) * 8);
    if (!h->replay_u || !h->replay_off) return fail(MCS_ERR_NOMEM, "malloc");
    memcpy(h->replay_u, u, (size_t)tot * 8); memcpy(h->replay_off, off, (size_t)(n + 1) * 8);
    h->replay_n = n;
    return MCS_OK;
}

int mcs_trace_enable(McsHandle* h, const int64_t* idx, int32_t n_trace, int32_t max_steps) {
    if (!h || n_trace < 0 || max_steps < 0) return fail(MCS_ERR_ARG, "bad argument");
    free(h->trace_idx); free(h->trace_recs); free(h->trace_cnt);
    h->trace_idx = NULL; h->trace_recs = NULL; h->trace_cnt = NULL;
    h->n_trace = n_trace; h->trace_max = max_steps;
    if (n_trace == 0) return MCS_OK;
    h->trace_idx = malloc((size_t)n_trace * 8);
    h->trace_recs = calloc((size_t)n_trace * (size_t)(max_steps > 0 ? max_steps : 1), sizeof(McsTraceRec));
    h->trace_cnt = calloc((size_t)n_trace, 4);
    if (!h->trace_idx || !h->trace_recs || !h->trace_cnt) return fail(MCS_ERR_NOMEM, "malloc");
    memcpy(h->trace_idx, idx, (size_t)n_trace * 8);
    return MCS_OK;
}
int mcs_trace_get(McsHandle* h, McsTraceRec* recs, int32_t* n_rec) {
    if (!h || !h->n_trace) return fail(MCS_ERR_STATE, "trace not enabled");
    CPY(recs, h->trace_recs, (size_t)h->n_trace * h->trace_max);
    CPY(n_rec, h->trace_cnt, h->n_trace);
    return MCS_OK;
}

int mcs_get_timing(McsHandle* h, McsTiming* out, int32_t reset) {
    (void)h; (void)reset;
    if (out) memset(out, 0, sizeof *out);
    return MCS_OK;
}
int mcs_measure_fp64_peak(McsHandle* h, double* t) { (void)h; (void)t; return fail(MCS_ERR_UNSUPPORTED, "cpu oracle"); }
int mcs_measure_scatter_peak(McsHandle* h, double* r) { (void)h; (void)r; return fail(MCS_ERR_UNSUPPORTED, "cpu oracle"); }
int mcs_selftest_math(McsHandle* h, int64_t n, int64_t* a, int64_t* b) { (void)h; (void)n; (void)a; (void)b; return fail(MCS_ERR_UNSUPPORTED, "cpu oracle"); }
int mcs_measure_atomic_peak(McsHandle* h, int64_t n, double* g) {
    (void)h; (void)n; (void)g;
    return fail(MCS_ERR_UNSUPPORTED, "cpu oracle");
}

/* ------------------------------------------------------------------------------------------ */
/* Test-only hooks (not part of include/mcs.h): let the known-answer tests call the restated    */
/* reference functions one at a time.                                                           */
MCS_API void mcso_transform_p_PS(McsHandle* h, double aa, double pb, double pperp, double gam_pf, double phi, double ux,
                                 double gsf, double bcos, double bsin, double out[5]) {
    double psk[3];
    transform_p_PS(h, aa, pb, pperp, gam_pf, phi, ux, 0.0, ux, gsf, bcos, bsin, &out[0], psk, &out[4]);
    out[1] = psk[0]; out[2] = psk[1]; out[3] = psk[2];
}
MCS_API void mcso_transform_p_PSP(McsHandle* h, double aa, double io[5] /* ptot pb pperp gam phi */, const double old6[6],
                                  const double new6[6] /* ux uz ut gsf bcos bsin */) {
    transform_p_PSP(h, aa, &io[1], &io[2], &io[3], &io[4], old6[0], old6[1], old6[2], old6[3], old6[4], old6[5], new6[0],
                    new6[1], new6[2], new6[3], new6[4], new6[5], &io[0]);
}
MCS_API void mcso_scattering(McsHandle* h, uint32_t stream, int n_kicks, double aa, double gyro_denom, double ptot,
                             double gam_pf, double xn_per, double io[4] /* gyro_period pb pperp phi */) {
    Rng rng; memset(&rng, 0, sizeof rng);
    rng.key[0] = (uint32_t)h->cfg.seed; rng.key[1] = (uint32_t)(h->cfg.seed >> 32); rng.ctr[1] = stream;
    for (int k = 0; k < n_kicks; k++) scattering(h, &rng, aa, gyro_denom, ptot, gam_pf, xn_per, &io[0], &io[1], &io[2], &io[3]);
}
MCS_API int mcso_psd_bin_momentum(McsHandle* h, double p) { return get_psd_bin_momentum(h, p); }
MCS_API int mcso_psd_bin_angle(McsHandle* h, double px, double p) { return get_psd_bin_angle(h, px, p); }
MCS_API double mcso_radiation_loss(McsHandle* h, double B2, double p, double dt) { return radiation_loss(h, B2, p, dt); }
MCS_API double mcso_mod2pi(double x) { return mod2pi(x); }
MCS_API void mcso_philox(uint64_t seed, uint32_t c1, uint32_t c2, uint32_t c3, int n, double* out) {
    Rng rng; memset(&rng, 0, sizeof rng);
    rng.key[0] = (uint32_t)seed; rng.key[1] = (uint32_t)(seed >> 32); rng.ctr[1] = c1; rng.ctr[2] = c2; rng.ctr[3] = c3;
    for (int i = 0; i < n; i++) out[i] = rng_uniform(&rng);
}
MCS_API void mcso_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
