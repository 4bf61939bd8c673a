/*
 * mcs.h — C-ABI of the B200-native per-particle transport loop.
 *
 * This is the drop-in boundary for ONE hot path of abhro/MonteCarloScattering.jl:
 * the particle loop of `main_loops` (reference src/main_loops.jl:228-292, i.e.
 * `particle_loop` src/particle_loop.jl:1-508 followed by `particle_finish!`
 * src/particle_finish.jl:46-107) and the between-pcut population management
 * (`pcut_finalize` src/cuts.jl:100-124, `new_pcut` src/cuts.jl:34-98).
 *
 * The reference has no FFI of its own (pure Julia); the entry points below are what a
 * `ccall` from main_loops.jl would bind (see INTEGRATION.md for the Julia stub).
 * Conventions: plain C, blocking calls, caller owns every host buffer, the library owns
 * every device buffer, return 0 on success and <0 on error (text via mcs_last_error()).
 * All physical quantities are cgs Float64, bit-identical to the reference's Unitful
 * wrappers (src/cgstypes.jl:8-21).  Grid arrays are the parents of the reference's
 * OffsetVectors: C index k == Julia offset index k for the 0:n_grid+1 axis.
 * Arrays the reference indexes 1:n_grid (pxx_flux, num_crossings, eps_target, pools, the
 * third PSD axis) are passed/returned as n_grid contiguous values, C index i-1.
 */
#ifndef MCS_H
#define MCS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MCS_API __attribute__((visibility("default")))
#else
#define MCS_API
#endif

#define MCS_ABI_VERSION 1

/* compile-time caps mirrored from src/parameters.jl:9-32 */
#define MCS_NA_C 100      /* na_c: max pcuts / tcuts                          */
#define MCS_PSD_MAX 200   /* psd_max: escape-PSD / coupled-spectra axis 0:200 */
#define MCS_MAX_IONS 8
#define MCS_MAX_XSPEC 16

/* error codes */
#define MCS_OK 0
#define MCS_ERR_ARG (-1)
#define MCS_ERR_CUDA (-2)
#define MCS_ERR_STATE (-3)
#define MCS_ERR_NOMEM (-4)
#define MCS_ERR_UNSUPPORTED (-5)
#define MCS_ERR_COMM (-6)

/* rng_mode */
#define MCS_RNG_PHILOX 0 /* Philox4x32-10, counter=(block, i_prt, i_pcut|i_ion<<16, i_iter), key=seed */
#define MCS_RNG_REPLAY 1 /* consume a recorded per-particle uniform stream (mcs_replay_set_stream)  */

/* compat flags: every deviation from the as-written reference is switchable */
#define MCS_COMPAT_RETRO_KEEP_NEW_PITCH 1u /* B-6: keep the large-angle-scattered pitch in retro_time
                                              (prob_return.jl:309-310 vs :329-330); off = as written */
#define MCS_COMPAT_DEFAULT (MCS_COMPAT_RETRO_KEEP_NEW_PITCH)

/* reasons a particle left the helix loop (particle_loop.jl:138, particle_finish.jl:79-105) */
#define MCS_FATE_SAVED 0    /* crossed the current pcut: stored in the *_saved arrays            */
#define MCS_FATE_DOWNSTREAM 1
#define MCS_FATE_FEB_PMAX 2
#define MCS_FATE_AGE 3
#define MCS_FATE_ZERO_ENERGY 4
#define MCS_FATE_ERROR 5    /* NaN position / replay stream exhausted: reference would throw    */

/* Scalars and flags of particle_loop's argument list (particle_loop.jl:1-31). */
typedef struct McsConfig {
    int32_t abi_version; /* = MCS_ABI_VERSION */
    int32_t device;      /* CUDA device ordinal; -1 = current/LOCAL_RANK */

    /* physical constants; filled by mcs_default_config */
    double mp_g, c_cms, qcgs_esu, E_rel_pt, rad_loss_fac;

    /* shock scalars */
    double gam0, beta0, u0, u2, bmag2;
    double pe_crit, gam_e_crit, eta_mfp;

    /* PSD binning (get_psd_bins.jl:16-97) */
    double psd_mom_min, psd_cos_fine, delta_cos, psd_theta_min;
    int32_t psd_bins_per_dec_mom, psd_bins_per_dec_theta;
    int32_t num_psd_mom_bins, num_psd_theta_bins;

    double energy_transfer_frac;
    double feb_upstream, feb_downstream, x_grid_stop;
    double B_CMBz;
    double xn_per_fine, xn_per_coarse;
    double age_max;

    int32_t n_grid;     /* zones; grid arrays have n_grid+2 nodes */
    int32_t i_grid_feb, i_shock;
    int32_t n_ions;
    int64_t n_pts_max;  /* capacity: >= every population size incl. after splitting */
    int64_t na_cr;      /* thermal-crossing log capacity (parameters.jl:25)          */

    int32_t n_xspec;
    double x_spec[MCS_MAX_XSPEC];
    int32_t n_tcuts;
    double tcuts[MCS_NA_C];
    double inj_fracs[MCS_MAX_IONS];

    /* flags (particle_loop.jl:24) */
    int32_t do_rad_losses, do_retro, do_tcuts, dont_DSA, dont_scatter, use_custom_frg, use_custom_epsB;

    /* build-specific */
    int32_t helix_cap;  /* 10000, particle_loop.jl:162 */
    int64_t retro_cap;  /* safety cap on one retro_time call; reference is unbounded */
    uint64_t seed;
    uint32_t compat;
    int32_t rng_mode;
    int32_t threads;    /* CPU oracle only: OpenMP threads (0/1 = serial, ordered) */
    int32_t bin_thermal; /* also bin every thermal crossing on the fly (McsTallies.therm_d2N_*) so that the
                            unbounded crossing log is not needed: 0 = off, 1 = on */
    int32_t dynamic_queue; /* 0 (default) = particles dealt to warps in a fixed interleaved order: the flux arrays and scalars, summed
                              per warp and then over warps and blocks in a fixed order, are bitwise reproducible run to run;
                              1 = global atomic work queue (idle lanes take the next particles: +4 % steps/s, the order of the
                              flux sums then varies in the last bits).  Per-particle results and the exact-accumulator tallies
                              (det_tallies) do not depend on it. */
    int32_t det_tallies;   /* 1 (default) = the phase-space histogram, escape PSDs, coupled / x_spec / efficiency spectra and the
                              energy pool are accumulated as exact fixed-point sums (integer atomics on 10 x 32-bit digits per
                              cell): bitwise identical run to run, for any schedule and any number of GPUs;
                              0 = red.global.add.f64 (order of the adds varies in the last bits).  CUDA library only.
                              The flux arrays and scalars use ordered per-warp partials in either case (see dynamic_queue). */
} McsConfig;

/* Per-species scalars read inside the loop (main_loops.jl:97-100, utils.jl:72-96). */
typedef struct McsSpecies {
    double aa;          /* mass / m_p                                           */
    double zz_esu;      /* charge as used in gyro_denom = 1/(zz*B)              */
    double n0;          /* far-upstream number density, density(species[i_ion]) */
    double pmax_cutoff; /* get_pmax_cutoff, ion_init.jl:55-72                   */
    double electron_weight_fac;
} McsSpecies;

/* Per-ion tallies, PURE SUMS (the reference's 1e-99 floors are added by the caller).
 * Any pointer may be NULL to skip that output. Layouts are the reference's. */
typedef struct McsTallies {
    double* pxx_flux;      /* [n_grid]                                   all_flux.jl:230 */
    double* pxz_flux;      /* [n_grid]                                   all_flux.jl:231 */
    double* energy_flux;   /* [n_grid]                                   all_flux.jl:232 */
    double* psd;           /* [(M+2)*(T+2)*n_grid] column-major           all_flux.jl:236 */
    int64_t* num_crossings;/* [n_grid]                                   all_flux.jl:254 */
    int64_t n_cr_count;    /* out: records held in the thermal log        all_flux.jl:243 */
    int64_t n_cr_overflow; /* out: records beyond na_cr (reference: scratch file, :250)    */
    int64_t* therm_grid;   /* [na_cr]  (order is NOT the reference's; consumers only bin) */
    double* therm_px_sk;   /* [na_cr] */
    double* therm_ptot_sk; /* [na_cr] */
    double* therm_weight;  /* [na_cr] */
    double* esc_psd_feb_upstream;   /* [(psd_max+1)^2]             particle_finish.jl:82 */
    double* esc_psd_feb_downstream; /* [(psd_max+1)^2]             particle_finish.jl:79 */
    double* esc_energy_eff;         /* [psd_max+1] this ion        particle_finish.jl:95 */
    double* esc_num_eff;            /* [psd_max+1] this ion        particle_finish.jl:96 */
    double* weight_coupled;         /* [na_c] this ion                    cuts.jl:156    */
    double* spectra_coupled;        /* [(psd_max+1)*na_c] this ion        cuts.jl:161    */
    double* energy_transfer_pool;   /* [n_grid] donated this ion   particle_loop.jl:681  */
    double* spectra_sf;             /* [(psd_max+1)*n_xspec]              all_flux.jl:178 */
    double* spectra_pf;             /* [(psd_max+1)*n_xspec]              all_flux.jl:185 */
    /* SURVEY 8(f1) — device-side forms of what the host consumers build from the tallies; index = jth + (T+2)*(k + (M+2)*(i-1)),
     * i.e. the reference's [jth, k, i] order restricted to the bins in use.  Filled only with cfg.bin_thermal = 1. */
    double* therm_d2N_sf;           /* thermal crossings binned in the shock frame     particle_counter.jl:426-445 */
    double* therm_d2N_pf;           /* ... boosted to the local plasma frame first     thermo_calcs.jl:133-164     */
    double* dNdp_cr_sf;             /* [(M+2)*n_grid]: sum over angle of psd           particle_counter.jl:81-85   */
    /* scalars (out) */
    double esc_flux, px_esc_feb, energy_esc_feb;       /* particle_finish.jl:81,91,92  */
    double sum_P_downstream, sum_KE_downstream;        /* particle_loop.jl:485-486     */
    double px_esc_upstream, energy_esc_upstream;       /* all_flux.jl:155-158          */
    /* statistics (out) */
    int64_t n_helix_steps, n_retro_steps;              /* the metric: scattering steps */
    int64_t n_warn_pperp, n_warn_psd_mom, n_neg_sqrt, n_retro_capped, n_errors;
    int64_t n_fate[6];
} McsTallies;

/* The 12-field particle record of main_loops.jl:212-226 as parallel arrays; NULL = skip. */
typedef struct McsPopulation {
    double* weight;
    double* ptot_pf;
    double* pb_pf;
    double* x_cm;
    double* xn_per;
    double* prp_x_cm;
    double* acctime_sec;
    double* phi_rad;
    int64_t* grid;
    int64_t* tcut;
    uint8_t* downstream;
    uint8_t* inj;
} McsPopulation;

/* One trajectory sample (replay-mode parity): state at the end of a helix-loop pass. */
typedef struct McsTraceRec {
    double x_cm, ptot_pf, pb_pf, phi_rad, acctime_sec, prp_x_cm;
    int32_t i_grid, helix_count;
    int32_t flags;      /* bit0 downstream, bit1 inj, bit2 came from retro_time, bits 8.. i_return+1 */
    int32_t n_draws;    /* uniforms consumed so far */
} McsTraceRec;

typedef struct McsTiming {
    double transport_ms; /* sum of transport-kernel durations, CUDA events on the library stream */
    double split_ms, reduce_ms, h2d_ms, d2h_ms, comm_ms;
    double ion_loop_ms;  /* device time of whole mcs_run_ion calls (transport + split + counts + comm) */
    int64_t transport_launches, other_launches;
    int64_t local_steps, local_particles; /* this rank's scattering steps / particles entered into pcuts (before any all-reduce) */
    int64_t local_reds;  /* red.global operations this rank issued into the tallies (FP64 cells + crossing counts) */
} McsTiming;

typedef struct McsHandle McsHandle;

MCS_API const char* mcs_last_error(void);
MCS_API const char* mcs_backend(void); /* "cuda-sm_100a" or "cpu-oracle" */

/* sizeof of {McsConfig, McsSpecies, McsTallies, McsPopulation, McsTraceRec, McsTiming}: lets an FFI
 * binder (Julia struct, ctypes) assert that its mirror of the layouts matches this build. */
MCS_API int mcs_abi_sizes(int32_t out[6]);

/* Fill constants / caps with the reference's values (parameters.jl). */
MCS_API void mcs_default_config(McsConfig* cfg);

/* Allocate device state for one host thread / one GPU. */
MCS_API int mcs_create(const McsConfig* cfg, McsHandle** out);
MCS_API int mcs_destroy(McsHandle* h);

/* Multi-GPU (one process per GPU). The id is an ncclUniqueId (128 bytes) made on rank 0 by
 * mcs_comm_unique_id and carried to the other ranks by the caller (torch.distributed / MPI / file). */
MCS_API int mcs_comm_unique_id(void* id128);
MCS_API int mcs_comm_init(McsHandle* h, int rank, int nranks, const void* id128);

/* Shock profile: nine grid arrays of n_grid+2 nodes (particle_loop.jl:22,26), the per-iteration
 * eps_target[n_grid] (iter_init.jl:1-15) and the frozen energy_recv_pool[n_grid] (main_loops.jl:164). */
MCS_API int mcs_set_profile(McsHandle* h, int32_t n_grid, const double* x_grid_cm, const double* ux_sk,
                            const double* uz_sk, const double* utot, const double* gam_sf,
                            const double* gam_ef, const double* beta_ef, const double* btot,
                            const double* theta, const double* eps_target, const double* energy_recv_pool);

/* Start an ion species: zero the per-ion tallies (clear_psd!, ion_init.jl:1-16) and upload this rank's
 * shard of the initial population (assign_particle_properties_to_population!, ion_init.jl:29-53).
 * `first_global` is the 0-based global index of pop[0] (RNG counters use global indices);
 * fields left NULL take the reference's initial values (downstream=inj=false, xn_per=fine,
 * prp_x=x_grid_stop, acctime=0, tcut=1). */
MCS_API int mcs_begin_ion(McsHandle* h, int32_t i_iter, int32_t i_ion, const McsSpecies* sp, int64_t n_pts,
                          int64_t first_global, const McsPopulation* pop);

/* init_pop (initializers.jl:977-1134) in run-length form, so that the injected population is generated
 * where it is used instead of being drawn on the host and copied (0.5 GB per ion at 1e7 particles).  Particle j of the
 * momentum-ordered population (j = 0 .. n-1, n = sum of bin_count) belongs to the bin b with bin_start[b] <= j <
 * bin_start[b+1] and gets  ptot = bin_ptot[b], weight = bin_weight[b], x = x_cm, grid = grid, two uniforms (U1, U2) from
 * the Philox block with counter (0, j, i_ion << 16, i_iter) — pcut field 0, never used by the transport — and
 *   mode MCS_INJ_UPSTREAM        pb = (ptot * 2) * (U1 - 0.5)                                   initializers.jl:1006
 *   mode MCS_INJ_FASTPUSH_NONREL vx = lo + (hi - lo) * sqrt(U1);  pb = gfac * (vx - u_stop)     initializers.jl:1119-1131
 *   mode MCS_INJ_FASTPUSH_REL    bx = lo + (hi - lo) * sqrt(U1);  pb = gfac * ((bx - bu) / (1 - bx * bu) * c),  bu = u_stop / c
 *                                                                                               initializers.jl:1095-1117
 *   phi = (2 pi) * U2                                                                           ion_init.jl:51
 * with lo, hi, gfac the per-bin bin_lo / bin_hi / bin_gfac (computed by the caller, so that only IEEE sqrt / add / mul /
 * div remain per particle).  Slot s of the population holds particle j = perm(s): perm_stride = 0 is the identity;
 * perm_stride = K lays the K strided sub-sequences (0, K, 2K, ...), (1, K+1, ...), ... end to end, which makes every
 * contiguous shard a fair sample of the distribution.  The remaining fields take the reference's initial
 * values (ion_init.jl:29-53). */
enum { MCS_INJ_UPSTREAM = 0, MCS_INJ_FASTPUSH_NONREL = 1, MCS_INJ_FASTPUSH_REL = 2 };
typedef struct McsInjection {
    int32_t n_bins;
    int32_t mode;
    const double* bin_ptot;   /* [n_bins] */
    const double* bin_weight; /* [n_bins] */
    const int64_t* bin_start; /* [n_bins + 1] exclusive prefix sum of the per-bin particle counts */
    const double* bin_lo;     /* [n_bins] (fast push) */
    const double* bin_hi;     /* [n_bins] (fast push) */
    const double* bin_gfac;   /* [n_bins] (fast push) */
    double x_cm;
    double u_stop;
    int64_t grid;
    int32_t perm_stride;
    int32_t reserved;
} McsInjection;
/* mcs_begin_ion with the population generated in place: this rank holds slots first_global .. first_global + n_local - 1
 * of the n_total = bin_start[n_bins] particles. */
MCS_API int mcs_begin_ion_generate(McsHandle* h, int32_t i_iter, int32_t i_ion, const McsSpecies* sp, int64_t first_global,
                                   int64_t n_local, const McsInjection* inj);

/* == main_loops.jl:184-292 for one pcut, on device. n_saved / n_steps are this rank's. */
MCS_API int mcs_run_pcut(McsHandle* h, int32_t i_pcut, double pcut, double pcut_prev, int64_t* n_saved,
                         int64_t* n_steps);

/* == cuts.jl:34-98 on device: i_mult = max(n_pts_target / n_saved_global, 1), order-preserving clones.
 * Outputs: this rank's new population size, the global one, and i_mult. */
MCS_API int mcs_split(McsHandle* h, int64_t n_pts_target, int64_t* n_pts_new_local,
                      int64_t* n_pts_new_global, int64_t* i_mult);

/* Host-driven variant for callers that do the rank exchange themselves (MPI, gloo, Distributed.jl):
 * the caller supplies i_mult (from the GLOBAL n_saved) and the global index of this rank's first child
 * (= i_mult * number of saved particles on lower ranks). */
MCS_API int mcs_split_explicit(McsHandle* h, int64_t i_mult, int64_t first_global_child, int64_t* n_pts_new_local);

/* Whole pcut loop of one ion (main_loops.jl:179-317) without host round trips of particle data.
 * n_pcuts_run = pcuts entered; per-pcut global counts are written when the arrays are non-NULL. */
MCS_API int mcs_run_ion(McsHandle* h, const double* pcuts, int32_t n_pcuts, double p_pcut_hi,
                        int64_t n_pts_pcut, int64_t n_pts_pcut_hi, int32_t* n_pcuts_run,
                        int64_t* n_used_per_pcut, int64_t* n_saved_per_pcut);

/* Finish an ion: sum tallies over ranks (NCCL) and copy them to the caller's arrays. */
MCS_API int mcs_end_ion(McsHandle* h, McsTallies* out);

/* the pressure consumer of the tallies, thermo_calcs (/root/reference/src/thermo_calcs.jl:31-355, called from
 * ion_finalize.jl:38-47), evaluated where the PSD lives instead of shipping (M+2)(T+2)n_grid doubles to the host first.
 * Per zone: d2N_pf = 1e-99 + thermal crossings boosted to the plasma frame (McsTallies.therm_d2N_pf, :133-164) + every CR
 * cell of psd re-binned by the boost of its bin centre (:178-207); normalised to zone_pop (:209-226); then the three
 * normalisation cases and the pressure / energy-density sums (:242-352).  Needs cfg.bin_thermal = 1 and the tallies of an
 * ended ion (mcs_end_ion) unless the three input arrays below are given.  Species constants (aa, zz unused, T0, n0) are the
 * McsSpecies of the last mcs_begin_ion; beta0/gam0 and the bin geometry are the McsConfig's. */
typedef struct McsThermoIn {
    const double* cos_center; /* [T+1]  -(cos_lo + cos_hi)/2 of angle bin jth = 0..T          thermo_calcs.jl:59-75  */
    const double* pt_center;  /* [M+1]  centre momentum [g cm/s] of bin k = 0..M              thermo_calcs.jl:77-82  */
    const double* zone_pop;   /* [n_grid] particles per zone, set_grid_volumes!      particle_counter.jl:1463-1523   */
    double temperature_K;     /* T0_ion[i_ion] (McsSpecies carries no temperature)                                   */
    /* optional HOST arrays replacing the device-resident tallies (NULL = use the ion just ended); layouts of McsTallies */
    const double* psd;            /* [(M+2)(T+2)n_grid] */
    const double* therm_d2N_pf;   /* [(T+2)(M+2)n_grid] */
    const int64_t* num_crossings; /* [n_grid]           */
} McsThermoIn;
/* out arrays [n_grid] each (NULL = skip): P_psd_par, P_psd_perp, energy_density_psd as returned by thermo_calcs, and the
 * normalised zone totals d2N_pop (:225).  The CR re-binning adds FP64 cells with atomics: sums agree with the serial order
 * to rounding (~1e-15 relative), not bit for bit. */
MCS_API int mcs_thermo(McsHandle* h, const McsThermoIn* in, double* P_psd_par, double* P_psd_perp, double* energy_density_psd,
                       double* d2N_pop);

/* Inspection (tests, host-side new_pcut, replay parity). `which`: 0 = current population (*_new),
 * 1 = *_saved arrays of the last pcut (sparse, with l_save). n = number of entries to copy. */
MCS_API int mcs_get_population(McsHandle* h, int32_t which, int64_t n, McsPopulation* out, uint8_t* l_save);
MCS_API int64_t mcs_population_size(McsHandle* h);
/* Per-particle outcome of the last pcut: fate (MCS_FATE_*), helix passes, retro passes, uniforms drawn. */
MCS_API int mcs_get_fates(McsHandle* h, int64_t n, int32_t* fate, int32_t* helix_count, int64_t* retro_steps,
                          int64_t* n_draws);

/* Replay mode: particle i (local index) draws u[offsets[i]], u[offsets[i]+1], ... < offsets[i+1]. */
MCS_API int mcs_replay_set_stream(McsHandle* h, const double* u, const int64_t* offsets, int64_t n_particles);
/* Record up to max_steps McsTraceRec per listed particle during the next mcs_run_pcut. */
MCS_API int mcs_trace_enable(McsHandle* h, const int64_t* local_idx, int32_t n_trace, int32_t max_steps);
MCS_API int mcs_trace_get(McsHandle* h, McsTraceRec* recs /*[n_trace*max_steps]*/, int32_t* n_recorded /*[n_trace]*/);

MCS_API int mcs_get_timing(McsHandle* h, McsTiming* out, int32_t reset);

/* Micro-benchmarks that pin the roofline denominators on the device the handle lives on
 * (MEASURED_PEAKS.json has no FP64 entry): DFMA TFLOP/s and scattered FP64 atomicAdd G-ops/s. */
MCS_API int mcs_measure_fp64_peak(McsHandle* h, double* tflops);
MCS_API int mcs_measure_atomic_peak(McsHandle* h, int64_t n_cells, double* gops);
/* Rate [steps/s] of the bare arithmetic of one bulk scattering step (Philox + kick + phase + move)
 * with no control flow around it: the practical ceiling of the transport kernel's hot path on this device. */
MCS_API int mcs_measure_scatter_peak(McsHandle* h, double* steps_per_s);
/* Device self-test of the branch-free sqrt / division sequences of the fast loop (csrc/mcs_math.cuh): n operand pairs
 * drawn over the kernel's ranges, each compared bit for bit with the IEEE sqrt() and `/`; returns the mismatch counts. */
MCS_API int mcs_selftest_math(McsHandle* h, int64_t n, int64_t* n_bad_sqrt, int64_t* n_bad_div);

#ifdef __cplusplus
}
#endif
#endif /* MCS_H */
