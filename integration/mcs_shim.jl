# mcs_shim.jl — Julia side of the drop-in boundary (include/mcs.h) for MonteCarloScattering.jl.
#
# What it replaces in the reference (all line numbers: /root/reference/src/):
#   main_loops.jl:164        energy_recv_pool .= energy_transfer_pool            -> mcs_set_profile(..., recv_pool)
#   main_loops.jl:179-317    the whole pcut loop of one ion: l_save/zero! resets, the `for i_prt` loop calling
#                            particle_loop (particle_loop.jl:1-508) and particle_finish! (particle_finish.jl:46-107),
#                            pcut_finalize (cuts.jl:100-124) and new_pcut (cuts.jl:34-98)   -> mcs_run_ion
#   ion_init.jl:1-16         clear_psd!                                            -> done by mcs_begin_ion
# Everything else of main_loops (iteration / ion loops, init_pop, ion_finalize, smoothing) stays as it is.
#
# Usage inside main_loops.jl (sketch of the patched loop_ion body):
#
#     include("integration/mcs_shim.jl"); using .McsShim
#     h = McsShim.create(cfg)                                  # once, after setup_profile
#     ...
#     for i_ion in 1:n_ions
#         ... init_pop, assign_particle_properties_to_population! as before ...
#         McsShim.run_ion_gpu!(h, i_iter, i_ion; <the arrays main_loops already holds>)
#         ... ion_finalize as before ...
#     end
#
# No Julia runtime exists in the image this repository is built in: this file is checked there only structurally
# (tests/test_julia_shim.py parses the struct definitions below and compares names, order, sizes and offsets with the
# ctypes mirror that the GPU tests run through).  `McsShim.check_abi()` repeats that check against the loaded library.
module McsShim

export McsConfig, McsSpecies, McsPopulation, McsInjection, McsTallies, McsTraceRec, McsTiming
export create, destroy, check_abi, default_config, set_profile!, begin_ion!, run_ion!, end_ion!, run_ion_gpu!, thermo_gpu, f64ptr
export McsThermoIn

const LIBMCS = get(ENV, "MCS_LIB", joinpath(@__DIR__, "..", "montecarloscattering.jl_b200", "libmcs_b200.so"))

const MCS_ABI_VERSION = Int32(1)
const MCS_NA_C = 100
const MCS_PSD_MAX = 200
const MCS_MAX_IONS = 8
const MCS_MAX_XSPEC = 16
const MCS_RNG_PHILOX = Int32(0)
const MCS_RNG_REPLAY = Int32(1)

# ---- layouts: field for field include/mcs.h -------------------------------------------------------------------------

struct McsConfig
    abi_version::Int32
    device::Int32
    mp_g::Float64
    c_cms::Float64
    qcgs_esu::Float64
    E_rel_pt::Float64
    rad_loss_fac::Float64
    gam0::Float64
    beta0::Float64
    u0::Float64
    u2::Float64
    bmag2::Float64
    pe_crit::Float64
    gam_e_crit::Float64
    eta_mfp::Float64
    psd_mom_min::Float64
    psd_cos_fine::Float64
    delta_cos::Float64
    psd_theta_min::Float64
    psd_bins_per_dec_mom::Int32
    psd_bins_per_dec_theta::Int32
    num_psd_mom_bins::Int32
    num_psd_theta_bins::Int32
    energy_transfer_frac::Float64
    feb_upstream::Float64
    feb_downstream::Float64
    x_grid_stop::Float64
    B_CMBz::Float64
    xn_per_fine::Float64
    xn_per_coarse::Float64
    age_max::Float64
    n_grid::Int32
    i_grid_feb::Int32
    i_shock::Int32
    n_ions::Int32
    n_pts_max::Int64
    na_cr::Int64
    n_xspec::Int32
    x_spec::NTuple{16, Float64}
    n_tcuts::Int32
    tcuts::NTuple{100, Float64}
    inj_fracs::NTuple{8, Float64}
    do_rad_losses::Int32
    do_retro::Int32
    do_tcuts::Int32
    dont_DSA::Int32
    dont_scatter::Int32
    use_custom_frg::Int32
    use_custom_epsB::Int32
    helix_cap::Int32
    retro_cap::Int64
    seed::UInt64
    compat::UInt32
    rng_mode::Int32
    threads::Int32
    bin_thermal::Int32
    dynamic_queue::Int32
    det_tallies::Int32
end

struct McsSpecies
    aa::Float64
    zz_esu::Float64
    n0::Float64
    pmax_cutoff::Float64
    electron_weight_fac::Float64
end

struct McsTallies
    pxx_flux::Ptr{Float64}
    pxz_flux::Ptr{Float64}
    energy_flux::Ptr{Float64}
    psd::Ptr{Float64}
    num_crossings::Ptr{Int64}
    n_cr_count::Int64
    n_cr_overflow::Int64
    therm_grid::Ptr{Int64}
    therm_px_sk::Ptr{Float64}
    therm_ptot_sk::Ptr{Float64}
    therm_weight::Ptr{Float64}
    esc_psd_feb_upstream::Ptr{Float64}
    esc_psd_feb_downstream::Ptr{Float64}
    esc_energy_eff::Ptr{Float64}
    esc_num_eff::Ptr{Float64}
    weight_coupled::Ptr{Float64}
    spectra_coupled::Ptr{Float64}
    energy_transfer_pool::Ptr{Float64}
    spectra_sf::Ptr{Float64}
    spectra_pf::Ptr{Float64}
    therm_d2N_sf::Ptr{Float64}
    therm_d2N_pf::Ptr{Float64}
    dNdp_cr_sf::Ptr{Float64}
    esc_flux::Float64
    px_esc_feb::Float64
    energy_esc_feb::Float64
    sum_P_downstream::Float64
    sum_KE_downstream::Float64
    px_esc_upstream::Float64
    energy_esc_upstream::Float64
    n_helix_steps::Int64
    n_retro_steps::Int64
    n_warn_pperp::Int64
    n_warn_psd_mom::Int64
    n_neg_sqrt::Int64
    n_retro_capped::Int64
    n_errors::Int64
    n_fate::NTuple{6, Int64}
end

struct McsPopulation
    weight::Ptr{Float64}
    ptot_pf::Ptr{Float64}
    pb_pf::Ptr{Float64}
    x_cm::Ptr{Float64}
    xn_per::Ptr{Float64}
    prp_x_cm::Ptr{Float64}
    acctime_sec::Ptr{Float64}
    phi_rad::Ptr{Float64}
    grid::Ptr{Int64}
    tcut::Ptr{Int64}
    downstream::Ptr{UInt8}
    inj::Ptr{UInt8}
end

struct McsInjection
    n_bins::Int32
    mode::Int32
    bin_ptot::Ptr{Float64}
    bin_weight::Ptr{Float64}
    bin_start::Ptr{Int64}
    bin_lo::Ptr{Float64}
    bin_hi::Ptr{Float64}
    bin_gfac::Ptr{Float64}
    x_cm::Float64
    u_stop::Float64
    grid::Int64
    perm_stride::Int32
    reserved::Int32
end

struct McsTraceRec
    x_cm::Float64
    ptot_pf::Float64
    pb_pf::Float64
    phi_rad::Float64
    acctime_sec::Float64
    prp_x_cm::Float64
    i_grid::Int32
    helix_count::Int32
    flags::Int32
    n_draws::Int32
end

struct McsTiming
    transport_ms::Float64
    split_ms::Float64
    reduce_ms::Float64
    h2d_ms::Float64
    d2h_ms::Float64
    comm_ms::Float64
    ion_loop_ms::Float64
    transport_launches::Int64
    other_launches::Int64
    local_steps::Int64
    local_particles::Int64
    local_reds::Int64
end

struct McsThermoIn
    cos_center::Ptr{Float64}
    pt_center::Ptr{Float64}
    zone_pop::Ptr{Float64}
    temperature_K::Float64
    psd::Ptr{Float64}
    therm_d2N_pf::Ptr{Float64}
    num_crossings::Ptr{Int64}
end

# ---- plumbing ---------------------------------------------------------------------------------------------------------

last_error() = unsafe_string(ccall((:mcs_last_error, LIBMCS), Cstring, ()))
check(rc::Integer) = rc == 0 ? nothing : error("libmcs: ", last_error(), " (code ", rc, ")")

"sizeof of the six boundary structs as this build of the library sees them; must equal Julia's."
function check_abi()
    out = zeros(Int32, 6)
    check(ccall((:mcs_abi_sizes, LIBMCS), Cint, (Ptr{Int32},), out))
    mine = Int32[sizeof(McsConfig), sizeof(McsSpecies), sizeof(McsTallies), sizeof(McsPopulation), sizeof(McsTraceRec), sizeof(McsTiming)]
    out == mine || error("libmcs ABI mismatch: library ", out, " vs Julia ", mine)
    return true
end

"McsConfig with the reference's constants and caps (parameters.jl, constants.jl:30) filled in by the library."
function default_config()
    r = Ref{McsConfig}()
    ccall((:mcs_default_config, LIBMCS), Cvoid, (Ref{McsConfig},), r)
    return r[]
end

"Copy of `c` with the named fields replaced (McsConfig is immutable so that its layout is C's)."
function with(c::T; kw...) where {T}
    vals = map(fieldnames(T)) do f
        haskey(kw, f) ? convert(fieldtype(T, f), kw[f]) : getfield(c, f)
    end
    return T(vals...)
end

ntuple_pad(v, n) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, n)

"Pointer to the Float64 payload of a vector of Unitful quantities (cgstypes.jl:8-21: same bits) or of an OffsetArray."
f64ptr(v::AbstractArray) = Ptr{Float64}(pointer(parent(v)))
i64ptr(v::AbstractArray) = Ptr{Int64}(pointer(parent(v)))
u8ptr(v::AbstractArray) = Ptr{UInt8}(pointer(parent(v)))

"""
    config_from_main(; kw...) -> McsConfig

The scalars of particle_loop's argument list (particle_loop.jl:1-31) as one struct.  Every keyword is the bare Float64 /
Int value in cgs (`ustrip` of the Unitful quantity main() holds): γ₀ β₀ u₀ u₂ bmag₂ pₑ_crit γₑ_crit η_mfp, the PSD scalars
of MonteCarloScattering.jl:284-332, energy_transfer_frac, feb_upstream feb_downstream x_grid_stop B_CMBz, xn_per_fine
xn_per_coarse age_max, n_grid i_grid_feb i_shock n_ions, n_pts_max, x_spec tcuts inj_fracs and the flags.
"""
function config_from_main(; γ₀, β₀, u₀, u₂, bmag₂, pₑ_crit, γₑ_crit, η_mfp,
        psd_mom_min, psd_cos_fine, Δcos, psd_θ_min, psd_bins_per_dec_mom, psd_bins_per_dec_θ, num_psd_mom_bins, num_psd_θ_bins,
        energy_transfer_frac, feb_upstream, feb_downstream, x_grid_stop, B_CMBz, xn_per_fine, xn_per_coarse, age_max,
        n_grid, i_grid_feb, i_shock, n_ions, n_pts_max, x_spec = Float64[], tcuts = Float64[], inj_fracs = ones(n_ions),
        do_rad_losses = false, do_retro = true, do_tcuts = false, dont_DSA = false, dont_scatter = false,
        use_custom_frg = false, use_custom_εB = false, seed = 210, na_cr = 1_000_000, device = -1, det_tallies = 1)
    c = default_config()
    return with(c; device = device, gam0 = γ₀, beta0 = β₀, u0 = u₀, u2 = u₂, bmag2 = bmag₂, pe_crit = pₑ_crit,
        gam_e_crit = γₑ_crit, eta_mfp = η_mfp, psd_mom_min = psd_mom_min, psd_cos_fine = psd_cos_fine, delta_cos = Δcos,
        psd_theta_min = psd_θ_min, psd_bins_per_dec_mom = psd_bins_per_dec_mom, psd_bins_per_dec_theta = psd_bins_per_dec_θ,
        num_psd_mom_bins = num_psd_mom_bins, num_psd_theta_bins = num_psd_θ_bins, energy_transfer_frac = energy_transfer_frac,
        feb_upstream = feb_upstream, feb_downstream = feb_downstream, x_grid_stop = x_grid_stop, B_CMBz = B_CMBz,
        xn_per_fine = xn_per_fine, xn_per_coarse = xn_per_coarse, age_max = age_max, n_grid = n_grid, i_grid_feb = i_grid_feb,
        i_shock = i_shock, n_ions = n_ions, n_pts_max = n_pts_max, na_cr = na_cr,
        n_xspec = length(x_spec), x_spec = ntuple_pad(x_spec, MCS_MAX_XSPEC),
        n_tcuts = length(tcuts), tcuts = ntuple_pad(tcuts, MCS_NA_C), inj_fracs = ntuple_pad(inj_fracs, MCS_MAX_IONS),
        do_rad_losses = do_rad_losses, do_retro = do_retro, do_tcuts = do_tcuts, dont_DSA = dont_DSA, dont_scatter = dont_scatter,
        use_custom_frg = use_custom_frg, use_custom_epsB = use_custom_εB, seed = seed, det_tallies = det_tallies)
end

"Allocate device state for this process' GPU (mcs_create)."
function create(cfg::McsConfig)
    check_abi()
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mcs_create, LIBMCS), Cint, (Ref{McsConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    return h[]
end
destroy(h) = check(ccall((:mcs_destroy, LIBMCS), Cint, (Ptr{Cvoid},), h))

"Shock profile of this iteration (nine `0:n_grid+1` OffsetVectors), ε_target and the frozen energy pool (main_loops.jl:164)."
function set_profile!(h, n_grid, x_grid_cm, uₓ_sk_grid, uz_sk_grid, utot_grid, γ_sf_grid, γ_ef_grid, β_ef_grid, btot_grid, θ_grid,
        ε_target, energy_recv_pool)
    check(ccall((:mcs_set_profile, LIBMCS), Cint,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        h, n_grid, f64ptr(x_grid_cm), f64ptr(uₓ_sk_grid), f64ptr(uz_sk_grid), f64ptr(utot_grid), f64ptr(γ_sf_grid),
        f64ptr(γ_ef_grid), f64ptr(β_ef_grid), f64ptr(btot_grid), f64ptr(θ_grid), f64ptr(ε_target), f64ptr(energy_recv_pool)))
end

"Upload the population assign_particle_properties_to_population! just filled (ion_init.jl:29-53); resets the per-ion tallies."
function begin_ion!(h, i_iter, i_ion, sp::McsSpecies, n_pts_use, weight_new, ptot_pf_new, pb_pf_new, x_PT_cm_new, grid_new, φ_rad_new;
        first_global = 0)
    pop = McsPopulation(f64ptr(weight_new), f64ptr(ptot_pf_new), f64ptr(pb_pf_new), f64ptr(x_PT_cm_new), C_NULL, C_NULL, C_NULL,
        f64ptr(φ_rad_new), i64ptr(grid_new), C_NULL, C_NULL, C_NULL)
    GC.@preserve weight_new ptot_pf_new pb_pf_new x_PT_cm_new grid_new φ_rad_new begin
        check(ccall((:mcs_begin_ion, LIBMCS), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{McsSpecies}, Int64, Int64, Ref{McsPopulation}),
            h, i_iter, i_ion, sp, n_pts_use, first_global, pop))
    end
end

"The whole pcut loop of one ion on the device (main_loops.jl:179-317)."
function run_ion!(h, pcuts_cgs::Vector{Float64}, p_pcut_hi, n_pts_pcut, n_pts_pcut_hi)
    n_run = Ref{Int32}(0)
    used = zeros(Int64, length(pcuts_cgs))
    saved = zeros(Int64, length(pcuts_cgs))
    check(ccall((:mcs_run_ion, LIBMCS), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Int64, Int64, Ref{Int32}, Ptr{Int64}, Ptr{Int64}),
        h, pcuts_cgs, length(pcuts_cgs), p_pcut_hi, n_pts_pcut, n_pts_pcut_hi, n_run, used, saved))
    return n_run[], used[1:n_run[]], saved[1:n_run[]]
end

"Scratch arrays with the reference's shapes (SURVEY App. A) that mcs_end_ion fills with PURE SUMS."
struct TallyScratch
    pxx::Vector{Float64}; pxz::Vector{Float64}; en::Vector{Float64}
    psd::Vector{Float64}; ncross::Vector{Int64}
    tg::Vector{Int64}; tpx::Vector{Float64}; tpt::Vector{Float64}; tw::Vector{Float64}
    esc_up::Vector{Float64}; esc_dn::Vector{Float64}; en_eff::Vector{Float64}; num_eff::Vector{Float64}
    wc::Vector{Float64}; sc::Vector{Float64}; pool::Vector{Float64}; sf::Vector{Float64}; pf::Vector{Float64}
end
function TallyScratch(cfg::McsConfig)
    ng = Int(cfg.n_grid); e1 = MCS_PSD_MAX + 1
    npsd = (cfg.num_psd_mom_bins + 2) * (cfg.num_psd_theta_bins + 2) * ng
    z(n) = zeros(Float64, n)
    TallyScratch(z(ng), z(ng), z(ng), z(npsd), zeros(Int64, ng), zeros(Int64, cfg.na_cr), z(cfg.na_cr), z(cfg.na_cr), z(cfg.na_cr),
        z(e1 * e1), z(e1 * e1), z(e1), z(e1), z(MCS_NA_C), z(e1 * MCS_NA_C), z(ng), z(e1 * MCS_MAX_XSPEC), z(e1 * MCS_MAX_XSPEC))
end

"Finish the ion: all-reduce over GPUs inside the library, copy the sums into `s`; returns the scalar part."
function end_ion!(h, s::TallyScratch)
    t = Ref(McsTallies(pointer(s.pxx), pointer(s.pxz), pointer(s.en), pointer(s.psd), pointer(s.ncross), 0, 0,
        pointer(s.tg), pointer(s.tpx), pointer(s.tpt), pointer(s.tw), pointer(s.esc_up), pointer(s.esc_dn), pointer(s.en_eff),
        pointer(s.num_eff), pointer(s.wc), pointer(s.sc), pointer(s.pool), pointer(s.sf), pointer(s.pf), C_NULL, C_NULL, C_NULL,
        0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0, 0, 0, 0, 0, 0, 0, ntuple(_ -> Int64(0), 6)))
    GC.@preserve s check(ccall((:mcs_end_ion, LIBMCS), Cint, (Ptr{Cvoid}, Ref{McsTallies}), h, t))
    return t[]
end

"""
    thermo_gpu(h, n_grid, cos_center, pt_center, zone_pop, T₀) -> (P_psd_par, P_psd_perp, energy_density_psd)

Drop-in for the `thermo_calcs(...)` call of ion_finalize.jl:38-47, after `end_ion!` of the same ion (needs
`bin_thermal = 1` in the McsConfig: the thermal crossings are then binned as they happen and the crossing log /
scratch file of thermo_calcs.jl:96-164 is not read).  `cos_center` (0:num_psd_θ_bins) and `pt_center`
(0:num_psd_mom_bins, in g cm/s) are the two arrays thermo_calcs.jl:55-82 builds, `zone_pop` is the one
get_normalized_dNdp returns (ion_finalize.jl:25); pass their parents / ustrip'ed payloads.  Results are plain
Float64 in cgs (dyn/cm², erg/cm³).
"""
function thermo_gpu(h, n_grid, cos_center::Vector{Float64}, pt_center::Vector{Float64}, zone_pop::Vector{Float64}, T₀::Float64)
    P_par, P_perp, e_dens = zeros(n_grid), zeros(n_grid), zeros(n_grid)
    tin = Ref(McsThermoIn(pointer(cos_center), pointer(pt_center), pointer(zone_pop), T₀, C_NULL, C_NULL, C_NULL))
    GC.@preserve cos_center pt_center zone_pop P_par P_perp e_dens check(ccall((:mcs_thermo, LIBMCS), Cint,
        (Ptr{Cvoid}, Ref{McsThermoIn}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        h, tin, P_par, P_perp, e_dens, C_NULL))
    return P_par, P_perp, e_dens
end

"""
    run_ion_gpu!(h, cfg, scratch, i_iter, i_ion; ...) -> (n_pts_use_last, ∑P_downstream, ∑KEdensity_downstream)

Drop-in for main_loops.jl:164-317 of ONE ion.  Arguments are the arrays main_loops already holds; Unitful arrays are passed
as they are (their payload is Float64).  On return the reference's tallies have been incremented exactly where
particle_loop / particle_finish! / tcut_track! would have incremented them:
  pxx_flux pxz_flux energy_flux (all_flux.jl:230-232), psd (:236), num_crossings + thermal log (:242-254),
  esc_psd_feb_* esc_flux pₓ_esc_feb energy_esc_feb esc_energy_eff esc_num_eff (particle_finish.jl:79-96),
  weight_coupled spectra_coupled (cuts.jl:156-161), energy_transfer_pool (particle_loop.jl:681), spectra_sf/pf (all_flux.jl:178-185).
"""
function run_ion_gpu!(h, cfg::McsConfig, s::TallyScratch, i_iter, i_ion;
        species_aa, species_zz_esu, species_n0, pmax_cutoff, electron_weight_fac,
        x_grid_cm, uₓ_sk_grid, uz_sk_grid, utot_grid, γ_sf_grid, γ_ef_grid, β_ef_grid, btot_grid, θ_grid,
        ε_target, energy_transfer_pool, energy_recv_pool,
        n_pts_use, weight_new, ptot_pf_new, pb_pf_new, x_PT_cm_new, grid_new, φ_rad_new,
        pcuts, p_pcut_hi, n_pts_pcut, n_pts_pcut_hi,
        pxx_flux, pxz_flux, energy_flux, psd, num_crossings, therm_grid, therm_pₓ_sk, therm_ptot_sk, therm_weight,
        esc_psd_feb_upstream, esc_psd_feb_downstream, esc_energy_eff, esc_num_eff, esc_flux, pₓ_esc_feb, energy_esc_feb,
        weight_coupled, spectra_coupled, spectra_sf, spectra_pf, ∑P_downstream, ∑KEdensity_downstream)
    ng = Int(cfg.n_grid)
    parent(energy_recv_pool) .= parent(energy_transfer_pool)                       # main_loops.jl:164
    set_profile!(h, ng, x_grid_cm, uₓ_sk_grid, uz_sk_grid, utot_grid, γ_sf_grid, γ_ef_grid, β_ef_grid, btot_grid, θ_grid,
        ε_target, energy_recv_pool)
    sp = McsSpecies(species_aa, abs(species_zz_esu), species_n0, pmax_cutoff, isfinite(electron_weight_fac) ? electron_weight_fac : 0.0)
    begin_ion!(h, i_iter, i_ion, sp, n_pts_use, weight_new, ptot_pf_new, pb_pf_new, x_PT_cm_new, grid_new, φ_rad_new)
    n_run, used, saved = run_ion!(h, collect(reinterpret(Float64, parent(pcuts))), Float64(p_pcut_hi), n_pts_pcut, n_pts_pcut_hi)
    t = end_ion!(h, s)
    # the library returns pure sums: add them where the reference accumulates (its arrays hold the 1e-99 floors already)
    f(v) = reinterpret(Float64, parent(v))
    f(pxx_flux) .+= s.pxx; f(pxz_flux) .+= s.pxz; f(energy_flux) .+= s.en
    vec(f(psd)) .+= s.psd                                                            # (0:M+1, 0:T+1, 1:n_grid), column-major
    parent(num_crossings) .+= s.ncross
    n_cr = Int(t.n_cr_count)                                                         # thermal log: order differs, consumers only bin
    parent(therm_grid)[1:n_cr] .= s.tg[1:n_cr]; f(therm_pₓ_sk)[1:n_cr] .= s.tpx[1:n_cr]
    f(therm_ptot_sk)[1:n_cr] .= s.tpt[1:n_cr]; f(therm_weight)[1:n_cr] .= s.tw[1:n_cr]
    vec(f(esc_psd_feb_upstream)) .+= s.esc_up; vec(f(esc_psd_feb_downstream)) .+= s.esc_dn
    e1 = MCS_PSD_MAX + 1
    view(f(esc_energy_eff), :, i_ion) .+= s.en_eff; view(f(esc_num_eff), :, i_ion) .+= s.num_eff
    view(f(weight_coupled), :, i_ion) .+= s.wc
    vec(view(f(spectra_coupled), :, :, i_ion)) .+= s.sc
    f(energy_transfer_pool) .+= s.pool
    nx = Int(cfg.n_xspec)
    vec(f(spectra_sf))[1:(e1 * nx)] .+= s.sf[1:(e1 * nx)]; vec(f(spectra_pf))[1:(e1 * nx)] .+= s.pf[1:(e1 * nx)]
    f(esc_flux)[i_ion] += t.esc_flux
    f(pₓ_esc_feb)[i_ion, i_iter] += t.px_esc_feb; f(energy_esc_feb)[i_ion, i_iter] += t.energy_esc_feb
    t.n_errors == 0 || error("libmcs: ", t.n_errors, " particles hit a condition on which the reference throws (all_flux.jl:73-75, prob_return.jl:134)")
    n_last = isempty(used) ? n_pts_use : used[end]
    return n_last, ∑P_downstream + t.sum_P_downstream, ∑KEdensity_downstream + t.sum_KE_downstream, n_cr
end

end # module
