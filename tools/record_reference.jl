# record_reference.jl — make the REAL reference pin the oracle and the CUDA path (for users who have Julia >= 1.12).
#
#   julia --project=/path/to/MonteCarloScattering.jl tools/record_reference.jl /path/to/MonteCarloScattering.jl OUT_DIR [N_TRACE] [MAX_PASSES]
#
# run from a directory that holds the `mc_in.toml` to record (the reference reads ./mc_in.toml).  It runs the reference's
# own main() with two of its source files patched IN MEMORY (nothing is written into the checkout):
#   src/particle_loop.jl   `rng = Random.Xoshiro(iseed_mod)` (particle_loop.jl:41) becomes a recording wrapper around the
#                          same Xoshiro, so every uniform the particle consumes is logged in order; one line before
#                          "# If particle escaped downstream" (particle_loop.jl:478) logs the state at the end of the pass;
#   src/main_loops.jl      hooks at the start of an ion, before and after the particle loop of a pcut, and at the end of the ion.
# Output: OUT_DIR/manifest.json + raw little-endian arrays (no NPZ writer in Julia's stdlib; numpy reads them with
# np.fromfile).  Copy OUT_DIR to tests/golden/julia/<name>/ and run
#   python -m pytest tests/test_reference_fixtures.py           (oracle; CPU)
#   python -m pytest tests/test_reference_fixtures.py -m gpu    (CUDA, replay mode)
# Both replay the recorded uniform streams (mcs_replay_set_stream) through every pcut, compare integer decisions exactly,
# continuous state to 1e-12 relative along the traced passes and the end-of-ion tallies to 1e-10.
#
# NOT RUN in the repository's build image (no Julia there); the consuming side is exercised with a fixture of the same
# layout written by tools/make_fixture_from_oracle.py.
using Random, Printf

const REF = abspath(ARGS[1])
const OUT = abspath(ARGS[2])
const N_TRACE = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 64      # particles per pcut whose passes are traced
const MAX_PASSES = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : 200  # passes traced per particle
mkpath(OUT)

module McsRecorder
using Random
mutable struct RecordingRNG <: Random.AbstractRNG
    inner::Random.Xoshiro
    log::Vector{Float64}
end
RecordingRNG(seed::Integer) = RecordingRNG(Random.Xoshiro(seed), Float64[])
function Random.rand(r::RecordingRNG)
    u = Random.rand(r.inner)
    push!(r.log, u)
    return u
end
const CURRENT = Ref{Any}(nothing)      # per-pcut collector
const MANIFEST = Dict{String, Any}("format" => "mcs-reference-fixture-1", "pcut_records" => Any[], "ion_records" => Any[])
const OUTDIR = Ref("")
const NTRACE = Ref(64)
const MAXP = Ref(200)

strip_units(x) = x
strip_units(x::AbstractArray) = collect(reinterpret(Float64, parent(x)))
raw(x::AbstractArray{Bool}) = collect(UInt8.(parent(x)))
raw(x::AbstractArray{<:Integer}) = collect(Int64.(parent(x)))
raw(x::AbstractArray) = eltype(x) <: AbstractFloat ? collect(Float64.(parent(x))) : collect(reinterpret(Float64, parent(x)))
scalar(x::Bool) = x
scalar(x::Integer) = Int(x)
scalar(x::AbstractFloat) = Float64(x)
scalar(x) = Float64(x.val)   # Unitful.Quantity: the cgs payload

function put!(name::String, a::AbstractArray)
    v = raw(a)
    dt = eltype(v) == UInt8 ? "u1" : (eltype(v) == Int64 ? "i8" : "f8")
    open(joinpath(OUTDIR[], name * ".bin"), "w") do io; write(io, v); end
    return Dict("file" => name * ".bin", "dtype" => dt, "shape" => collect(size(v)))
end

mutable struct PcutRec
    key::String
    n::Int
    draws::Vector{Vector{Float64}}
    trace::Vector{Vector{NTuple{8, Float64}}}
end

function ion_begin(i_iter, i_ion; scalars, arrays)
    rec = Dict{String, Any}("i_iter" => i_iter, "i_ion" => i_ion, "scalars" => Dict(string(k) => scalar(v) for (k, v) in pairs(scalars)))
    for (k, v) in pairs(arrays)
        rec[string(k)] = put!(@sprintf("it%d_ion%d_begin_%s", i_iter, i_ion, k), v)
    end
    push!(MANIFEST["ion_records"], rec)
end
function ion_end(i_iter, i_ion; scalars, arrays)
    rec = MANIFEST["ion_records"][end]
    rec["end_scalars"] = Dict(string(k) => scalar(v) for (k, v) in pairs(scalars))
    for (k, v) in pairs(arrays)
        rec["end_" * string(k)] = put!(@sprintf("it%d_ion%d_end_%s", i_iter, i_ion, k), v)
    end
end
function pcut_begin(i_iter, i_ion, i_pcut, n_pts_use; arrays)
    key = @sprintf("it%d_ion%d_pcut%d", i_iter, i_ion, i_pcut)
    rec = Dict{String, Any}("i_iter" => i_iter, "i_ion" => i_ion, "i_pcut" => i_pcut, "n_pts_use" => n_pts_use)
    for (k, v) in pairs(arrays)
        rec["new_" * string(k)] = put!(key * "_new_" * string(k), view(parent(v), 1:n_pts_use))
    end
    push!(MANIFEST["pcut_records"], rec)
    CURRENT[] = PcutRec(key, n_pts_use, [Float64[] for _ in 1:n_pts_use], [NTuple{8, Float64}[] for _ in 1:min(n_pts_use, NTRACE[])])
end
function pcut_end(; arrays)
    c = CURRENT[]::PcutRec
    rec = MANIFEST["pcut_records"][end]
    for (k, v) in pairs(arrays)
        rec["saved_" * string(k)] = put!(c.key * "_saved_" * string(k), view(parent(v), 1:c.n))
    end
    off = Int64[0]
    for d in c.draws; push!(off, off[end] + length(d)); end
    rec["draws_off"] = put!(c.key * "_draws_off", off)
    rec["draws"] = put!(c.key * "_draws", reduce(vcat, c.draws; init = Float64[]))
    toff = Int64[0]
    for t in c.trace; push!(toff, toff[end] + length(t)); end
    flat = Float64[]
    for t in c.trace, r in t; append!(flat, r); end
    rec["trace_off"] = put!(c.key * "_trace_off", toff)
    rec["trace"] = put!(c.key * "_trace", reshape(flat, 8, :))   # rows: x, ptot, pb, phi, acctime, prp_x, i_grid, n_draws
end
# called from the patched particle_loop
new_rng(iseed_mod) = RecordingRNG(iseed_mod)
function particle_done(i_prt, rng::RecordingRNG)
    c = CURRENT[]::PcutRec
    c.draws[i_prt] = rng.log
end
function pass_end(i_prt, rng::RecordingRNG, x, ptot, pb, phi, acct, prp, i_grid)
    c = CURRENT[]::PcutRec
    if i_prt <= length(c.trace) && length(c.trace[i_prt]) < MAXP[]
        push!(c.trace[i_prt], (scalar(x), scalar(ptot), scalar(pb), Float64(phi), scalar(acct), scalar(prp), Float64(i_grid), Float64(length(rng.log))))
    end
end
end # module McsRecorder

McsRecorder.OUTDIR[] = OUT
McsRecorder.NTRACE[] = N_TRACE
McsRecorder.MAXP[] = MAX_PASSES

import MonteCarloScattering
const MCS = MonteCarloScattering

function patch(text::String, anchor::String, replacement::String; where::String)
    occursin(anchor, text) || error("anchor not found in $where: ", repr(anchor), " — this recorder matches the reference as surveyed (SURVEY.md); adapt the anchors to your checkout")
    return replace(text, anchor => replacement; count = 1)
end

# ---- particle_loop.jl ------------------------------------------------------------------------------------------------
pl = read(joinpath(REF, "src", "particle_loop.jl"), String)
pl = patch(pl, "rng = Random.Xoshiro(iseed_mod)", "rng = Main.McsRecorder.new_rng(iseed_mod)"; where = "particle_loop.jl:41")
pl = patch(pl, "        # If particle escaped downstream, handle final calculations here",
    "        Main.McsRecorder.pass_end(i_prt, rng, r_PT_cm.x, ptot_pf, pb_pf, φ_rad, acctime_sec, prp_x_cm, i_grid)\n" *
    "        # If particle escaped downstream, handle final calculations here"; where = "particle_loop.jl:478")
pl = patch(pl, "    end # loop_helix\n", "    end # loop_helix\n    Main.McsRecorder.particle_done(i_prt, rng)\n"; where = "particle_loop.jl:499")
Base.include_string(MCS, pl, "particle_loop.jl (recording)")

# ---- main_loops.jl -----------------------------------------------------------------------------------------------------
ml = read(joinpath(REF, "src", "main_loops.jl"), String)
ml = patch(ml, "            weight_running = weight_in[1]\n",
    "            weight_running = weight_in[1]\n" *
    "            Main.McsRecorder.ion_begin(i_iter, i_ion;\n" *
    "                scalars = (; γ₀, β₀, u₀, u₂, bmag₂, pₑ_crit, γₑ_crit, η_mfp, psd_mom_min, psd_cos_fine, Δcos, psd_θ_min,\n" *
    "                    psd_bins_per_dec_mom, psd_bins_per_dec_θ, num_psd_mom_bins, num_psd_θ_bins, energy_transfer_frac,\n" *
    "                    feb_upstream, feb_downstream, x_grid_stop, B_CMBz, xn_per_fine, xn_per_coarse, age_max, n_grid, i_grid_feb,\n" *
    "                    i_shock, n_ions, n_pts_max, n_pts_pcut, n_pts_pcut_hi, n_xspec, n_tcuts, do_rad_losses, do_retro, do_tcuts,\n" *
    "                    dont_DSA, dont_scatter, use_custom_frg, use_custom_εB, electron_weight_fac, aa, zz, m, pmax_cutoff,\n" *
    "                    n0 = MonteCarloScattering.density(species[i_ion]), n_pts_use,\n" *
    "                    p_pcut_hi = MonteCarloScattering.pcut_hi(energy_pcut_hi, MonteCarloScattering.E_rel_pt, MonteCarloScattering.mass(species[i_ion]))),\n" *
    "                arrays = (; x_grid_cm, uₓ_sk_grid, uz_sk_grid, utot_grid, γ_sf_grid, γ_ef_grid, β_ef_grid, btot_grid, θ_grid, ε_target,\n" *
    "                    energy_transfer_pool, pcuts, tcuts, x_spec, inj_fracs, pxx_flux, pxz_flux, energy_flux))\n"; where = "main_loops.jl:160")
ml = patch(ml, "                for i_prt in 1:n_pts_use # loop_pt\n",
    "                Main.McsRecorder.pcut_begin(i_iter, i_ion, i_pcut, n_pts_use; arrays = (; weight = weight_new, ptot_pf = ptot_pf_new,\n" *
    "                    pb_pf = pb_pf_new, x_cm = x_PT_cm_new, xn_per = xn_per_new, prp_x_cm = prp_x_cm_new, acctime_sec = acctime_sec_new,\n" *
    "                    phi_rad = φ_rad_new, grid = grid_new, tcut = tcut_new, downstream = downstream_new, inj = inj_new))\n" *
    "                for i_prt in 1:n_pts_use # loop_pt\n"; where = "main_loops.jl:228")
ml = patch(ml, "                # Conclusion of particle loop\n",
    "                # Conclusion of particle loop\n" *
    "                Main.McsRecorder.pcut_end(; arrays = (; l_save, weight = weight_saved, ptot_pf = ptot_pf_saved, pb_pf = pb_pf_saved,\n" *
    "                    x_cm = x_PT_cm_saved, xn_per = xn_per_saved, prp_x_cm = prp_x_cm_saved, acctime_sec = acctime_sec_saved,\n" *
    "                    phi_rad = φ_rad_saved, grid = grid_saved, tcut = tcut_saved, downstream = downstream_saved, inj = inj_saved))\n"; where = "main_loops.jl:295")
ml = patch(ml, "            # Conclusion of pcuts loop\n",
    "            # Conclusion of pcuts loop\n" *
    "            Main.McsRecorder.ion_end(i_iter, i_ion; scalars = (; ∑P_downstream, ∑KEdensity_downstream),\n" *
    "                arrays = (; pxx_flux, pxz_flux, energy_flux, psd, num_crossings, esc_psd_feb_upstream, esc_psd_feb_downstream,\n" *
    "                    esc_energy_eff, esc_num_eff, esc_flux, pₓ_esc_feb, energy_esc_feb, weight_coupled, spectra_coupled, energy_transfer_pool))\n"; where = "main_loops.jl:317")
Base.include_string(MCS, ml, "main_loops.jl (recording)")

# ---- run ---------------------------------------------------------------------------------------------------------------
isfile("mc_in.toml") || error("run this from a directory containing the mc_in.toml to record")
Base.invokelatest(MCS.main)

# minimal JSON writer (no JSON package in the reference's dependency set)
function json(io, x::AbstractDict)
    print(io, "{"); first = true
    for (k, v) in x
        first || print(io, ","); first = false
        print(io, "\"", k, "\":"); json(io, v)
    end
    print(io, "}")
end
json(io, x::AbstractVector) = (print(io, "["); for (i, v) in enumerate(x); i > 1 && print(io, ","); json(io, v); end; print(io, "]"))
json(io, x::AbstractString) = print(io, "\"", x, "\"")
json(io, x::Bool) = print(io, x ? "true" : "false")
json(io, x::Integer) = print(io, x)
json(io, x::AbstractFloat) = isfinite(x) ? print(io, repr(Float64(x))) : print(io, "\"", x, "\"")
open(joinpath(OUT, "manifest.json"), "w") do io
    json(io, McsRecorder.MANIFEST)
end
println("recorded ", length(McsRecorder.MANIFEST["pcut_records"]), " pcuts of ", length(McsRecorder.MANIFEST["ion_records"]), " ions into ", OUT)
