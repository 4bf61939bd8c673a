# Wall-clock timing of the reference's own CPU loop, for users who have Julia 1.12 and a checkout of
# abhro/MonteCarloScattering.jl.  UNTESTED in this repository's build image (no Julia there); the numbers in
# BASELINE.md come from the C restatement in oracle/, not from this script.
#
#   julia --project=/path/to/MonteCarloScattering.jl --threads=auto tools/time_reference.jl /path/to/run_dir
#
# run_dir must hold the mc_in.toml to time (the reference reads it from the working directory,
# src/MonteCarloScattering.jl:68).  The script reports seconds per iteration; the reference does not export a step
# counter, so steps/s is obtained by dividing the step count of the same configuration reported by
# `python bench.py` (`config.steps_per_iteration` in its JSON line) by this time.
using TOML
import MonteCarloScattering

run_dir = length(ARGS) >= 1 ? ARGS[1] : pwd()
cd(run_dir) do
    cfg = TOML.parsefile("mc_in.toml")
    n_iter = get(cfg, "num-iterations", 1)
    MonteCarloScattering.main(String[])            # first call compiles; discard
    t = @elapsed MonteCarloScattering.main(String[])
    println("{\"impl\": \"julia-reference\", \"threads\": $(Threads.nthreads()), \"iterations\": $n_iter, ",
            "\"s_total\": $t, \"s_per_iteration\": $(t / n_iter)}")
end
