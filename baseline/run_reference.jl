# Out-of-container cross-check (NOT run in the build image: no Julia, no network).
#
# Runs the UNMODIFIED reference (DaanVrancken/HubbardTN at the pinned Manifest) for the parity configs and
# dumps what the device path is compared against: energy per site, densities, full bond dimensions and the
# entanglement spectrum per sector, plus wall times.  Usage, from a checkout of the reference:
#
#     julia --project=. -t auto path/to/run_reference.jl out.json
#
# then compare with `python tools/run_config.py C1` / `bench.py` output (`groundstate` block) of this repo.
using Pkg
Pkg.instantiate()
include(joinpath(pwd(), "src", "HubbardFunctions.jl"))
import .HubbardFunctions as hf
using MPSKit, TensorKit, JSON

configs = Dict(
    "C1"   => () -> hf.OB_Sim([1.0], [8.0], 0.0, 1, 1, 2.0),                       # test/Spin.jl-like, SU(2)
    "OB_U0" => () -> hf.OB_Sim([1.0], [0.0], 0.0, 1, 1, 2.0),                      # test/OB.jl:21
    "OB_U5" => () -> hf.OB_Sim([1.0], [5.0], 0.0, 1, 1, 2.0),                      # test/OB.jl:44
    "C2"   => () -> hf.OB_Sim([1.0, 0.2], [6.0], 0.0, 1, 1, 5.0),
)

out = Dict{String,Any}()
for (name, make) in configs
    model = make()
    t = @elapsed dictionary = hf.produce_groundstate(model; force=true)
    ψ = dictionary["groundstate"]; H = dictionary["ham"]
    E = sum(real(expectation_value(ψ, H))) / length(H)
    spectra = [Dict(string(c) => collect(v) for (c, v) in pairs(MPSKit.entanglement_spectrum(ψ, i).data)) for i in 1:length(ψ)]
    out[name] = Dict("seconds" => t, "energy_per_site" => E, "delta" => dictionary["delta"],
                     "density" => hf.density_state(model), "dim_state" => hf.dim_state(ψ),
                     "entanglement_spectrum" => spectra, "threads" => Threads.nthreads())
end
open(length(ARGS) > 0 ? ARGS[1] : "reference_results.json", "w") do io
    JSON.print(io, out, 1)
end
