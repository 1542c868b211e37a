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
    # exchange terms: no reference test pins a model with J != 0; these two settle the sign convention of HF:445-450 /
    # 563-616 that this repo could only take from the second-quantised operators (hubbardfunctions.mb_terms docstring):
    # compare with tests/test_gpu_groundstate.py::test_polyacetylene_model_with_exchange_both_symmetries
    "OB_J"  => () -> hf.OB_Sim([1.0], [4.0], 0.0, [0.5], 1, 1, 2.0),
    "C4_polyacetylene" => () -> hf.MB_Sim([0.000 3.803 -0.548 0.000; 3.803 0.000 2.977 -0.501],
                                          [10.317 6.264 0.000 0.000; 6.264 10.317 6.162 0.000],
                                          [0.000 0.123 0.000 0.000; 0.123 0.000 0.113 0.000], 1, 1, 2.5, 20; code = "c4_polyacetylene"),
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
