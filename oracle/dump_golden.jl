# dump_golden.jl -- regenerate golden vectors from the REAL LDPCDecoders.jl (needs Julia; UNEXECUTED
# in the build image).  Writes, for every tests/golden/*.npz-equivalent case, a .txt triple the
# Python tests can be pointed at to pin oracle/bp_oracle.c against the reference itself:
#   julia --project=/path/to/LDPCDecoders.jl oracle/dump_golden.jl outdir
using LDPCDecoders, DelimitedFiles, SparseArrays
outdir = length(ARGS) > 0 ? ARGS[1] : "golden_from_julia"
mkpath(outdir)
for f in filter(x -> endswith(x, ".H.txt"), readdir("tests/golden"; join=true))
    name = replace(basename(f), ".H.txt" => "")
    H = Int.(readdlm(f))
    meta = readdlm(joinpath("tests/golden", name * ".meta.txt"))
    per, max_iters = Float64(meta[1]), Int(meta[2])
    syndromes = Int.(readdlm(joinpath("tests/golden", name * ".syndromes.txt")))
    dec = BeliefPropagationDecoder(H, per, max_iters)
    errors = zeros(Int, size(H, 2), size(syndromes, 2))
    _, success = batchdecode!(dec, syndromes, errors)
    writedlm(joinpath(outdir, name * ".errors.txt"), errors)
    writedlm(joinpath(outdir, name * ".converged.txt"), Int.(success))
end
