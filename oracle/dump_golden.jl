# dump_golden.jl -- pin oracle/bp_oracle.c against the REAL LDPCDecoders.jl.  Needs Julia; UNEXECUTED in the build image
# (no julia binary there, no network), which is why DESIGN.md says "parity unpinned".  Anyone with Julia runs
#
#     julia --project=/path/to/LDPCDecoders.jl oracle/dump_golden.jl [outdir]
#
# from the repository root.  For every case under tests/golden/julia_twins/ (written by tests/golden/make_golden.py: the
# [[144,12,12]] gross code, the d = 15 surface code, the HGP-1600 code and the (1000,10,9) Gallager matrix; H as 1-based
# "row col" pairs, the syndromes, and the ORACLE's outputs) it
#   1. builds BeliefPropagationDecoder(H, per, max_iters) and runs batchdecode! (belief_propagation.jl:220-231), then
#      decode! per column to read the executed iteration count is not possible (the package does not expose it), so
#      errors and converged flags are compared;
#   2. builds BeliefPropagationOSDDecoder(H, per, max_iters_osd; osd_order = 0) and decodes every column
#      (belief_propagation_osd.jl:49-125) -- this exercises the real exp.(log_probabs) sort key, the one place where the
#      oracle knowingly deviates (it sorts on 1/R; see profiles/r2_osd_key_study.txt for how often that matters);
#   3. writes the package's outputs next to the oracle's and PRINTS the number of differing columns per case;
#   4. exits with status 1 if any BP column differs (BP+OSD differences are reported separately, with the indices of the
#      columns, because a small number of them is the documented sort-key deviation and not an oracle bug).
using LDPCDecoders, DelimitedFiles, SparseArrays

twins = joinpath("tests", "golden", "julia_twins")
outdir = length(ARGS) > 0 ? ARGS[1] : "golden_from_julia"
mkpath(outdir)
bp_bad_total = 0
for f in sort(filter(x -> endswith(x, ".H.coo.txt"), readdir(twins; join=true)))
    name = replace(basename(f), ".H.coo.txt" => "")
    meta = readdlm(joinpath(twins, name * ".meta.txt"))
    s, n, per, max_iters = Int(meta[1]), Int(meta[2]), Float64(meta[3]), Int(meta[4])
    ij = Int.(readdlm(f))
    H = BitMatrix(Matrix(sparse(ij[:, 1], ij[:, 2], trues(size(ij, 1)), s, n)))
    syndromes = Int.(readdlm(joinpath(twins, name * ".syndromes.txt")))
    B = size(syndromes, 2)

    dec = BeliefPropagationDecoder(H, per, max_iters)
    errors = zeros(Int, n, B)
    _, success = batchdecode!(dec, syndromes, errors)
    want_e = Int.(readdlm(joinpath(twins, name * ".errors.txt")))
    want_c = vec(Int.(readdlm(joinpath(twins, name * ".converged.txt"))))
    bad = [i for i in 1:B if errors[:, i] != want_e[:, i] || Int(success[i]) != want_c[i]]
    writedlm(joinpath(outdir, name * ".errors.txt"), errors)
    writedlm(joinpath(outdir, name * ".converged.txt"), Int.(success))
    println(name, ": BP  ", length(bad), " of ", B, " columns differ from the oracle", isempty(bad) ? "" : "  -> " * string(bad))
    global bp_bad_total += length(bad)

    osd_meta = Int.(readdlm(joinpath(twins, name * ".osd_meta.txt")))
    osd = BeliefPropagationOSDDecoder(H, per, osd_meta[1]; osd_order=0)
    osd_errors = zeros(Int, n, B)
    osd_conv = zeros(Int, B)
    for i in 1:B
        guess, conv = decode!(osd, syndromes[:, i])
        osd_errors[:, i] .= guess
        osd_conv[i] = Int(conv)
    end
    want_oe = Int.(readdlm(joinpath(twins, name * ".osd_errors.txt")))
    want_oc = vec(Int.(readdlm(joinpath(twins, name * ".osd_converged.txt"))))
    bad_osd = [i for i in 1:B if osd_errors[:, i] != want_oe[:, i] || osd_conv[i] != want_oc[i]]
    writedlm(joinpath(outdir, name * ".osd_errors.txt"), osd_errors)
    writedlm(joinpath(outdir, name * ".osd_converged.txt"), osd_conv)
    println(name, ": BP+OSD-0 (max_iters ", osd_meta[1], ", ", osd_meta[2], " columns reach OSD)  ", length(bad_osd), " of ", B,
            " columns differ from the oracle", isempty(bad_osd) ? "" : "  -> " * string(bad_osd))
end
if bp_bad_total > 0
    println("FAILED: the oracle's BP restatement disagrees with LDPCDecoders.jl on ", bp_bad_total, " columns")
    exit(1)
end
println("BP restatement agrees with LDPCDecoders.jl on every column")
