# belief_propagation_osd_b200.jl -- OPTIONAL addition after src/decoders/belief_propagation_osd.jl (which stays as it is).
#
# UNEXECUTED IN THE BUILD ENVIRONMENT (no julia binary in the image); mirrored by
# ldpcdecoders.jl_b200/decoder.py: BeliefPropagationOSDDecoder / batchdecode_b.
#
# With julia/belief_propagation_b200.jl in place, BeliefPropagationOSDDecoder already runs its BP stage on the GPU for every
# osd_order: decode!(::BeliefPropagationOSDDecoder, syndrome) (belief_propagation_osd.jl:49-61) calls
# decode!(decoder.bp_decoder, syndrome) and reads decoder.bp_decoder.scratch.log_probabs.  This file only adds a faster
# batch method for osd_order = 0 (BASELINE config 4): BP and the OSD-0 elimination (osd(..., Val(0)), :63-125) of the
# syndromes BP left unconverged run in ONE library call, instead of the generic column loop of abstract_decoder.jl:31-42.
# Orders 1..12 take the same call with the library option "osd_order" set (osdk_kernel: the order-O search of :127-209 on
# EVERY column, as decode! does for osd_order > 0); larger orders fall through to the generic loop (BP on the GPU,
# osd(..., Val{O}) in Julia).
function batchdecode!(decoder::BeliefPropagationOSDDecoder, syndromes::AbstractMatrix, errors::AbstractMatrix,
                      converged::AbstractVector{Bool})
    if !(0 <= decoder.osd_order <= 12) || decoder.bp_decoder.variant !== :sumproduct
        return invoke(batchdecode!, Tuple{AbstractDecoder,AbstractMatrix,AbstractMatrix,AbstractVector{Bool}},
                      decoder, syndromes, errors, converged)
    end
    @assert size(syndromes, 2) == size(errors, 2)
    @assert size(syndromes, 2) == length(converged)
    bp = decoder.bp_decoder
    B = size(syndromes, 2)
    syn = syndromes isa _B200_IN ? syndromes : Matrix{Int64}(syndromes)
    err = errors isa _B200_OUT ? errors : Matrix{Int64}(undef, bp.n, B)
    conv = converged isa Vector{Bool} ? converged : Vector{Bool}(undef, B)
    _b200_check(ccall((:ldpcb200_set_option, LDPCB200_LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), bp.handle, "osd_order", decoder.osd_order))
    GC.@preserve syn err conv begin
        _b200_check(ccall((:ldpcb200_bposd_decode_batch, LDPCB200_LIB), Cint,
                          (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
                          bp.handle, B, _b200_ptr(syn), _b200_fmt(syn), bp.s, _b200_ptr(err), _b200_fmt(err), bp.n,
                          pointer(conv), C_NULL, C_NULL, C_NULL))
    end
    err === errors || (errors .= err)
    conv === converged || (converged .= conv)      # BP's converged flag, as :60 returns
    return errors, converged
end
