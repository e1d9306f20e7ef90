# belief_propagation_b200.jl -- drop-in REPLACEMENT for src/decoders/belief_propagation.jl of LDPCDecoders.jl.
#
# UNEXECUTED IN THE BUILD ENVIRONMENT (no julia binary in the image).  The Python mirror
# ldpcdecoders.jl_b200/decoder.py drives exactly the same C entry points with the same argument
# conventions and is what the test-suite runs; keep the two in step.
#
# How a maintainer uses it: in src/LDPCDecoders.jl replace
#       include("decoders/belief_propagation.jl")
# by    include("decoders/belief_propagation_b200.jl")          (this file, copied next to the old one)
# and, optionally, add include("decoders/belief_propagation_osd_b200.jl") after belief_propagation_osd.jl.
# Nothing else changes: the names, signatures, field names and error behaviour below are those of
# belief_propagation.jl (:38-67 struct + constructor, :83-91 reset!, :121-188 decode!, :220-231 batchdecode!), so
#   * user code, the package's tests and docs keep calling BeliefPropagationDecoder(H, per, max_iters), decode!,
#     batchdecode! (3- and 4-argument forms: abstract_decoder.jl:44-48 forwards to the method defined here);
#   * BeliefPropagationOSDDecoder (belief_propagation_osd.jl:17-61) works unchanged on top: it calls decode! on its
#     bp_decoder and reads bp_decoder.scratch.log_probabs, both of which exist here -- so the BP stage of BP+OSD runs
#     on the same kernels for every osd_order.
# What is different: the dense s x n message matrices of the scratch space are gone (messages live on the GPU), the
# struct is mutable (it owns a library handle with a finalizer), and two keyword arguments exist that default to the
# reference behaviour:  devices (GPU ordinals the batch is sharded over)  and  variant (:sumproduct = the bit-exact
# replica of this file's original arithmetic, :minsum, :fast32).
using SparseArrays

const LDPCB200_LIB = get(ENV, "LDPCB200_LIB", "libldpcb200.so")
const _FMT_U8, _FMT_I64, _FMT_BITS, _FMT_PACKED32, _FMT_F64 = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const _B200_VARIANTS = Dict(:sumproduct => Int32(0), :minsum => Int32(1), :fast32 => Int32(2))

struct LDPCB200Error <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::LDPCB200Error) = print(io, "libldpcb200 error ", e.code, ": ", e.msg)

function _b200_check(rc::Integer)
    rc == 0 && return nothing
    throw(LDPCB200Error(Int32(rc), unsafe_string(ccall((:ldpcb200_last_error, LDPCB200_LIB), Cstring, ()))))
end

"Host-side scratch: what callers read after `decode!` (belief_propagation_osd.jl:51-52 reads `log_probabs`)."
struct BeliefPropagationScratchSpace
    log_probabs::Vector{Float64}
    channel_probs::Vector{Float64}
    err::Vector{Float64}
end
BeliefPropagationScratchSpace(n, s, per) = BeliefPropagationScratchSpace(zeros(n), fill(per, n), zeros(n))

"""
    BeliefPropagationDecoder(H, per::Float64, max_iters::Int; devices=Int32[0], variant=:sumproduct)

Sum-product belief propagation decoder; the Tanner graph of `H` and all messages live on the listed B200 GPUs behind
`libldpcb200.so`.  Fields `per, max_iters, s, n, sparse_H, sparse_HT, scratch` as before.
"""
mutable struct BeliefPropagationDecoder <: AbstractDecoder
    per::Float64
    max_iters::Int
    s::Int
    n::Int
    sparse_H::SparseArrays.SparseMatrixCSC{Bool,Int}
    sparse_HT::SparseArrays.SparseMatrixCSC{Bool,Int}
    scratch::BeliefPropagationScratchSpace
    handle::Ptr{Cvoid}
    variant::Symbol
end

function BeliefPropagationDecoder(H, per::Float64, max_iters::Int; devices=Int32[0], variant::Symbol=:sumproduct)
    s, n = size(H)
    sparse_H = SparseArrays.SparseMatrixCSC{Bool,Int}(dropzeros(sparse(H)))
    sparse_HT = SparseArrays.SparseMatrixCSC{Bool,Int}(sparse(sparse_H'))
    colptr = Vector{Int64}(sparse_H.colptr)          # n+1 entries, 1-based
    rowval = Vector{Int64}(sparse_H.rowval)          # ascending inside a column (SparseMatrixCSC invariant)
    devs = Vector{Int32}(devices)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve colptr rowval devs begin
        _b200_check(ccall((:ldpcb200_create, LDPCB200_LIB), Cint,
                          (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Float64, Int32, Int32, Ptr{Int32}, Int32, Ptr{Ptr{Cvoid}}),
                          s, n, colptr, rowval, Int32(1), per, Int32(max_iters), _B200_VARIANTS[variant], devs, Int32(length(devs)), out))
    end
    dec = BeliefPropagationDecoder(per, max_iters, s, n, sparse_H, sparse_HT, BeliefPropagationScratchSpace(n, s, per), out[], variant)
    finalizer(dec) do d
        d.handle == C_NULL || ccall((:ldpcb200_destroy, LDPCB200_LIB), Cint, (Ptr{Cvoid},), d.handle)
        d.handle = C_NULL
    end
    return dec
end

"Messages are (re)initialised inside the kernels on every decode; only the host scratch is cleared."
function reset!(bp_decoder::BeliefPropagationDecoder)
    bp_decoder.scratch.log_probabs .= 0.0
    bp_decoder.scratch.channel_probs .= bp_decoder.per
    bp_decoder.scratch.err .= 0.0
    bp_decoder
end

# element format of a Julia array at the C boundary (anything else is converted to Matrix{Int64} first)
_b200_fmt(::BitMatrix) = _FMT_BITS
_b200_fmt(::Matrix{Bool}) = _FMT_U8
_b200_fmt(::Matrix{UInt8}) = _FMT_U8
_b200_fmt(::Matrix{Int64}) = _FMT_I64
_b200_fmt(::Matrix{Float64}) = _FMT_F64
_b200_ptr(A::BitMatrix) = Ptr{Cvoid}(pointer(A.chunks))
_b200_ptr(A::Matrix) = Ptr{Cvoid}(pointer(A))
const _B200_IN = Union{BitMatrix,Matrix{Bool},Matrix{UInt8},Matrix{Int64}}
const _B200_OUT = Union{BitMatrix,Matrix{Bool},Matrix{UInt8},Matrix{Int64},Matrix{Float64}}

"""
    decode!(decoder::BeliefPropagationDecoder, syndrome::AbstractVector) -> (decoder.scratch.err, converged)

One syndrome; returns the aliased Float64 0/1 vector `scratch.err` and refreshes `scratch.log_probabs`, as before.
"""
function decode!(decoder::BeliefPropagationDecoder, syndrome::AbstractVector)
    reset!(decoder)
    syn = reshape(Vector{Int64}(syndrome), decoder.s, 1)
    err = reshape(decoder.scratch.err, decoder.n, 1)
    conv = Vector{Bool}(undef, 1)
    ratio = ones(Float64, decoder.n)
    GC.@preserve syn err conv ratio begin
        _b200_check(ccall((:ldpcb200_decode_batch, LDPCB200_LIB), Cint,
                          (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Ptr{Int32}, Ptr{Float64}, Ptr{Int64}),
                          decoder.handle, 1, pointer(syn), _FMT_I64, decoder.s, pointer(err), _FMT_F64, decoder.n,
                          pointer(conv), C_NULL, decoder.max_iters > 0 ? pointer(ratio) : C_NULL, C_NULL))
    end
    if decoder.variant === :sumproduct
        decoder.scratch.log_probabs .= log.(1 ./ ratio)      # posterior ratio R = P(1)/P(0); same expression, Julia's own log
    else
        decoder.scratch.log_probabs .= ratio                 # the LLR variants return log(P(0)/P(1)) directly
    end
    return decoder.scratch.err, conv[1]
end

"""
    batchdecode!(decoder::BeliefPropagationDecoder, syndromes, errors, success) -> (errors, success)

All columns in ONE library call (sharded over `devices`); `errors` and `success` are written in place.
"""
function batchdecode!(decoder::BeliefPropagationDecoder, syndromes::AbstractMatrix, errors::AbstractMatrix, success::AbstractVector{Bool})
    @assert size(syndromes, 2) == size(errors, 2)
    @assert size(syndromes, 2) == length(success)
    B = size(syndromes, 2)
    syn = syndromes isa _B200_IN ? syndromes : Matrix{Int64}(syndromes)
    err = errors isa _B200_OUT ? errors : Matrix{Int64}(undef, decoder.n, B)
    conv = success isa Vector{Bool} ? success : Vector{Bool}(undef, B)
    GC.@preserve syn err conv begin
        _b200_check(ccall((:ldpcb200_decode_batch, LDPCB200_LIB), Cint,
                          (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Ptr{Int32}, Ptr{Float64}, Ptr{Int64}),
                          decoder.handle, B, _b200_ptr(syn), _b200_fmt(syn), decoder.s, _b200_ptr(err), _b200_fmt(err), decoder.n,
                          pointer(conv), C_NULL, C_NULL, C_NULL))
    end
    err === errors || (errors .= err)
    conv === success || (success .= conv)
    return errors, success
end

# ---- extras that have no counterpart in the original file (thin wrappers of include/ldpcb200.h) -------------------------
"Logical operators `L` (k x n, k <= 64) for failure counting in `sample_decode_score`; `nothing` removes them."
function set_logicals!(decoder::BeliefPropagationDecoder, L)
    if L === nothing
        _b200_check(ccall((:ldpcb200_set_logicals, LDPCB200_LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int32),
                          decoder.handle, 0, C_NULL, C_NULL, Int32(1)))
        return decoder
    end
    Ls = SparseArrays.SparseMatrixCSC{Bool,Int}(dropzeros(sparse(L)))
    colptr = Vector{Int64}(Ls.colptr)
    rowval = Vector{Int64}(Ls.rowval)
    GC.@preserve colptr rowval _b200_check(ccall((:ldpcb200_set_logicals, LDPCB200_LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Int32),
                                                 decoder.handle, size(Ls, 1), colptr, rowval, Int32(1)))
    return decoder
end

"""
    sample_decode_score(decoder, shots; first=0, seed=12345, per_channel=decoder.per, osd=false) -> NamedTuple

The loop of test/test_bp_decoder.jl:19-30 (draw errors, take syndromes, decode, compare) entirely on the GPUs.
"""
function sample_decode_score(decoder::BeliefPropagationDecoder, shots::Integer; first::Integer=0, seed::Integer=12345,
                             per_channel::Float64=decoder.per, osd::Bool=false)
    out = zeros(Int64, 8)
    GC.@preserve out _b200_check(ccall((:ldpcb200_sample_decode_score, LDPCB200_LIB), Cint,
                                       (Ptr{Cvoid}, Int64, Int64, UInt64, Float64, Int32, Ptr{Int64}),
                                       decoder.handle, shots, first, UInt64(seed), per_channel, Int32(osd), out))
    return (shots=out[1], converged=out[2], iterations=out[3], exact_matches=out[4], syndrome_satisfied=out[5],
            failures=out[6], residual_weight=out[7], osd_processed=out[8])
end
