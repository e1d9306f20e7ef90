# LDPCDecodersB200.jl -- Julia-side binding of libldpcb200.so (include/ldpcb200.h).
#
# UNEXECUTED IN THE BUILD ENVIRONMENT (no julia binary in the image); the Python mirror
# ldpcdecoders.jl_b200/decoder.py drives exactly the same C entry points and is what the tests
# exercise.  Drop this file next to LDPCDecoders.jl and `include` it after the package: it adds a
# GPU decoder type with the reference's constructor / decode! / batchdecode! / reset! methods
# (src/decoders/belief_propagation.jl:61-67, 83-91, 121-188, 220-231), and the BP+OSD decoder with osd_order = 0
# (src/decoders/belief_propagation_osd.jl:17-125).
module LDPCDecodersB200

using SparseArrays
import LDPCDecoders: AbstractDecoder, decode!, batchdecode!, reset!

const LIB = get(ENV, "LDPCB200_LIB", "libldpcb200.so")

const FMT_U8, FMT_I64, FMT_BITS, FMT_PACKED32, FMT_F64 = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)

struct B200Error <: Exception
    code::Int32
    msg::String
end

function check(rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:ldpcb200_last_error, LIB), Cstring, ()))
    throw(B200Error(Int32(rc), msg))
end

"Scratch the reference exposes and BP+OSD reads (belief_propagation_osd.jl:51-52)."
struct B200ScratchSpace
    log_probabs::Vector{Float64}
    err::Vector{Float64}
end

"""
    B200BeliefPropagationDecoder(H, per::Float64, max_iters::Int; devices=Int32[0])

Same fields as `BeliefPropagationDecoder` (belief_propagation.jl:38-59); the Tanner graph and all
messages live on the GPU(s) behind `handle`.
"""
mutable struct B200BeliefPropagationDecoder <: AbstractDecoder
    per::Float64
    max_iters::Int
    s::Int
    n::Int
    sparse_H::SparseMatrixCSC{Bool,Int}
    sparse_HT::SparseMatrixCSC{Bool,Int}
    scratch::B200ScratchSpace
    handle::Ptr{Cvoid}
end

function B200BeliefPropagationDecoder(H, per::Float64, max_iters::Int; devices::Vector{Int32}=Int32[0])
    s, n = size(H)
    sparse_H = SparseMatrixCSC{Bool,Int}(dropzeros(sparse(H)))
    sparse_HT = SparseMatrixCSC{Bool,Int}(sparse(sparse_H'))
    colptr = Vector{Int64}(sparse_H.colptr)     # 1-based, n+1
    rowval = Vector{Int64}(sparse_H.rowval)     # 1-based, ascending inside a column
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve colptr rowval devices begin
        check(ccall((:ldpcb200_create, LIB), Cint,
                    (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Float64, Int32, Int32, Ptr{Int32}, Int32, Ptr{Ptr{Cvoid}}),
                    s, n, colptr, rowval, Int32(1), per, Int32(max_iters), Int32(0), devices, Int32(length(devices)), out))
    end
    dec = B200BeliefPropagationDecoder(per, max_iters, s, n, sparse_H, sparse_HT,
                                       B200ScratchSpace(zeros(n), zeros(n)), out[])
    finalizer(d -> (d.handle != C_NULL && ccall((:ldpcb200_destroy, LIB), Cint, (Ptr{Cvoid},), d.handle); d.handle = C_NULL), dec)
    return dec
end

"No host scratch to clear: messages are (re)initialised inside the kernels."
function reset!(dec::B200BeliefPropagationDecoder)
    dec.scratch.log_probabs .= 0.0
    dec.scratch.err .= 0.0
    dec
end

# element format of a dense Julia matrix / BitMatrix at the C boundary
fmt_of(::BitMatrix) = FMT_BITS
fmt_of(::Matrix{Bool}) = FMT_U8
fmt_of(::Matrix{UInt8}) = FMT_U8
fmt_of(::Matrix{Int64}) = FMT_I64
fmt_of(::Matrix{Float64}) = FMT_F64
hostptr(A::BitMatrix) = pointer(A.chunks)
hostptr(A::Matrix) = pointer(A)

function batchdecode!(dec::B200BeliefPropagationDecoder, syndromes::AbstractMatrix, errors::AbstractMatrix,
                      success::AbstractVector{Bool})
    @assert size(syndromes, 2) == size(errors, 2)
    @assert size(syndromes, 2) == length(success)
    B = size(syndromes, 2)
    syn = syndromes isa Union{BitMatrix,Matrix{Bool},Matrix{UInt8},Matrix{Int64}} ? syndromes : Matrix{Int64}(syndromes)
    err = errors isa Union{BitMatrix,Matrix{Bool},Matrix{UInt8},Matrix{Int64},Matrix{Float64}} ? errors : Matrix{Int64}(undef, dec.n, B)
    conv = success isa Vector{Bool} ? success : Vector{Bool}(undef, B)
    GC.@preserve syn err conv begin
        check(ccall((:ldpcb200_decode_batch, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Ptr{Int32}, Ptr{Float64}, Ptr{Int64}),
                    dec.handle, B, hostptr(syn), fmt_of(syn), dec.s, hostptr(err), fmt_of(err), dec.n,
                    pointer(conv), C_NULL, C_NULL, C_NULL))
    end
    err === errors || (errors .= err)
    conv === success || (success .= conv)
    return errors, success
end

function decode!(dec::B200BeliefPropagationDecoder, syndrome::AbstractVector)
    syn = reshape(Vector{Int64}(syndrome), dec.s, 1)
    err = reshape(dec.scratch.err, dec.n, 1)            # aliased Float64 0.0/1.0, as :187 returns
    conv = Vector{Bool}(undef, 1)
    ratio = ones(Float64, dec.n)
    GC.@preserve syn err conv ratio begin
        check(ccall((:ldpcb200_decode_batch, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Ptr{Int32}, Ptr{Float64}, Ptr{Int64}),
                    dec.handle, 1, pointer(syn), FMT_I64, dec.s, pointer(err), FMT_F64, dec.n, pointer(conv), C_NULL,
                    dec.max_iters > 0 ? pointer(ratio) : C_NULL, C_NULL))
    end
    dec.scratch.log_probabs .= log.(1 ./ ratio)        # belief_propagation.jl:163, Julia's own log
    return dec.scratch.err, conv[1]
end

# ---- BP + OSD-0  (src/decoders/belief_propagation_osd.jl:17-61; osd(..., Val(0)) :63-125) -------------
"""
    B200BeliefPropagationOSDDecoder(H, per::Float64, max_iters::Int; osd_order::Int=0)

Same fields as `BeliefPropagationOSDDecoder` (belief_propagation_osd.jl:17-24).  Only `osd_order = 0` is
implemented on the GPU; BP and the OSD-0 elimination of the unconverged syndromes run in one library call.
"""
struct B200BeliefPropagationOSDDecoder <: AbstractDecoder
    bp_decoder::B200BeliefPropagationDecoder
    H::BitMatrix
    osd_order::Int
end

function B200BeliefPropagationOSDDecoder(H, per::Float64, max_iters::Int; osd_order::Int=0, kw...)
    osd_order == 0 || error("B200BeliefPropagationOSDDecoder: only osd_order = 0 runs on the GPU")
    return B200BeliefPropagationOSDDecoder(B200BeliefPropagationDecoder(H, per, max_iters; kw...), BitMatrix(H), 0)
end

reset!(dec::B200BeliefPropagationOSDDecoder) = (reset!(dec.bp_decoder); dec)

# generic batchdecode! (abstract_decoder.jl:31-42) over decode!(::BeliefPropagationOSDDecoder), done in one call
function batchdecode!(dec::B200BeliefPropagationOSDDecoder, syndromes::AbstractMatrix, errors::AbstractMatrix,
                      converged::AbstractVector{Bool})
    @assert size(syndromes, 2) == size(errors, 2)
    @assert size(syndromes, 2) == length(converged)
    bp = dec.bp_decoder
    B = size(syndromes, 2)
    syn = syndromes isa Union{BitMatrix,Matrix{Bool},Matrix{UInt8},Matrix{Int64}} ? syndromes : Matrix{Int64}(syndromes)
    err = errors isa Union{BitMatrix,Matrix{Bool},Matrix{UInt8},Matrix{Int64},Matrix{Float64}} ? errors : Matrix{Int64}(undef, bp.n, B)
    conv = converged isa Vector{Bool} ? converged : Vector{Bool}(undef, B)
    GC.@preserve syn err conv begin
        check(ccall((:ldpcb200_bposd_decode_batch, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32, Int64, Ptr{UInt8}, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
                    bp.handle, B, hostptr(syn), fmt_of(syn), bp.s, hostptr(err), fmt_of(err), bp.n,
                    pointer(conv), C_NULL, C_NULL, C_NULL))
    end
    err === errors || (errors .= err)
    conv === converged || (converged .= conv)
    return errors, converged
end

# decode! returns a fresh Bool vector and BP's converged flag (belief_propagation_osd.jl:60)
function decode!(dec::B200BeliefPropagationOSDDecoder, syndrome::AbstractVector)
    errors = Matrix{Bool}(undef, dec.bp_decoder.n, 1)
    conv = Vector{Bool}(undef, 1)
    batchdecode!(dec, reshape(Vector{Int64}(syndrome), dec.bp_decoder.s, 1), errors, conv)
    return vec(errors), conv[1]
end

end # module
