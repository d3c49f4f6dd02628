# GenericSchurCUDA.jl — Julia host shim over libgschur_cuda (include/gschur_cuda.h).
#
# Adds CUDA-backed methods behind GenericSchur.jl's own entry points so that `gschur!`, `gschur`,
# `schur!` and `eigvals!` return the same `LinearAlgebra.Schur{T, Z, values}` objects as the pure-Julia
# code they replace:
#   gschur!(A::StridedMatrix{Complex{T}}; wantZ, scale)   src/GenericSchur.jl:350-372
#   gschur!(A::StridedMatrix{T<:AbstractFloat}; wantZ, scale)   src/GenericSchur.jl:805-835
#   LinearAlgebra.schur!    src/pirates.jl:8-10        LinearAlgebra.eigvals!  src/pirates.jl:17-27
#   LinearAlgebra.hessenberg! -> _hessenberg!   src/pirates.jl:232, src/hessenberg.jl:3-17
#
# NOTE: no Julia toolchain exists in the build environment, so this file has never been executed; it is
# deliberately thin and mirrors genericschur.jl_b200/__init__.py (which is exercised by the test-suite over the
# same C ABI) line for line.  `Float64x2` below stands for any isbits double-double type laid out as two
# consecutive Float64 (hi, lo): MultiFloats.Float64x2 or DoubleFloats.Double64.
module GenericSchurCUDA

using LinearAlgebra
import GenericSchur
import GenericSchur: gschur!, UnconvergedException

const libgschur = get(ENV, "GSCHUR_CUDA_LIB", "libgschur_cuda")

const GSCHUR_F64, GSCHUR_C64, GSCHUR_DD, GSCHUR_CDD = Cint(0), Cint(1), Cint(2), Cint(3)
const GSCHUR_ERR_ARG, GSCHUR_ERR_CUDA, GSCHUR_ERR_SIZE, GSCHUR_ERR_SUBDIAG = -1, -2, -3, -4

_kind(::Type{Float64}) = GSCHUR_F64
_kind(::Type{ComplexF64}) = GSCHUR_C64
# double-double element types: any isbits struct of two Float64 limbs, high limb first
_kind(::Type{T}) where {T <: AbstractFloat} =
    (isbitstype(T) && sizeof(T) == 16) ? GSCHUR_DD : throw(MethodError(gschur!, (Matrix{T},)))
_kind(::Type{Complex{T}}) where {T <: AbstractFloat} =
    (isbitstype(T) && sizeof(T) == 16) ? GSCHUR_CDD : throw(MethodError(gschur!, (Matrix{Complex{T}},)))

_lasterr() = unsafe_string(ccall((:gschur_cuda_last_error, libgschur), Cstring, ()))

function _check(rc::Integer, maxiter::Integer)
    rc == 0 && return
    rc > 0 && throw(UnconvergedException("iteration limit $maxiter reached"))
    rc == GSCHUR_ERR_SUBDIAG && throw(ArgumentError("algorithm assumes real subdiagonal"))
    msg = _lasterr()
    rc == GSCHUR_ERR_ARG && occursin("DimensionMismatch", msg) && throw(DimensionMismatch(msg))
    rc == GSCHUR_ERR_ARG && throw(ArgumentError(msg))
    error("libgschur_cuda error $rc: $msg")
end

"""
    gschur_batched!(A::Array{T,3}; wantZ=true, scale=true, maxiter=100n, devices=Cint[0]) -> (T, Z, values, info)

Schur decomposition of every `A[:, :, b]` on the GPU(s).  `A` is overwritten by the Schur forms.
"""
function gschur_batched!(A::Array{T, 3}; wantZ::Bool = true, scale::Bool = true,
        maxiter::Integer = 100 * size(A, 1), devices::Vector{Cint} = Cint[0]) where {T}
    n = LinearAlgebra.checksquare(view(A, :, :, 1))
    batch = size(A, 3)
    CT = T <: Complex ? T : Complex{T}
    Z = wantZ ? similar(A) : Array{T, 3}(undef, 0, 0, 0)
    w = Array{CT, 2}(undef, n, batch)
    info = zeros(Int32, batch)
    rc = ccall((:gschur_cuda_batched, libgschur), Cint,
        (Cint, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Cint,
         Ptr{Int32}, Ptr{UInt32}, Ptr{Cint}, Cint, UInt32),
        _kind(T), n, batch, A, n, n * n, wantZ ? pointer(Z) : C_NULL, n, n * n, w, scale, maxiter,
        info, C_NULL, devices, length(devices), 0)
    _check(rc, maxiter)
    return A, Z, w, info
end

# --- the drop-in methods: same signatures and return types as the reference ---------------------------------
for ET in (Float64, ComplexF64)
    @eval function gschur!(A::Matrix{$ET}; wantZ::Bool = true, scale::Bool = true,
            maxiter::Integer = 100 * size(A, 1), kwargs...)
        n = LinearAlgebra.checksquare(A)           # DimensionMismatch for non-square input, as the reference
        A3 = reshape(A, n, n, 1)
        T3, Z3, w, _ = gschur_batched!(A3; wantZ = wantZ, scale = scale, maxiter = maxiter)
        Z = wantZ ? reshape(Z3, n, n) : Matrix{$ET}(undef, 0, 0)     # src/GenericSchur.jl:334, 698
        return LinearAlgebra.Schur(reshape(T3, n, n), Z, vec(w))
    end
end

# double-double element types go through the same entry (kind 2 / 3); declared generically so that any
# two-limb type dispatches here while BigFloat, Float16 ... keep using the pure-Julia methods.
function gschur_dd!(A::Matrix{T}; wantZ::Bool = true, scale::Bool = true,
        maxiter::Integer = 100 * size(A, 1)) where {T}
    n = LinearAlgebra.checksquare(A)
    T3, Z3, w, _ = gschur_batched!(reshape(A, n, n, 1); wantZ = wantZ, scale = scale, maxiter = maxiter)
    Z = wantZ ? reshape(Z3, n, n) : Matrix{T}(undef, 0, 0)
    return LinearAlgebra.Schur(reshape(T3, n, n), Z, vec(w))
end

"""
    hessenberg_cuda!(A) -> (factors, τ, Q)     (LinearAlgebra.hessenberg! for T<:STypes, src/pirates.jl:232)
"""
function hessenberg_cuda!(A::Matrix{T}) where {T}
    n = LinearAlgebra.checksquare(A)
    τ = Vector{T}(undef, max(n - 1, 0))
    Q = similar(A)
    rc = ccall((:gschur_cuda_hessenberg_batched, libgschur), Cint,
        (Cint, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cint}, Cint, UInt32),
        _kind(T), n, 1, A, n, n * n, τ, Q, n, n * n, C_NULL, 0, 0)
    _check(rc, 0)
    return A, τ, Q
end

end # module
