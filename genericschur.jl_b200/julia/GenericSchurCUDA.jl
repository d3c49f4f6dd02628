# GenericSchurCUDA.jl — Julia host shim over libgschur_cuda (include/gschur_cuda.h).
#
# Adds CUDA-backed methods at GenericSchur.jl's own dispatch points so that `gschur!`, `gschur`, `schur!`, `eigvals!`
# and `hessenberg!` keep their signatures and return the same `LinearAlgebra.Schur{T, Z, values}` /
# `LinearAlgebra.Hessenberg` objects as the pure-Julia code they replace:
#
#   reference method (file:line)                                        shim method below              C symbol
#   gschur!(A::StridedMatrix{Complex{T}}; wantZ, scale)   GenericSchur.jl:350-372   gschur!(::Matrix{ComplexF64/Complex{DD}})   gschur_cuda_batched
#   gschur!(A::StridedMatrix{T}; wantZ, scale, Zarg, Zwrk)         :805-835         gschur!(::Matrix{Float64/DD})               gschur_cuda_batched, gschur_cuda_large
#   gschur!(H::Hessenberg{Complex{RT}}, Z; maxiter, checksd)       :194-335         gschur!(::Hessenberg{ComplexF64,...}, Z)    gschur_cuda_batched + GSCHUR_FLAG_HESS_INPUT
#   gschur!(H::Hessenberg{T}, Z; maxiter)                          :513-699         gschur!(::Hessenberg{Float64,...}, Z)       gschur_cuda_batched + GSCHUR_FLAG_HESS_INPUT
#   LinearAlgebra.schur!(A)  pirates.jl:8-10,  eigvals!(A)  :17-27                  reach the methods above through gschur!     —
#   LinearAlgebra.hessenberg!(A) pirates.jl:232 -> _hessenberg! hessenberg.jl:3-17  _hessenberg!(::Matrix{...})                 gschur_cuda_hessenberg_batched / _large
#
# Everything the library does not cover is forwarded to the package's own method with `invoke`, never refused:
# matrices above the batched kernels' size limit (except Float64, which has the large-matrix path), strided views,
# and calls that pass keywords the C ABI has no parameter for (`maxinner`, `tol`, `Zwrk`, `standardize`).
#
# STATUS: EXPERIMENTAL — no Julia toolchain exists in the build environment or on the GPU boxes, so this file has not
# been executed.  The same C ABI is exercised by the Python mirror genericschur.jl_b200/__init__.py (ctypes), function for
# function (gschur_ <-> gschur!, gschur_hess_ <-> gschur!(H, Z), hessenberg_ <-> hessenberg!, _gschur_large_).
# `DD` below is any isbits two-limb double-double type laid out as (hi, lo) Float64: MultiFloats.Float64x2,
# DoubleFloats.Double64.
module GenericSchurCUDA

using LinearAlgebra
import GenericSchur
import GenericSchur: gschur!, _hessenberg!, UnconvergedException

const libgschur = get(ENV, "GSCHUR_CUDA_LIB", "libgschur_cuda")

const GSCHUR_F64, GSCHUR_C64, GSCHUR_DD, GSCHUR_CDD = Cint(0), Cint(1), Cint(2), Cint(3)
const GSCHUR_ERR_ARG, GSCHUR_ERR_CUDA, GSCHUR_ERR_SIZE, GSCHUR_ERR_SUBDIAG = -1, -2, -3, -4
const GSCHUR_FLAG_HESS_INPUT, GSCHUR_FLAG_CHECK_SUBDIAG = UInt32(0x2), UInt32(0x4)

# ---- element kinds ------------------------------------------------------------------------------------------
_twolimb(::Type{T}) where {T} = isbitstype(T) && sizeof(T) == 16 && T <: AbstractFloat
_kind(::Type{Float64}) = GSCHUR_F64
_kind(::Type{ComplexF64}) = GSCHUR_C64
_kind(::Type{T}) where {T <: AbstractFloat} = _twolimb(T) ? GSCHUR_DD : nothing
_kind(::Type{Complex{T}}) where {T <: AbstractFloat} = _twolimb(T) ? GSCHUR_CDD : nothing
_kind(::Type) = nothing

_maxn(kind) = Int(ccall((:gschur_cuda_max_batched_n, libgschur), Cint, (Cint,), kind))
_lasterr() = unsafe_string(ccall((:gschur_cuda_last_error, libgschur), Cstring, ()))
_largeerr() = unsafe_string(ccall((:gschur_cuda_large_last_error, libgschur), Cstring, ()))

function _check(rc::Integer, maxiter::Integer)
    rc == 0 && return
    rc > 0 && throw(UnconvergedException("iteration limit $maxiter reached"))
    rc == GSCHUR_ERR_SUBDIAG && throw(ArgumentError("algorithm assumes real subdiagonal"))
    msg = _lasterr()
    rc == GSCHUR_ERR_ARG && occursin("DimensionMismatch", msg) && throw(DimensionMismatch(msg))
    rc == GSCHUR_ERR_ARG && throw(ArgumentError(msg))
    error("libgschur_cuda error $rc: $msg")
end

# keywords the C ABI carries; anything else sends the call to the pure-Julia method
const _SUPPORTED = (:wantZ, :scale, :maxiter, :Zarg, :checksd)
_supported(kwargs) = all(k -> k in _SUPPORTED, keys(kwargs))

# ---- the batched entry (new spelling; the reference's equivalent is Threads.@threads over gschur!) ----------------
"""
    gschur_batched!(A::Array{T,3}; wantZ=true, scale=true, maxiter=100n, devices=Cint[0], Z=nothing, flags=0)
        -> (T, Z, values, info)

Schur decomposition of every `A[:, :, b]` on the GPU(s).  `A` is overwritten by the Schur forms.
"""
function gschur_batched!(A::Array{T, 3}; wantZ::Bool = true, scale::Bool = true,
        maxiter::Integer = 100 * size(A, 1), devices::Vector{Cint} = Cint[0],
        Z::Union{Nothing, Array{T, 3}} = nothing, flags::UInt32 = UInt32(0)) where {T}
    size(A, 1) == size(A, 2) || throw(DimensionMismatch("matrix is not square: dimensions are $(size(A)[1:2])"))
    kind = _kind(T)
    kind === nothing && throw(ArgumentError("element type $T has no CUDA kernel"))
    n, batch = size(A, 1), size(A, 3)
    CT = T <: Complex ? T : Complex{T}
    Zb = wantZ ? (Z === nothing ? similar(A) : Z) : Array{T, 3}(undef, 0, 0, 0)
    wantZ && size(Zb) != size(A) && throw(DimensionMismatch("second dimension of Z must match H"))
    w = Array{CT, 2}(undef, n, batch)
    info = zeros(Int32, batch)
    (n == 0 || batch == 0) && return A, Zb, w, info            # nothing to do (the library accepts it too)
    rc = ccall((:gschur_cuda_batched, libgschur), Cint,
        (Cint, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Cint,
         Ptr{Int32}, Ptr{UInt32}, Ptr{Cint}, Cint, UInt32),
        kind, n, batch, A, n, n * n, wantZ ? pointer(Zb) : C_NULL, n, n * n, w, scale, maxiter,
        info, C_NULL, devices, length(devices), flags)
    _check(rc, maxiter)
    return A, Zb, w, info
end

# one large Float64 matrix: blocked WY Hessenberg + multi-bulge QR with DMMA GEMM updates (regime 2)
function _gschur_large!(A::Matrix{Float64}; wantZ::Bool, scale::Bool, maxiter::Integer)
    n = LinearAlgebra.checksquare(A)
    Z = wantZ ? similar(A) : Matrix{Float64}(undef, 0, 0)
    w = Vector{ComplexF64}(undef, n)
    info = Ref{Cint}(0)
    rc = ccall((:gschur_cuda_large, libgschur), Cint,
        (Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{ComplexF64}, Cint, Ptr{Cint}, Ptr{Int64}, UInt32),
        n, A, n, wantZ ? pointer(Z) : C_NULL, n, w, scale, info, C_NULL, 0)
    rc < 0 && error("libgschur_cuda error $rc: $(_largeerr())")
    rc > 0 && throw(UnconvergedException("iteration limit $maxiter reached"))
    return LinearAlgebra.Schur(A, Z, w)
end

# ---- gschur!(A): same signatures and return types as the reference ---------------------------------------------
function _gschur_cuda!(A::Matrix{ET}; wantZ::Bool = true, scale::Bool = true,
        maxiter::Integer = 100 * size(A, 1), Zarg::Union{Nothing, Matrix{ET}} = nothing) where {ET}
    n = LinearAlgebra.checksquare(A)           # DimensionMismatch for non-square input, as the reference
    Z3 = (wantZ && Zarg !== nothing) ? reshape(Zarg, n, n, 1) : nothing        # the caller's Z buffer (src/GenericSchur.jl:807, 821)
    T3, Zo, w, _ = gschur_batched!(reshape(A, n, n, 1); wantZ = wantZ, scale = scale, maxiter = maxiter, Z = Z3)
    Z = wantZ ? reshape(Zo, n, n) : Matrix{ET}(undef, 0, 0)                    # src/GenericSchur.jl:334, 698
    return LinearAlgebra.Schur(reshape(T3, n, n), Z, vec(w))
end

for ET in (Float64, ComplexF64)
    @eval function gschur!(A::Matrix{$ET}; kwargs...)
        n = LinearAlgebra.checksquare(A)
        if !_supported(kwargs)
            return invoke(gschur!, Tuple{StridedMatrix{$ET}}, A; kwargs...)
        end
        if n <= _maxn(_kind($ET))
            return _gschur_cuda!(A; kwargs...)
        elseif $ET === Float64 && !haskey(kwargs, :Zarg)
            kw = Dict{Symbol, Any}(kwargs)
            return _gschur_large!(A; wantZ = get(kw, :wantZ, true), scale = get(kw, :scale, true),
                maxiter = get(kw, :maxiter, 100 * n))
        else
            return invoke(gschur!, Tuple{StridedMatrix{$ET}}, A; kwargs...)   # ComplexF64 above the batched limit
        end
    end
end

# Two-limb (double-double) element types.  These methods are what `schur!` / `eigvals!` (src/pirates.jl:8-27) reach for
# MultiFloats.Float64x2 / DoubleFloats.Double64 matrices — the types for which the piracy is actually exercised.
# Declared for every AbstractFloat so that no type import is needed; other element types (BigFloat, Float16, Float32)
# fall through to the pure-Julia method at once.
function gschur!(A::Matrix{T}; kwargs...) where {T <: AbstractFloat}
    if _kind(T) === GSCHUR_DD && _supported(kwargs) && LinearAlgebra.checksquare(A) <= _maxn(GSCHUR_DD)
        return _gschur_cuda!(A; kwargs...)
    end
    return invoke(gschur!, Tuple{StridedMatrix{T}}, A; kwargs...)
end
function gschur!(A::Matrix{Complex{T}}; kwargs...) where {T <: AbstractFloat}
    if _kind(Complex{T}) === GSCHUR_CDD && _supported(kwargs) && LinearAlgebra.checksquare(A) <= _maxn(GSCHUR_CDD)
        return _gschur_cuda!(A; kwargs...)
    end
    return invoke(gschur!, Tuple{StridedMatrix{Complex{T}}}, A; kwargs...)
end

# ---- gschur!(H::Hessenberg, Z) (src/GenericSchur.jl:194-210, 513-525) ---------------------------------------------
# H carries the factored form; like the reference (`_getdata(H)`, then zero below the sub-diagonal) the upper Hessenberg
# part of its storage is the input and is overwritten by T.  Z, when given, is updated in place.
function _gschur_hess_cuda!(HH::Matrix{ET}, Z::Union{Nothing, Matrix{ET}}, maxiter::Integer, checksd::Bool) where {ET}
    n = LinearAlgebra.checksquare(HH)
    if Z !== nothing
        size(Z, 2) == n || throw(DimensionMismatch("second dimension of Z must match H"))
        size(Z, 1) == n || throw(DimensionMismatch("Z must be n x n for the CUDA path"))
    end
    flags = GSCHUR_FLAG_HESS_INPUT | (checksd ? GSCHUR_FLAG_CHECK_SUBDIAG : UInt32(0))
    Z3 = Z === nothing ? nothing : reshape(Z, n, n, 1)
    T3, Zo, w, _ = gschur_batched!(reshape(HH, n, n, 1); wantZ = Z !== nothing, scale = false, maxiter = maxiter,
        Z = Z3, flags = flags)
    Zr = Z === nothing ? Matrix{ET}(undef, 0, 0) : Z
    return LinearAlgebra.Schur(reshape(T3, n, n), Zr, vec(w))
end

# argument types of the reference methods these shadow (src/GenericSchur.jl:194-198, 513-518)
_href(::Type{T}) where {T <: Complex} = Tuple{Hessenberg{T}, Any}
_href(::Type{T}) where {T <: AbstractFloat} = Tuple{Hessenberg{T}, Union{Nothing, AbstractMatrix}}

function gschur!(H::Hessenberg{T, <:UpperHessenberg{T, Matrix{T}}}, Z::Union{Nothing, Matrix{T}} = nothing;
        kwargs...) where {T <: Union{AbstractFloat, Complex{<:AbstractFloat}}}
    kind = _kind(T)
    n = size(H, 1)
    ok = kind !== nothing && _supported(kwargs) && n <= _maxn(kind) && (Z === nothing || size(Z) == (n, n))
    if !ok
        return invoke(gschur!, _href(T), H, Z; kwargs...)
    end
    kw = Dict{Symbol, Any}(kwargs)
    # the complex method checks for a real sub-diagonal unless told otherwise (src/GenericSchur.jl:196, 206-210);
    # the real method has nothing to check
    checksd = T <: Complex ? get(kw, :checksd, true) : false
    return _gschur_hess_cuda!(GenericSchur._getdata(H), Z, get(kw, :maxiter, 100 * n), checksd)
end

# ---- hessenberg! (src/pirates.jl:232 -> _hessenberg!, src/hessenberg.jl:3-17): returns LinearAlgebra.Hessenberg(A, τ) ----
function _hessenberg!(A::Matrix{T}) where {T <: Union{AbstractFloat, Complex{<:AbstractFloat}}}
    n = LinearAlgebra.checksquare(A)
    kind = _kind(T)
    τ = Vector{T}(undef, max(n - 1, 0))
    if kind !== nothing && n <= _maxn(kind)
        n == 0 && return Hessenberg(A, τ)
        rc = ccall((:gschur_cuda_hessenberg_batched, libgschur), Cint,
            (Cint, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cint}, Cint, UInt32),
            kind, n, 1, A, n, n * n, τ, C_NULL, n, n * n, C_NULL, 0, 0)
        _check(rc, 0)
        return Hessenberg(A, τ)                                    # src/hessenberg.jl:16
    elseif T === Float64
        rc = ccall((:gschur_cuda_hessenberg_large, libgschur), Cint,
            (Cint, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Cint, UInt32), n, A, n, τ, C_NULL, n, 0)
        rc == 0 || error("libgschur_cuda error $rc: $(_largeerr())")
        return Hessenberg(A, τ)
    end
    return invoke(_hessenberg!, Tuple{StridedMatrix{T}}, A)
end

"""
    hessenberg_cuda!(A) -> (F::Hessenberg, Q::Matrix)

`hessenberg!(A)` plus the explicit unitary factor (`_materializeQ`, src/hessenberg.jl:150-166) from the same kernel.
"""
function hessenberg_cuda!(A::Matrix{T}) where {T}
    n = LinearAlgebra.checksquare(A)
    kind = _kind(T)
    (kind === nothing || n > _maxn(kind)) && throw(ArgumentError("no batched CUDA kernel for $T at n = $n"))
    τ = Vector{T}(undef, max(n - 1, 0))
    Q = similar(A)
    rc = ccall((:gschur_cuda_hessenberg_batched, libgschur), Cint,
        (Cint, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Ptr{Cint}, Cint, UInt32),
        kind, n, 1, A, n, n * n, τ, Q, n, n * n, C_NULL, 0, 0)
    _check(rc, 0)
    return Hessenberg(A, τ), Q
end

# ---- SURVEY.md section 8(f) ranks 2 and 3: eigenvectors from the Schur form, balancing, triangularize ------------------
# geigvecs(S; left) (src/vectors.jl:12-20) for ComplexF64 Schur decompositions; other element types keep the package's method
function GenericSchur.geigvecs(S::Schur{ComplexF64, Matrix{ComplexF64}}; left::Bool = false)
    n = size(S.T, 1)
    n <= 128 || return invoke(GenericSchur.geigvecs, Tuple{Schur{T}} where {T}, S; left = left)
    V = Matrix{ComplexF64}(undef, n, n)
    haveZ = size(S.Z, 1) > 0
    rc = ccall((:gschur_cuda_eigvecs_batched, libgschur), Cint,
        (Cint, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Cint, UInt32),
        GSCHUR_C64, n, 1, S.T, n, n * n, haveZ ? pointer(S.Z) : C_NULL, n, n * n, V, n, n * n, left, 0)
    rc == 0 || error("libgschur_cuda error $rc: " * unsafe_string(ccall((:gschur_cuda_eigvecs_last_error, libgschur), Cstring, ())))
    return V
end

# balance!(A; scale, permute) => (Abal, B::Balancer) (src/balance.jl:33-199); the p / algo keywords go to the package's method
function GenericSchur.balance!(A::Matrix{T}; scale = true, permute = true, kwargs...) where {T <: Union{Float64, ComplexF64}}
    isempty(kwargs) || return invoke(GenericSchur.balance!, Tuple{AbstractMatrix{T}}, A; scale = scale, permute = permute, kwargs...)
    n = LinearAlgebra.checksquare(A)
    D = Vector{Float64}(undef, n)
    ii = zeros(Int32, 3)
    sp = zeros(Int32, n)
    info = zeros(Int32, 1)
    rc = ccall((:gschur_cuda_balance_batched, libgschur), Cint,
        (Cint, Cint, Int64, Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Cint, Cint, UInt32),
        _kind(T), n, 1, A, n, n * n, D, ii, sp, info, scale, permute, 0)
    rc == 0 || error("libgschur_cuda error $rc")
    info[1] == 0 || error("NaN encountered while balancing")
    ilo, ihi = Int(ii[1]), Int(ii[2])
    B = GenericSchur.Balancer{T}(ilo, ihi, Int.(sp[1:(ilo - 1)]), Int.(sp[(ihi + 1):n]), T.(D), ii[3] != 0)
    return A, B
end

# triangularize(S::Schur{Float64}) (src/triang.jl:9-43)
function GenericSchur.triangularize(S::Schur{Float64, Matrix{Float64}})
    n = size(S.T, 1)
    Tc = Matrix{ComplexF64}(undef, n, n)
    Zc = Matrix{ComplexF64}(undef, n, n)
    w = Vector{ComplexF64}(undef, n)
    rc = ccall((:gschur_cuda_triangularize_batched, libgschur), Cint,
        (Cint, Int64, Ptr{Float64}, Cint, Int64, Ptr{Float64}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt32),
        n, 1, S.T, n, n * n, S.Z, n, n * n, Tc, Zc, w, 0)
    rc == 0 || error("libgschur_cuda error $rc")
    return Schur(Tc, Zc, w)
end

release_workspace() = ccall((:gschur_cuda_release_workspace, libgschur), Cint, ())

end # module
