"""Host-side mirror of GenericSchur.jl's interface for the Schur hot path, over libgschur_cuda (B200, sm_100a).

The reference's host language (Julia) is not available in this environment; its shim is julia/GenericSchurCUDA.jl
(see INTEGRATION.md).  This module is the same thin layer in Python, over the same C ABI, with the reference's
names, argument meaning and error behaviour:

  gschur(A; wantZ, scale, maxiter)      src/GenericSchur.jl:343      (copying)
  gschur_(A; ...)                       gschur!  src/GenericSchur.jl:350-372, 805-835   (A is overwritten by T)
  schur / schur_                        LinearAlgebra.schur!  src/pirates.jl:8-10
  eigvals / eigvals_                    LinearAlgebra.eigvals! src/pirates.jl:17-27  (wantZ=false + sorteig!)
  hessenberg / hessenberg_              LinearAlgebra.hessenberg! src/pirates.jl:232 -> _hessenberg! src/hessenberg.jl:3-17
  gschur_hess_(H, Z)                    gschur!(H::Hessenberg, Z) src/GenericSchur.jl:194-335, 513-699
  Schur(T, Z, values)                   LinearAlgebra.Schur{Ty,S,C}  (fields T, Z / vectors, values)
  UnconvergedException                  src/GenericSchur.jl:39-45

Array layout is Julia's: column-major.  A single matrix is an (n, n) Fortran-ordered ndarray; a batch is
(n, n, batch) Fortran-ordered (`Array{T,3}`).  Element kinds:
  float64 (n, n[, batch])                   Float64
  complex128 (n, n[, batch])                ComplexF64
  float64 (2, n, n[, batch]) via DD(...)    double-double, limbs (hi, lo)
  float64 (4, n, n[, batch]) via CDD(...)   Complex{double-double}, (re.hi, re.lo, im.hi, im.lo)
Everything executes on the GPU; there is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import C64, CDD, DD, F64

__all__ = [
    "gschur", "gschur_", "gschur_batched_", "schur", "schur_", "eigvals", "eigvals_", "hessenberg", "hessenberg_",
    "gschur_hess_", "gschur_device_", "Schur", "Hessenberg", "UnconvergedException", "DimensionMismatch",
    "ArgumentError", "DDArray", "CDDArray", "F64", "C64", "DD", "CDD", "device_count", "launch_count", "release_workspace", "geigvecs", "eigen", "eigen_", "balance", "balance_", "balancer_lmul_", "triangularize", "Balancer",
]


class UnconvergedException(Exception):
    """Mirror of GenericSchur.UnconvergedException (src/GenericSchur.jl:39-45)."""


class DimensionMismatch(ValueError):
    """Mirror of Julia's DimensionMismatch (checksquare, src/GenericSchur.jl:354,811; :523-525)."""


class ArgumentError(ValueError):
    """Mirror of Julia's ArgumentError (src/GenericSchur.jl:206-210)."""


class DDArray(np.ndarray):
    """float64 array whose leading axis of length 2 holds (hi, lo) limbs of double-double numbers."""


class CDDArray(np.ndarray):
    """float64 array whose leading axis of length 4 holds (re.hi, re.lo, im.hi, im.lo)."""


def _as_dd(x):
    return np.asfortranarray(x, dtype=np.float64).view(DDArray)


def _as_cdd(x):
    return np.asfortranarray(x, dtype=np.float64).view(CDDArray)


class Schur:
    """LinearAlgebra.Schur: A = Z * T * Z'.  `values` pairs with diag(T).  Batched results carry a trailing axis."""

    def __init__(self, T, Z, values, info=None, stats=None):
        self.T = T
        self.Z = Z
        self.values = values
        self.info = info
        self.stats = stats

    @property
    def vectors(self):
        return self.Z

    @property
    def Schur(self):
        return self.T

    def __iter__(self):   # destructuring: T, Z, values = schur(A)
        return iter((self.T, self.Z, self.values))


class Hessenberg:
    """Result of hessenberg!: `factors` holds H on/above the sub-diagonal and the reflector tails below
    (H.H.data === A in the reference, src/hessenberg.jl:16), `tau` the reflector scalars, `Q` the explicit
    unitary factor (_materializeQ, src/hessenberg.jl:150-166)."""

    def __init__(self, factors, tau, Q):
        self.factors = factors
        self.tau = tau
        self.Q = Q

    @property
    def H(self):
        n = self.factors.shape[-2 if self.factors.ndim % 2 == 0 or self.factors.ndim == 2 else -3]
        F = np.array(self.factors, copy=True, order="F")
        if F.ndim == 2:
            return np.triu(F, -1)
        raise NotImplementedError("H view is provided for single f64/c64 matrices; use factors otherwise")


def device_count():
    return _lib.lib().gschur_cuda_device_count()


def launch_count():
    return int(_lib.lib().gschur_cuda_launch_count())


def measure_fp64_peak():
    """(TFLOP/s, ms) of a DFMA-only micro-kernel on the current device: the FP64 roofline denominator."""
    t = ctypes.c_double(0.0)
    ms = ctypes.c_double(0.0)
    rc = _lib.lib().gschur_cuda_measure_fp64_peak(ctypes.byref(t), ctypes.byref(ms))
    if rc != 0:
        raise RuntimeError(f"gschur_cuda_measure_fp64_peak failed: {rc}")
    return t.value, ms.value


def measure_dmma_peak():
    """(TFLOP/s, ms) of a DMMA-only (mma.sync.m8n8k4.f64) micro-kernel: the FP64 tensor roofline denominator of the
    large-matrix path's GEMM updates."""
    t = ctypes.c_double(0.0)
    ms = ctypes.c_double(0.0)
    rc = _lib.lib().gschur_cuda_measure_dmma_peak(ctypes.byref(t), ctypes.byref(ms))
    if rc != 0:
        raise RuntimeError(f"gschur_cuda_measure_dmma_peak failed: {rc}")
    return t.value, ms.value


def measure_l2_bandwidth():
    """(GB/s read + written, ms) of an L2-resident read-modify-write stream on the current device: the ceiling of
    stage B's Z stream."""
    g = ctypes.c_double(0.0)
    ms = ctypes.c_double(0.0)
    rc = _lib.lib().gschur_cuda_measure_l2_bandwidth(ctypes.byref(g), ctypes.byref(ms))
    if rc != 0:
        raise RuntimeError(f"gschur_cuda_measure_l2_bandwidth failed: {rc}")
    return g.value, ms.value


def max_batched_n(kind):
    return _lib.lib().gschur_cuda_max_batched_n(kind)


# ----------------------------------------------------------------------------------------------------------
def _kind_and_shape(A):
    """(kind, lead, n, batch or None) from an array in the layouts above; raises like the reference would."""
    if isinstance(A, CDDArray):
        kind, lead = CDD, 1
    elif isinstance(A, DDArray):
        kind, lead = DD, 1
    elif A.dtype == np.complex128:
        kind, lead = C64, 0
    elif A.dtype == np.float64:
        kind, lead = F64, 0
    else:
        # LinearAlgebra.schur! only has methods for T<:STypes (AbstractFloat / Complex{<:AbstractFloat}),
        # src/GenericSchur.jl:24; test/errors.jl:16-28 expects a MethodError for Int / Rational.
        raise TypeError(f"MethodError: no method matching gschur!(::Array{{{A.dtype}}})")
    core = A.shape[lead:]
    if lead and A.shape[0] != (2 if kind == DD else 4):
        raise DimensionMismatch("leading limb axis has the wrong length")
    if len(core) == 2:
        m, n = core
        batch = None
    elif len(core) == 3:
        m, n, batch = core
    else:
        raise DimensionMismatch(f"expected a matrix or a batch of matrices, got shape {A.shape}")
    if m != n:
        raise DimensionMismatch(f"matrix is not square: dimensions are ({m}, {n})")
    return kind, lead, n, batch


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _check_rc(rc, maxiter, n):
    if rc == 0:
        return
    msg = _lib.last_error()
    if rc > 0:
        raise UnconvergedException(f"iteration limit {maxiter if maxiter and maxiter > 0 else 100 * n} reached")
    if rc == _lib.ERR_SUBDIAG:
        raise ArgumentError("algorithm assumes real subdiagonal")
    if rc == _lib.ERR_ARG and "DimensionMismatch" in msg:
        raise DimensionMismatch(msg)
    if rc == _lib.ERR_ARG:
        raise ArgumentError(msg)
    raise RuntimeError(f"libgschur_cuda error {rc}: {msg}")


def _eig_array(kind, n, batch):
    shape_b = () if batch is None else (batch,)
    if kind in (F64, C64):
        return np.zeros((n,) + shape_b, dtype=np.complex128, order="F")
    return np.zeros((4, n) + shape_b, dtype=np.float64, order="F").view(CDDArray)


def _gschur_large_(A, wantZ, scale, check):
    """One large Float64 matrix (n > 128): blocked Hessenberg + multi-bulge QR with GEMM updates (regime 2)."""
    n = A.shape[0]
    Z = np.zeros_like(A) if wantZ else None
    w = np.zeros(n, dtype=np.complex128)
    info = ctypes.c_int(0)
    st = (ctypes.c_longlong * 3)()
    rc = _lib.lib().gschur_cuda_large(n, _ptr(A), n, _ptr(Z), n, _ptr(w), int(bool(scale)), ctypes.byref(info), st, 0)
    if rc < 0:
        raise RuntimeError(f"libgschur_cuda error {rc}: {_lib.lib().gschur_cuda_large_last_error().decode()}")
    if rc > 0 and check:
        raise UnconvergedException(f"iteration limit reached (active block ends at row {info.value})")
    if not wantZ:
        Z = np.zeros((0, 0), dtype=A.dtype)
    return Schur(A, Z, w, info=int(info.value), stats=np.array(list(st), dtype=np.int64))


def gschur_(A, wantZ=True, scale=True, maxiter=None, devices=None, check=True, flags=0, Z=None):
    """gschur!(A; wantZ, scale, maxiter): A (Fortran-ordered, see module doc) is overwritten by T.

    Works on one matrix or on a batch (trailing axis).  Returns Schur(T=A, Z, values[, info, stats]).
    `devices`: list of CUDA device ordinals the batch is split over (contiguous slices, no collective).
    A single Float64 matrix larger than the batched kernels take (n > 128) goes to the large-matrix path.
    """
    kind, lead, n, batch = _kind_and_shape(A)
    if not (A.flags.f_contiguous and A.flags.writeable):
        raise ArgumentError("A must be a writeable Fortran-ordered (column-major) array")
    if kind == F64 and batch is None and n > max_batched_n(F64) and flags == 0 and Z is None:
        return _gschur_large_(A, wantZ, scale, check)
    nb = 1 if batch is None else batch
    if wantZ:
        if Z is None:
            Z = np.zeros_like(A)
        elif Z.shape != A.shape or Z.dtype != A.dtype or not Z.flags.f_contiguous:
            raise DimensionMismatch("second dimension of Z must match H")
    else:
        Z = None
    w = _eig_array(kind, n, batch)
    info = np.zeros(nb, dtype=np.int32)
    stats = np.zeros((_lib.STATS_PER_MATRIX, nb), dtype=np.uint32, order="F")
    devs = None
    ndev = 0
    if devices is not None:
        devs = (ctypes.c_int * len(devices))(*devices)
        ndev = len(devices)
    rc = _lib.lib().gschur_cuda_batched(
        kind, n, nb, _ptr(A), n, n * n, _ptr(Z), n, n * n, _ptr(w), int(bool(scale)),
        int(maxiter) if maxiter else 0, _ptr(info), _ptr(stats), devs, ndev, flags)
    if check:
        _check_rc(rc, maxiter, n)
    if wantZ is False:
        # the reference returns a 0x0 matrix when wantZ=false (src/GenericSchur.jl:334, 698)
        Z = np.zeros((0, 0), dtype=A.dtype)
    return Schur(A, Z, w, info=info if batch is not None else int(info[0]), stats=stats)


gschur_batched_ = gschur_


def _copy_f(A):
    cls = type(A) if isinstance(A, (DDArray, CDDArray)) else None
    B = np.array(A, order="F", copy=True)
    return B.view(cls) if cls is not None else B


def gschur(A, **kw):
    """gschur(A) = gschur!(Matrix(A))  (src/GenericSchur.jl:343)."""
    A = np.asarray(A) if not isinstance(A, np.ndarray) else A
    _kind_and_shape(A)
    return gschur_(_copy_f(A), **kw)


def schur_(A, **kw):
    """LinearAlgebra.schur!(A) for T<:STypes (src/pirates.jl:8-10)."""
    return gschur_(A, **kw)


def schur(A, **kw):
    return gschur(A, **kw)


def _sorteig(values, sortby):
    """stdlib sorteig!: default eigsortby = λ -> (real(λ), imag(λ))."""
    if sortby is None:
        return values
    if values.ndim == 1:
        return np.array(sorted(values, key=sortby))
    out = np.empty_like(values)
    for b in range(values.shape[1]):
        out[:, b] = sorted(values[:, b], key=sortby)
    return out


def eigsortby(lam):
    return (lam.real, lam.imag)


def eigvals_(A, sortby=eigsortby, **kw):
    """LinearAlgebra.eigvals!(A; sortby) (src/pirates.jl:17-27): gschur!(A; wantZ=false) then sorteig!."""
    S = gschur_(A, wantZ=False, **kw)
    v = S.values
    if isinstance(v, CDDArray):
        return v   # double-double eigenvalues are returned unsorted limbs; sort on (hi) upstream if needed
    return _sorteig(v, sortby)


def eigvals(A, sortby=eigsortby, **kw):
    A = np.asarray(A) if not isinstance(A, np.ndarray) else A
    _kind_and_shape(A)
    return eigvals_(_copy_f(A), sortby=sortby, **kw)


def hessenberg_(A, wantQ=True, devices=None):
    """hessenberg!(A) (src/pirates.jl:232 -> _hessenberg!, src/hessenberg.jl:3-17) and _materializeQ."""
    kind, lead, n, batch = _kind_and_shape(A)
    if not (A.flags.f_contiguous and A.flags.writeable):
        raise ArgumentError("A must be a writeable Fortran-ordered (column-major) array")
    nb = 1 if batch is None else batch
    tshape = A.shape[:lead] + (max(n - 1, 0),) + (() if batch is None else (batch,))
    if kind == F64 and batch is None and n > max_batched_n(F64):
        # one large Float64 matrix: the blocked WY reduction of regime 2 (csrc/large_gehrd.cuh)
        tau = np.zeros(n - 1, dtype=np.float64)
        Q = np.zeros_like(A) if wantQ else None
        rc = _lib.lib().gschur_cuda_hessenberg_large(n, _ptr(A), n, _ptr(tau), _ptr(Q), n, 0)
        if rc != 0:
            raise RuntimeError(f"libgschur_cuda error {rc}: {_lib.lib().gschur_cuda_large_last_error().decode()}")
        return Hessenberg(A, tau, Q)
    tau = np.zeros(tshape, dtype=A.dtype, order="F")
    if isinstance(A, (DDArray, CDDArray)):
        tau = tau.view(type(A))
    Q = np.zeros_like(A) if wantQ else None
    devs = None
    ndev = 0
    if devices is not None:
        devs = (ctypes.c_int * len(devices))(*devices)
        ndev = len(devices)
    rc = _lib.lib().gschur_cuda_hessenberg_batched(kind, n, nb, _ptr(A), n, n * n, _ptr(tau), _ptr(Q), n, n * n,
                                                  devs, ndev, 0)
    _check_rc(rc, None, n)
    return Hessenberg(A, tau, Q)


def hessenberg(A, **kw):
    A = np.asarray(A) if not isinstance(A, np.ndarray) else A
    _kind_and_shape(A)
    return hessenberg_(_copy_f(A), **kw)


def gschur_hess_(H, Z=None, maxiter=None, checksd=True):
    """gschur!(H::Hessenberg, Z): H upper Hessenberg (overwritten by T), Z updated in place if given
    (src/GenericSchur.jl:194-335 complex, 513-699 real).  Raises ArgumentError for a non-real sub-diagonal
    (checksd) and DimensionMismatch when Z does not match (test/errors.jl:1-10)."""
    kind, lead, n, batch = _kind_and_shape(H)
    if Z is not None:
        if Z.shape[lead + 1] != n or Z.shape != H.shape:
            raise DimensionMismatch("second dimension of Z must match H")
    flags = _lib.FLAG_HESS_INPUT | (_lib.FLAG_CHECK_SUBDIAG if checksd else 0)
    return gschur_(H, wantZ=Z is not None, scale=False, maxiter=maxiter, flags=flags, Z=Z)


def geigvecs(S, left=False, normalize=True):
    """geigvecs(S; left) (src/vectors.jl:12-20): eigenvectors from a Schur decomposition `S` — _geigvecs! /
    _gleigvecs! on (S.T, S.Z) followed by _enormalize!.  ComplexF64, one matrix or a batch (trailing axis)."""
    T, Z = S.T, S.Z
    kind, lead, n, batch = _kind_and_shape(T)
    if kind != C64:
        raise ArgumentError("eigenvectors from the Schur form are implemented for ComplexF64 only")
    nb = 1 if batch is None else batch
    haveZ = Z is not None and Z.size > 0
    V = np.zeros_like(T, order="F")
    rc = _lib.lib().gschur_cuda_eigvecs_batched(kind, n, nb, _ptr(T), n, n * n, _ptr(Z) if haveZ else None, n, n * n,
                                               _ptr(V), n, n * n, int(bool(left)), 0 if normalize else 0x10)
    if rc != 0:
        raise RuntimeError(f"libgschur_cuda error {rc}: {_lib.lib().gschur_cuda_eigvecs_last_error().decode()}")
    return V


def eigen(A, **kw):
    """eigen(A) without balancing (src/pirates.jl:63-90 with permute = scale = false): (values, vectors) of a ComplexF64
    matrix or batch: gschur! followed by the eigenvectors of the Schur form, both on the GPU."""
    S = gschur(A, **kw)
    return S.values, geigvecs(S)


class Balancer:
    """Mirror of GenericSchur.Balancer (src/balance.jl:3-10): ilo, ihi, prow, pcol, D, trivial (batched: trailing axis)."""

    def __init__(self, ilo, ihi, D, perm, trivial):
        self.ilo, self.ihi, self.D, self.perm, self.trivial = ilo, ihi, D, perm, trivial

    @property
    def prow(self):
        return self.perm[: self.ilo - 1] if np.ndim(self.ilo) == 0 else None

    @property
    def pcol(self):
        return self.perm[self.ihi:] if np.ndim(self.ihi) == 0 else None

    def _raw(self):
        ii = np.stack([np.atleast_1d(self.ilo), np.atleast_1d(self.ihi), np.atleast_1d(self.trivial).astype(np.int32)],
                      axis=1).astype(np.int32)
        return np.ascontiguousarray(ii), np.asfortranarray(self.D, dtype=np.float64), np.asfortranarray(self.perm, dtype=np.int32)


def balance_(A, scale=True, permute=True):
    """balance!(A; scale, permute) => (Abal, B::Balancer)   (src/balance.jl:33-199).  Float64 / ComplexF64, one matrix or
    a batch (trailing axis); A is overwritten."""
    kind, lead, n, batch = _kind_and_shape(A)
    if kind not in (F64, C64):
        raise ArgumentError("balance! is implemented for Float64 and ComplexF64")
    if not (A.flags.f_contiguous and A.flags.writeable):
        raise ArgumentError("A must be a writeable Fortran-ordered (column-major) array")
    nb = 1 if batch is None else batch
    D = np.zeros((n, nb), order="F")
    perm = np.zeros((n, nb), dtype=np.int32, order="F")
    ii = np.zeros((nb, 3), dtype=np.int32)
    info = np.zeros(nb, dtype=np.int32)
    rc = _lib.lib().gschur_cuda_balance_batched(kind, n, nb, _ptr(A), n, n * n, _ptr(D), _ptr(ii), _ptr(perm), _ptr(info),
                                               int(bool(scale)), int(bool(permute)), 0)
    if rc != 0:
        raise RuntimeError(f"libgschur_cuda error {rc}: {_lib.lib().gschur_cuda_balance_last_error().decode()}")
    if (info != 0).any():
        raise RuntimeError("NaN encountered while balancing")
    if batch is None:
        return A, Balancer(int(ii[0, 0]), int(ii[0, 1]), D[:, 0], perm[:, 0], bool(ii[0, 2]))
    return A, Balancer(ii[:, 0].copy(), ii[:, 1].copy(), D, perm, ii[:, 2].astype(bool))


def balance(A, **kw):
    return balance_(_copy_f(np.asarray(A)), **kw)


def balancer_lmul_(B, V, inverse=False):
    """lmul!(B::Balancer, V) (right eigenvectors) / ldiv!(B, V) (inverse=True, left eigenvectors), src/balance.jl:203-260"""
    kind, lead, n, batch = _kind_and_shape(V)
    nb = 1 if batch is None else batch
    ii, D, perm = B._raw()
    rc = _lib.lib().gschur_cuda_balance_apply_batched(kind, n, nb, _ptr(V), n, n * n, _ptr(D), _ptr(ii), _ptr(perm),
                                                     int(bool(inverse)), 0)
    if rc != 0:
        raise RuntimeError(f"libgschur_cuda error {rc}: {_lib.lib().gschur_cuda_balance_last_error().decode()}")
    return V


def triangularize(S):
    """triangularize(S::Schur{<:Real}) => Schur{Complex} (src/triang.jl:9-43), one matrix or a batch."""
    T, Z = S.T, S.Z
    kind, lead, n, batch = _kind_and_shape(T)
    if kind != F64:
        raise ArgumentError("triangularize takes a real (Float64) Schur decomposition")
    nb = 1 if batch is None else batch
    haveZ = Z is not None and Z.size > 0
    shape = T.shape
    Tc = np.zeros(shape, dtype=np.complex128, order="F")
    Zc = np.zeros(shape, dtype=np.complex128, order="F") if haveZ else None
    w = np.zeros((n,) if batch is None else (n, batch), dtype=np.complex128, order="F")
    rc = _lib.lib().gschur_cuda_triangularize_batched(n, nb, _ptr(T), n, n * n, _ptr(Z) if haveZ else None, n, n * n,
                                                     _ptr(Tc), _ptr(Zc), _ptr(w), 0)
    if rc != 0:
        raise RuntimeError(f"libgschur_cuda error {rc}: {_lib.lib().gschur_cuda_balance_last_error().decode()}")
    return Schur(Tc, Zc if haveZ else np.zeros((0, 0), dtype=np.complex128), w)


def eigen_(A, permute=True, scale=True):
    """eigen!(A; permute, scale) (src/pirates.jl:63-90, the path without condition numbers): balance!, gschur!
    (+ triangularize for Float64), eigenvectors of the Schur form, lmul!(B, v), _enormalize! — every step on the GPU.
    Returns (values, vectors), unsorted (sortby = nothing)."""
    kind, lead, n, batch = _kind_and_shape(A)
    if kind not in (F64, C64):
        raise ArgumentError("eigen! is implemented for Float64 and ComplexF64")
    B = None
    if permute or scale:
        A, B = balance_(A, scale=scale, permute=permute)
    S = gschur_(A)
    if kind == F64:
        S = triangularize(S)
    v = geigvecs(S, normalize=False)
    if B is not None:
        balancer_lmul_(B, v)
    v = _enormalize_(v)
    return S.values, v


def _enormalize_(v):
    """_enormalize! (src/util.jl:572-592) through the eigenvector kernel's normalisation pass (identity T, Z = v)"""
    kind, lead, n, batch = _kind_and_shape(v)
    nb = 1 if batch is None else batch
    eye = np.zeros_like(v, order="F")
    idx = np.arange(n)
    eye[idx, idx, ...] = 1.0
    out = np.zeros_like(v, order="F")
    rc = _lib.lib().gschur_cuda_eigvecs_batched(C64, n, nb, _ptr(eye), n, n * n, _ptr(v), n, n * n, _ptr(out), n, n * n, 0, 0)
    if rc != 0:
        raise RuntimeError(f"libgschur_cuda error {rc}: {_lib.lib().gschur_cuda_eigvecs_last_error().decode()}")
    return out


def release_workspace():
    """Free the device / pinned buffers the library caches between calls (gschur_cuda_release_workspace)."""
    return _lib.lib().gschur_cuda_release_workspace()


def gschur_device_(kind, n, batch, A_ptr, Z_ptr, w_ptr, info_ptr=None, stats_ptr=None, scale=True, maxiter=0,
                   stream=None, lda=None, strideA=None, ldz=None, strideZ=None):
    """Device-resident, asynchronous gschur! over raw device pointers (e.g. torch tensors' data_ptr()) on the
    current CUDA device; enqueued on `stream` (a cudaStream_t as int).  Used by bench.py for the HBM-resident
    timing; convergence is reported through the info array only."""
    lda = n if lda is None else lda
    ldz = n if ldz is None else ldz
    strideA = n * n if strideA is None else strideA
    strideZ = n * n if strideZ is None else strideZ
    rc = _lib.lib().gschur_cuda_batched_async(
        kind, n, batch, ctypes.c_void_p(A_ptr), lda, strideA, ctypes.c_void_p(Z_ptr) if Z_ptr else None, ldz,
        strideZ, ctypes.c_void_p(w_ptr), int(bool(scale)), int(maxiter),
        ctypes.c_void_p(info_ptr) if info_ptr else None, ctypes.c_void_p(stats_ptr) if stats_ptr else None,
        ctypes.c_void_p(stream) if stream else None, 0)
    if rc != 0:
        _check_rc(rc, maxiter, n)
    return rc
