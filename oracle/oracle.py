"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this
module; the product package never does.  The oracle is a C++ restatement of the reference's algorithm
(oracle/gschur_oracle.hpp cites the reference file:line of every routine).

Array conventions (all column-major / Fortran order):
  kind 0  f64          (n, n) float64
  kind 1  c64          (n, n) complex128
  kind 2  dd           (2, n, n) float64   [hi, lo]
  kind 3  complex dd   (4, n, n) float64   [re.hi, re.lo, im.hi, im.lo]
Eigenvalues: complex128 (n,) for kinds 0/1, (4, n) float64 for kinds 2/3.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgschur_oracle.so")
_lib = None

KIND_F64, KIND_C64, KIND_DD, KIND_CDD = 0, 1, 2, 3


def build(force=False):
    """Compile oracle/libgschur_oracle.so with the Makefile beside this file."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        L = _lib
        vp, ci, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_long
        L.gso_gschur.argtypes = [ci, ci, vp, cl, vp, cl, vp, ci, ci, vp]
        L.gso_gschur_mp.argtypes = [ci, ci, vp, cl, vp, vp, vp, ci]
        L.gso_residuals.argtypes = [ci, ci, vp, cl, vp, cl, vp, cl, vp]
        L.gso_hessenberg.argtypes = [ci, ci, vp, cl, vp, vp, cl]
        L.gso_gschur_hess.argtypes = [ci, ci, vp, cl, vp, cl, vp, ci, ci]
        L.gso_balance.argtypes = [ci, ci, vp, cl, ci, ci, vp, vp, vp]
        L.gso_balance_apply.argtypes = [ci, ci, vp, cl, vp, vp, vp, ci]
        L.gso_eigvalscond.argtypes = [ci, ci, vp, cl, vp]
        L.gso_gs2x2.argtypes = [vp, vp, vp]
        L.gso_reflector.argtypes = [ci, ci, vp, vp]
        L.gso_gschur_batched.argtypes = [ci, ci, cl, vp, vp, vp, ci, ci, vp]
    return _lib


def kind_of(A):
    """Infer the element kind from an array in the conventions above (last two dims are the matrix)."""
    if np.iscomplexobj(A):
        return KIND_C64
    if A.ndim >= 3 and A.shape[0] == 2 and A.dtype == np.float64 and getattr(A, "_gs_kind", None) != 0:
        return KIND_DD
    return KIND_F64


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _prep(A, kind):
    lead = {0: (), 1: (), 2: (2,), 3: (4,)}[kind]
    dt = np.complex128 if kind == 1 else np.float64
    A = np.array(A, dtype=dt, order="F", copy=True)
    assert A.shape[: len(lead)] == lead and A.shape[-1] == A.shape[-2], (A.shape, kind)
    return A


def _empty_like_w(kind, n):
    return np.zeros(n, dtype=np.complex128) if kind < 2 else np.zeros((4, n), dtype=np.float64, order="F")


def gschur(A, kind, wantZ=True, scale=True, maxiter=0):
    """gschur!(copy(A); wantZ, scale) -> (T, Z or None, w, rc, stats).  rc: 0 ok, 1 unconverged."""
    T = _prep(A, kind)
    n = T.shape[-1]
    Z = np.zeros_like(T) if wantZ else None
    w = _empty_like_w(kind, n)
    stats = np.zeros(4, dtype=np.int64)
    rc = lib().gso_gschur(kind, n, _ptr(T), n, _ptr(Z), n, _ptr(w), int(scale), int(maxiter), _ptr(stats))
    return T, Z, w, rc, stats


def gschur_mp(A, kind, scale=True):
    """Decomposition in MPFR-256 (BigFloat(256) stand-in); outputs rounded to double-double arrays."""
    A = _prep(A, kind)
    n = A.shape[-1]
    lead = 4 if (kind & 1) else 2
    T = np.zeros((lead, n, n), order="F")
    Z = np.zeros((lead, n, n), order="F")
    w = np.zeros((4, n), order="F")
    rc = lib().gso_gschur_mp(kind, n, _ptr(A), n, _ptr(T), _ptr(Z), _ptr(w), int(scale))
    return T, Z, w, rc


def residuals(A, T, Z, kind):
    """(backward error ratio, orthogonality ratio, ||A||_F) in MPFR-256 — the reference's test statistics."""
    A, T, Z = _prep(A, kind), _prep(T, kind), _prep(Z, kind)
    n = A.shape[-1]
    out = np.zeros(3)
    lib().gso_residuals(kind, n, _ptr(A), n, _ptr(T), n, _ptr(Z), n, _ptr(out))
    return float(out[0]), float(out[1]), float(out[2])


def hessenberg(A, kind, wantQ=True):
    """_hessenberg!(copy(A)) (+ _materializeQ) -> (factors, tau, Q)."""
    F = _prep(A, kind)
    n = F.shape[-1]
    lead = F.shape[:-2]
    dt = F.dtype
    tau = np.zeros(lead + (max(n - 1, 1),), dtype=dt, order="F")
    Q = np.zeros_like(F) if wantQ else None
    rc = lib().gso_hessenberg(kind, n, _ptr(F), n, _ptr(tau), _ptr(Q), n)
    assert rc == 0
    return F, tau[..., : max(n - 1, 0)], Q


def gschur_hess(H, kind, Z=None, maxiter=0, checksd=True):
    """gschur!(Hessenberg(H), Z) -> (T, Z, w, rc); rc -2 = ArgumentError (complex sub-diagonal)."""
    T = _prep(H, kind)
    n = T.shape[-1]
    Zc = _prep(Z, kind) if Z is not None else None
    w = _empty_like_w(kind, n)
    rc = lib().gso_gschur_hess(kind, n, _ptr(T), n, _ptr(Zc), n, _ptr(w), int(maxiter), int(checksd))
    return T, Zc, w, rc


def eigvalscond(T, kind):
    T = _prep(T, kind)
    n = T.shape[-1]
    s = np.zeros(n)
    rc = lib().gso_eigvalscond(kind, n, _ptr(T), n, _ptr(s))
    assert rc == 0
    return s


def balance(A, scale=True, permute=True):
    """balance!(A) -> (Abal, D, sp, (ilo, ihi, trivial), rc)   (src/balance.jl:33-199); Float64 / ComplexF64"""
    kind = 1 if np.iscomplexobj(A) else 0
    Ab = _prep(A, kind)
    n = Ab.shape[-1]
    D = np.zeros(n)
    sp = np.zeros(n, dtype=np.int32)
    ii = np.zeros(3, dtype=np.int32)
    rc = lib().gso_balance(kind, n, _ptr(Ab), n, int(scale), int(permute), _ptr(D), _ptr(sp), _ptr(ii))
    return Ab, D, sp, (int(ii[0]), int(ii[1]), bool(ii[2])), rc


def balance_apply(V, D, sp, ii, inverse=False):
    """lmul!(B, V) / ldiv!(B, V)   (src/balance.jl:203-260)"""
    kind = 1 if np.iscomplexobj(V) else 0
    Vb = _prep(V, kind)
    n = Vb.shape[-1]
    iia = np.array([ii[0], ii[1], int(ii[2])], dtype=np.int32)
    rc = lib().gso_balance_apply(kind, n, _ptr(Vb), n, _ptr(np.ascontiguousarray(D, dtype=np.float64)),
                                 _ptr(np.ascontiguousarray(sp, dtype=np.int32)), _ptr(iia), int(inverse))
    assert rc == 0
    return Vb


def gs2x2(a, b, c, d):
    abcd = np.array([a, b, c, d], dtype=np.float64)
    csn = np.zeros(2)
    w = np.zeros(4)
    lib().gso_gs2x2(_ptr(abcd), _ptr(csn), _ptr(w))
    return abcd, csn, complex(w[0], w[1]), complex(w[2], w[3])


def reflector(x):
    x = np.array(x, copy=True)
    kind = 1 if np.iscomplexobj(x) else 0
    x = x.astype(np.complex128 if kind else np.float64)
    tau = np.zeros(1, dtype=x.dtype)
    lib().gso_reflector(kind, x.shape[0], _ptr(x), _ptr(tau))
    return x, tau[0]


def gschur_batched(A, kind, wantZ=True, scale=True, nthreads=1):
    """Batch driver used as the CPU baseline: A is (..., n, n, batch) Fortran-ordered; in place.
    Returns (T, Z, w, info)."""
    A = np.asfortranarray(A)
    n = A.shape[-2]
    batch = A.shape[-1]
    Z = np.zeros_like(A) if wantZ else None
    w = np.zeros((n, batch), dtype=np.complex128, order="F") if kind < 2 else np.zeros((4, n, batch), order="F")
    info = np.zeros(batch, dtype=np.int32)
    lib().gso_gschur_batched(kind, n, batch, _ptr(A), _ptr(Z), _ptr(w), int(scale), int(nthreads), _ptr(info))
    return A, Z, w, info


# ---- double-double helpers for the tests -----------------------------------------------------------------

def dd_from_float(x):
    x = np.asarray(x, dtype=np.float64)
    return np.asfortranarray(np.stack([x, np.zeros_like(x)], axis=0))


def dd_to_float(x):
    return x[0] + x[1]
