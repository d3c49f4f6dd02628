"""LAPACK test-matrix generators (xLATME / xLATMR) and LAPACK originals (xLAHQR, xGEHD2) reached through the
OpenBLAS that scipy bundles — the same routines the reference's tests call through test/TMGlib.jl:58-356,
with the same default iseed = [2518, 3899, 995, 397] (test/TMGlib.jl:66,134,282).

Used to regenerate the reference's matrix classes (test/complex.jl:221-389, test/real.jl:180-309) and to write the
committed fixtures under tests/golden/ (tests/golden/make_golden.py); nothing here runs on the GPU box's hot path.
"""
import ctypes
import glob
import os

import numpy as np

_lib = None
DEFAULT_ISEED = (2518, 3899, 995, 397)


def lib():
    global _lib
    if _lib is None:
        import scipy
        pat = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so")
        hits = glob.glob(pat)
        if not hits:
            raise OSError("scipy's bundled OpenBLAS not found")
        _lib = ctypes.CDLL(hits[0])
    return _lib


def _i(x):
    return ctypes.byref(ctypes.c_int(x))


def _d(x):
    return ctypes.byref(ctypes.c_double(x))


def _c(ch):
    return ctypes.c_char_p(ch.encode())


def _z(x):
    return (ctypes.c_double * 2)(x.real, x.imag)


def latme(n, anorm, imode, rcond, simrcond=1.0, complex_=False, iseed=DEFAULT_ISEED, dist=None, upper=True,
          simtrans=True, simmode=4):
    """latme!(A, anorm, imode, rcond, simrcond) as wrapped at test/TMGlib.jl:58-123 (complex) / :126-239 (real)."""
    L = lib()
    seed = (ctypes.c_int * 4)(*iseed)
    info = ctypes.c_int(0)
    one = ctypes.c_long(1)
    if complex_:
        A = np.zeros((n, n), dtype=np.complex128, order="F")
        d = np.zeros(n, dtype=np.complex128)
        ds = np.zeros(n)
        work = np.zeros(3 * n, dtype=np.complex128)
        L.scipy_zlatme_(_i(n), _c(dist or "D"), seed, d.ctypes, _i(imode), _d(rcond), _z(1.0 + 0j), _c("T"),
                        _c("T" if upper else "F"), _c("T" if simtrans else "F"), ds.ctypes, _i(simmode), _d(simrcond),
                        _i(n), _i(n), _d(anorm), A.ctypes, _i(n), work.ctypes, ctypes.byref(info), one, one, one, one)
    else:
        A = np.zeros((n, n), dtype=np.float64, order="F")
        d = np.zeros(n)
        ds = np.zeros(n)
        work = np.zeros(3 * n)
        ei = ctypes.c_char_p(b" ")
        L.scipy_dlatme_(_i(n), _c(dist or "S"), seed, d.ctypes, _i(imode), _d(rcond), _d(1.0), ei, _c("T"),
                        _c("T" if upper else "F"), _c("T" if simtrans else "F"), ds.ctypes, _i(simmode), _d(simrcond),
                        _i(n), _i(n), _d(anorm), A.ctypes, _i(n), work.ctypes, ctypes.byref(info), one, one, one, one,
                        one)
    if info.value != 0:
        raise RuntimeError(f"latme info = {info.value}")
    return A


def latmr(n, anorm, imode, rcond, complex_=False, sym="N", kl=None, ku=None, iseed=DEFAULT_ISEED, dist=None):
    """latmr!(A, anorm, imode, rcond; sym, kl, ku) as wrapped at test/TMGlib.jl:271-356 (defaults as there:
    grade 'N', pivoting 'N', sparse 0, pack 'N', rsign true)."""
    L = lib()
    kl = n if kl is None else kl
    ku = n if ku is None else ku
    seed = (ctypes.c_int * 4)(*iseed)
    info = ctypes.c_int(0)
    one = ctypes.c_long(1)
    iwork = np.zeros(n, dtype=np.int32)
    ipivot = np.zeros(n, dtype=np.int32)
    if complex_:
        A = np.zeros((n, n), dtype=np.complex128, order="F")
        d = np.zeros(n, dtype=np.complex128)
        dl = np.zeros(n, dtype=np.complex128)
        dr = np.zeros(n, dtype=np.complex128)
        L.scipy_zlatmr_(_i(n), _i(n), _c(dist or "D"), seed, _c(sym), d.ctypes, _i(imode), _d(rcond), _z(1.0 + 0j),
                        _c("T"), _c("N"), dl.ctypes, _i(1), _d(1.0), dr.ctypes, _i(1), _d(1.0), _c("N"), ipivot.ctypes,
                        _i(kl), _i(ku), _d(0.0), _d(anorm), _c("N"), A.ctypes, _i(n), iwork.ctypes,
                        ctypes.byref(info), one, one, one, one, one, one)
    else:
        A = np.zeros((n, n), dtype=np.float64, order="F")
        d = np.zeros(n)
        dl = np.zeros(n)
        dr = np.zeros(n)
        L.scipy_dlatmr_(_i(n), _i(n), _c(dist or "S"), seed, _c(sym), d.ctypes, _i(imode), _d(rcond), _d(1.0),
                        _c("T"), _c("N"), dl.ctypes, _i(1), _d(1.0), dr.ctypes, _i(1), _d(1.0), _c("N"), ipivot.ctypes,
                        _i(kl), _i(ku), _d(0.0), _d(anorm), _c("N"), A.ctypes, _i(n), iwork.ctypes,
                        ctypes.byref(info), one, one, one, one, one, one)
    if info.value != 0:
        raise RuntimeError(f"latmr info = {info.value}")
    return A


def lapack_gehd2(A):
    """LAPACK xGEHD2 (the Fortran original of src/hessenberg.jl:3-17): returns (factors, tau)."""
    L = lib()
    A = np.array(A, order="F", copy=True)
    n = A.shape[0]
    tau = np.zeros(max(n - 1, 1), dtype=A.dtype)
    work = np.zeros(n, dtype=A.dtype)
    info = ctypes.c_int(0)
    f = L.scipy_zgehd2_ if np.iscomplexobj(A) else L.scipy_dgehd2_
    f(_i(n), _i(1), _i(n), A.ctypes, _i(n), tau.ctypes, work.ctypes, ctypes.byref(info))
    assert info.value == 0
    return A, tau[: n - 1]


def lapack_lahqr(H, wantz=True):
    """LAPACK xLAHQR on an upper Hessenberg matrix (the Fortran original of src/GenericSchur.jl:194-335,
    513-699): returns (T, Z, w)."""
    L = lib()
    H = np.array(H, order="F", copy=True)
    n = H.shape[0]
    info = ctypes.c_int(0)
    if np.iscomplexobj(H):
        Z = np.eye(n, dtype=np.complex128, order="F")
        w = np.zeros(n, dtype=np.complex128)
        L.scipy_zlahqr_(_i(1), _i(1 if wantz else 0), _i(n), _i(1), _i(n), H.ctypes, _i(n), w.ctypes, _i(1), _i(n),
                        Z.ctypes, _i(n), ctypes.byref(info))
    else:
        Z = np.eye(n, order="F")
        wr = np.zeros(n)
        wi = np.zeros(n)
        L.scipy_dlahqr_(_i(1), _i(1 if wantz else 0), _i(n), _i(1), _i(n), H.ctypes, _i(n), wr.ctypes, wi.ctypes,
                        _i(1), _i(n), Z.ctypes, _i(n), ctypes.byref(info))
        w = wr + 1j * wi
    assert info.value == 0, info.value
    return H, Z, w


def lapack_lanv2(a, b, c, d):
    """LAPACK DLANV2 (the Fortran original of _gs2x2!, src/GenericSchur.jl:716-803)."""
    L = lib()
    v = [ctypes.c_double(x) for x in (a, b, c, d)]
    out = [ctypes.c_double(0.0) for _ in range(6)]   # rt1r rt1i rt2r rt2i cs sn
    L.scipy_dlanv2_(*[ctypes.byref(x) for x in v], *[ctypes.byref(x) for x in out])
    return [x.value for x in v], [x.value for x in out]
