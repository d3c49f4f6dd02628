"""Shared test helpers: the reference's acceptance checks (test/complex.jl:1-34, test/real.jl:3-74,
test/testfuncs.jl:57-86) restated over numpy, and the reference's test-matrix classes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from __graft_entry__ import load_oracle, load_package  # noqa: E402

ULP = np.finfo(np.float64).eps
UNFL = np.finfo(np.float64).tiny
OVFL = 1.0 / UNFL
ULPINV = 1.0 / ULP
RTULPI = 1.0 / np.sqrt(ULP)
SIMRCONDS = [1.0, RTULPI, 0.0]
MAGNS = [1.0, OVFL * ULP, UNFL * ULPINV]
SAFEMIN = UNFL


def fnorm(A):
    """Frobenius norm without over/underflow (inputs span unfl/ulp .. ovfl*ulp)."""
    m = float(np.max(np.abs(A))) if A.size else 0.0
    if m == 0 or not np.isfinite(m):
        return m
    return m * float(np.linalg.norm(A / m))


def csort(v):
    """csort of test/testfuncs.jl:1-2: sort by (real, imag)."""
    v = np.asarray(v)
    return v[np.lexsort((v.imag, v.real))]


def checkblocks(T):
    """test/real.jl:3-22: every 2x2 diagonal block is in standard form."""
    n = T.shape[0]
    ok = True
    for j in range(n - 1):
        if T[j + 1, j] != 0:
            ok &= T[j, j] == T[j + 1, j + 1]
            ok &= T[j, j + 1] != 0
            ok &= np.sign(T[j, j + 1]) * np.sign(T[j + 1, j]) < 0
    return bool(ok)


def checkeigvals(T, w, tol):
    """test/testfuncs.jl:57-86."""
    n = T.shape[0]
    ok = True
    for j in range(n):
        ok &= T[j, j] == w[j].real
    if n > 1:
        if T[1, 0] == 0:
            ok &= w[0].imag == 0
        if T[n - 1, n - 2] == 0:
            ok &= w[n - 1].imag == 0
    for j in range(n - 1):
        if T[j + 1, j] != 0:
            t = np.sqrt(abs(T[j + 1, j])) * np.sqrt(abs(T[j, j + 1]))
            cmp = max(ULP * t, SAFEMIN)
            ok &= abs(w[j].imag - t) / cmp < tol
            ok &= abs(w[j + 1].imag + t) / cmp < tol
        elif j > 0 and T[j + 1, j] == 0 and T[j, j - 1] == 0:
            ok &= w[j].imag == 0
    return bool(ok)


def structure_ok(T, w, kind, tol=20):
    """The exact (`==`) post-conditions: test/complex.jl:8,20 and test/real.jl:33,35,45."""
    if kind == 1:
        if not np.all(np.tril(T, -1) == 0):
            return False, "T not upper triangular"
        if not np.array_equal(csort(w), csort(np.diag(T))):
            return False, "values != diag(T)"
        return True, ""
    if not np.all(np.tril(T, -2) == 0):
        return False, "T not quasi-triangular"
    if not checkblocks(T):
        return False, "2x2 block not in standard form"
    if not checkeigvals(T, w, tol):
        return False, "values inconsistent with T"
    return True, ""


def match_eigs(w, wref, tol=None):
    """Match two spectra (the reference's tests sort both; sorting is fragile for close conjugate pairs and for
    clusters of ill-conditioned eigenvalues, so an optimal assignment on the tolerance-scaled distance is used).
    Returns the per-eigenvalue distances aligned with wref."""
    from scipy.optimize import linear_sum_assignment
    w = np.asarray(w)
    wref = np.asarray(wref)
    D = np.abs(w[None, :] - wref[:, None])           # D[j, i] = |w_i - wref_j|
    if tol is None:
        cost = D
    else:
        cost = np.minimum(D / np.asarray(tol)[:, None], 1e6) ** 2
    r, c = linear_sum_assignment(cost)
    out = np.zeros(len(wref))
    out[r] = D[r, c]
    return out


def randu(rng, n):
    """random unitary, test/testfuncs.jl:94-105"""
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Q, R = np.linalg.qr(A)
    d = np.diag(R)
    return np.asfortranarray(Q * (d / np.abs(d)))


def rando(rng, n):
    """random orthogonal, test/testfuncs.jl:108-119"""
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return np.asfortranarray(Q)


def godunov():
    """test/testfuncs.jl:127-142"""
    A = np.array([
        [289, 2064, 336, 128, 80, 32, 16],
        [1152, 30, 1312, 512, 288, 128, 32],
        [-29, -2000, 756, 384, 1008, 224, 48],
        [512, 128, 640, 0, 640, 512, 128],
        [1053, 2256, -504, -384, -756, 800, 208],
        [-287, -16, 1712, -128, 1968, -30, 2032],
        [-2176, -287, -1565, -512, -541, -1152, -289]], dtype=np.float64)
    return np.asfortranarray(A), np.array([-4, -2, -1, 0, 1, 2, 4], dtype=np.float64), 7.0e16


def reference_classes(complex_, sizes=(4, 32), seed=1234):
    """Yield (name, A, tol) over the matrix classes of test/complex.jl:88-135,221-389 and
    test/real.jl:104-138,180-309 (Float64 / ComplexF64 instances).  LAPACK-generated classes are produced
    with the same iseed through tests/tmg.py."""
    import tmg
    rng = np.random.default_rng(seed + (1 if complex_ else 0))

    def rnd(n):
        A = rng.random((n, n))
        if complex_:
            A = A + 1j * rng.random((n, n))
        return np.asfortranarray(A)

    for n in sizes:
        yield f"rand_n{n}", rnd(n), 10
        yield f"normal_n{n}", (randu(rng, n) if complex_ else rando(rng, n)), 20
        dt = np.complex128 if complex_ else np.float64
        yield f"jordan_n{n}", np.asfortranarray(np.diag(np.ones(n - 1, dtype=dt), 1) + np.eye(n, dtype=dt)), 20
        kmagn = [1, 1, 1, 1, 1, 1, 1, 1, 2, 3]
        kmode = [4, 3, 1, 5, 4, 3, 1, 5, 5, 5]
        kconds = [1, 1, 1, 1, 2, 2, 2, 2, 2, 2]
        for j in range(10):
            A = tmg.latme(n, MAGNS[kmagn[j] - 1], kmode[j], ULPINV, SIMRCONDS[kconds[j] - 1], complex_=complex_)
            yield f"latme{j}_n{n}", A, 20
        yield f"diag_n{n}", tmg.latmr(n, 1.0, 6, 1.0, complex_=complex_, kl=0, ku=0), 20
        yield f"sym_n{n}", tmg.latmr(n, 1.0, 6, 1.0, complex_=complex_, sym="H" if complex_ else "S"), 20
        for j, m in enumerate(MAGNS):
            yield f"latmr_general{j}_n{n}", tmg.latmr(n, m, 6, 1.0, complex_=complex_), 20
        yield f"triangular_n{n}", tmg.latmr(n, 1.0, 6, 1.0, complex_=complex_, kl=0), 20
    # Hessenberg-test scalings (test/complex.jl:88-96, test/real.jl:104-116)
    A = rnd(32)
    yield "rand_tiny_n32", np.asfortranarray(A * (100 * UNFL)), 10
    if not complex_:
        yield "rand_huge_n32", np.asfortranarray(A * (np.finfo(np.float64).max / 100)), 10
    if complex_:
        yield "short_sweep_n5", tmg.latme(5, 1.0, 4, ULPINV, 1.0, complex_=True, iseed=(4066, 2905, 502, 2389)), 20
