"""GPU tests of the eigenvectors-from-the-Schur-form path (run with -m gpu).

  * vectest of the reference (test/testfuncs.jl:5-34): || A VR - VR diag(w) || / (n ||A|| ulp) < tol and the same for
    the left vectors with A', on the matrix classes of test/complex.jl;
  * the LAPACK original the reference states it is based on (ZTREVC, through scipy's OpenBLAS) after the reference's
    own normalisation (_enormalize!, src/util.jl:572-592, restated in numpy below);
  * the normalisation itself; the overflow-guarded solve on a badly scaled triangular matrix.
"""
import numpy as np
import pytest

from common import ULP, fnorm, reference_classes

pytestmark = pytest.mark.gpu


def enormalize(v):
    """_enormalize! (src/util.jl:572-592): unit 2-norm, the component of largest modulus real."""
    v = v.copy()
    for j in range(v.shape[1]):
        s = 1.0 / np.linalg.norm(v[:, j])
        i0 = int(np.argmax(np.abs(v[:, j]) ** 2))
        t = s * np.conj(v[i0, j]) / np.sqrt(np.abs(v[i0, j]) ** 2)
        v[:, j] *= t
        v[i0, j] = v[i0, j].real
    return v


def _vectest(A, S, VR, VL, tol, name):
    n = A.shape[0]
    w = S.values
    nA = fnorm(A)
    r = fnorm(A @ VR - VR * w[None, :]) / (n * nA * ULP)
    l = fnorm(A.conj().T @ VL - VL * np.conj(w)[None, :]) / (n * nA * ULP)
    assert r < tol and l < tol, (name, r, l)


def test_vectest_random_batches(gs):
    rng = np.random.default_rng(2024)
    for n, batch in ((5, 64), (32, 32), (64, 16)):
        A = np.asfortranarray(rng.random((n, n, batch)) + 1j * rng.random((n, n, batch)))
        S = gs.gschur(A)
        VR = gs.geigvecs(S)
        VL = gs.geigvecs(S, left=True)
        for b in range(batch):
            Sb = gs.Schur(S.T[:, :, b], S.Z[:, :, b], S.values[:, b])
            _vectest(A[:, :, b], Sb, VR[:, :, b], VL[:, :, b], 20, f"random n={n} b={b}")
            for V in (VR[:, :, b], VL[:, :, b]):
                np.testing.assert_allclose(np.linalg.norm(V, axis=0), 1.0, rtol=0, atol=1e-13)
                i0 = np.argmax(np.abs(V), axis=0)
                big = V[i0, np.arange(n)]
                assert np.all(big.imag == 0) and np.all(big.real > 0)


def test_vectest_reference_classes(gs):
    """the matrix classes of test/complex.jl (latme / latmr / latms generators, same iseed) through schur + eigvecs;
    tolerance as the reference's vtol for its general classes"""
    for name, A, tol in reference_classes(True):
        n = A.shape[0]
        nA = fnorm(A)
        if nA < 16 * n * np.finfo(float).tiny / ULP or nA > 1e150:
            continue     # the reference rescales these before geigvecs (test/testfuncs.jl:9-17); covered by the scaled test
        S = gs.gschur(np.asfortranarray(A))
        VR = gs.geigvecs(S)
        VL = gs.geigvecs(S, left=True)
        _vectest(A, S, VR, VL, tol, name)       # vectest(A, S, tol) as in test/complex.jl:32


def test_against_lapack_triangular_solves(gs):
    """The triangular systems of src/vectors.jl:79-108, 405-436 solved by LAPACK (ZTRTRS through
    scipy.linalg.solve_triangular — scipy has no ZTREVC wrapper) on the same T, Z, normalised with the reference's
    _enormalize!"""
    from scipy.linalg import solve_triangular
    rng = np.random.default_rng(7)
    n = 24
    A = np.asfortranarray(rng.random((n, n)) + 1j * rng.random((n, n)))
    S = gs.gschur(A)
    VR = gs.geigvecs(S)
    VL = gs.geigvecs(S, left=True)
    T, Z = S.T, S.Z
    vr = np.zeros((n, n), dtype=complex)
    vl = np.zeros((n, n), dtype=complex)
    for k in range(n):
        lam = T[k, k]
        x = solve_triangular(T[:k, :k] - lam * np.eye(k), -T[:k, k]) if k > 0 else np.zeros(0, dtype=complex)
        vr[:, k] = Z[:, k] + Z[:, :k] @ x
        m = n - k - 1
        y = (solve_triangular((T[k + 1:, k + 1:] - lam * np.eye(m)).conj().T, -np.conj(T[k, k + 1:]), lower=True)
             if m > 0 else np.zeros(0, dtype=complex))
        vl[:, k] = Z[:, k] + Z[:, k + 1:] @ y
    np.testing.assert_allclose(VR, enormalize(vr), rtol=0, atol=1e-11)
    np.testing.assert_allclose(VL, enormalize(vl), rtol=0, atol=1e-11)
    # eigenvectors of T itself (no Z), raw scaling of _geigvecs!: largest abs1 component equal to one
    St = gs.Schur(S.T, np.zeros((0, 0), dtype=complex), S.values)
    X = gs.geigvecs(St, normalize=False)
    assert np.allclose(np.max(np.abs(X.real) + np.abs(X.imag), axis=0), 1.0, atol=1e-14)
    assert np.all(np.tril(X, -1) == 0)
    res = fnorm(S.T @ X - X * S.values[None, :]) / (n * fnorm(S.T) * ULP)
    assert res < 20, res


def test_overflow_guarded_solve(gs):
    """a triangular matrix whose plain back substitution overflows (huge off-diagonal part, clustered diagonal): the
    scaled path of _usolve! / _cusolve! (src/util.jl:194-298, 369-455) must return finite, accurate vectors"""
    rng = np.random.default_rng(11)
    n = 40
    T = np.triu(rng.random((n, n)) + 1j * rng.random((n, n)), 1) * 1e200
    T = T + np.diag((1.0 + 1e-8 * np.arange(n)) * (1 + 0.5j))
    T = np.asfortranarray(T)
    St = gs.Schur(T, np.zeros((0, 0), dtype=complex), np.diag(T).copy())
    for left in (False, True):
        X = gs.geigvecs(St, left=left)
        assert np.all(np.isfinite(X))
        M = T.conj().T if left else T
        lam = np.conj(np.diag(T)) if left else np.diag(T)
        # column-wise residual relative to ||T|| ||x|| (the vectors have unit norm)
        R = M @ X - X * lam[None, :]
        assert np.max(np.linalg.norm(R, axis=0)) / (fnorm(T) * n * ULP) < 100


def test_eigvecs_errors(gs):
    S = gs.gschur(np.asfortranarray(np.random.default_rng(0).random((4, 4))))
    with pytest.raises(gs.ArgumentError):
        gs.geigvecs(S)       # Float64: not built (the reference goes through triangularize)
