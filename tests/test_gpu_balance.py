"""GPU tests of the balancing pre-step, its back-transformation and triangularize (-m gpu).
The CUDA kernels must agree with the oracle bit for bit: permutations and power-of-two scalings are exact operations
and the order of the floating-point operations behind the decisions is fixed on both sides."""
import numpy as np
import pytest

from common import ULP, fnorm
from test_balance_oracle import unbal_classic

pytestmark = pytest.mark.gpu


def _badly_scaled(rng, n, batch, complex_):
    A = rng.random((n, n, batch)) * np.exp(rng.normal(0, 8, (n, 1, batch))) / np.exp(rng.normal(0, 8, (1, n, batch)))
    if complex_:
        A = A * np.exp(2j * np.pi * rng.random((n, n, batch)))
    return np.asfortranarray(A)


@pytest.mark.parametrize("complex_", [False, True])
def test_balance_matches_oracle_exactly(gs, O, complex_):
    rng = np.random.default_rng(17 + complex_)
    for n, batch in ((3, 40), (12, 64), (33, 16), (64, 8)):
        A = _badly_scaled(rng, n, batch, complex_)
        # isolate some eigenvalues in a few matrices (rows / columns that are zero off the diagonal)
        for b in range(0, batch, 3):
            k = int(rng.integers(n))
            d = A[k, k, b]
            A[k, :, b] = 0
            A[k, k, b] = d
            if n > 4:
                k2 = int(rng.integers(n))
                if k2 != k:
                    d = A[k2, k2, b]
                    A[:, k2, b] = 0
                    A[k2, k2, b] = d
        Ab, B = gs.balance(A)
        for b in range(batch):
            Ao, D, sp, (ilo, ihi, trivial), rc = O.balance(A[:, :, b])
            assert rc == 0
            assert (int(B.ilo[b]), int(B.ihi[b]), bool(B.trivial[b])) == (ilo, ihi, trivial), (n, b)
            np.testing.assert_array_equal(B.D[:, b], D)
            np.testing.assert_array_equal(B.perm[:, b], sp)
            np.testing.assert_array_equal(Ab[:, :, b], Ao)
    # single matrix, permute / scale switches
    A1 = np.asfortranarray(_badly_scaled(rng, 10, 1, complex_)[:, :, 0])
    for kw in ({"scale": False}, {"permute": False}, {"scale": False, "permute": False}):
        Ab, B = gs.balance(A1, **kw)
        Ao, D, sp, ii, rc = O.balance(A1, **kw)
        np.testing.assert_array_equal(Ab, Ao)
        assert (B.ilo, B.ihi, B.trivial) == ii


def test_classic_unbalanced_eigen(gs):
    """test/balance.jl:56-71: eigen with balancing is far more accurate than without on the classic example, and the
    back-transformed left eigenvectors satisfy A' Vl = Vl conj(Lambda)"""
    A, lam = unbal_classic()
    Ac = np.asfortranarray(A.astype(complex))
    wu, vu = gs.eigen_(Ac.copy(order="F"), permute=False, scale=False)
    wb, vb = gs.eigen_(Ac.copy(order="F"))
    ru = np.max(np.abs(np.sort(wu.real) - lam))
    rb = np.max(np.abs(np.sort(wb.real) - lam))
    assert ru > 100 * rb, (ru, rb)
    assert fnorm(Ac @ vb - vb * wb[None, :]) / (fnorm(Ac) * 3 * ULP) < 100
    # real input goes through triangularize
    wr, vr = gs.eigen_(np.asfortranarray(A.copy()))
    assert np.max(np.abs(np.sort(wr.real) - lam)) < 100 * max(rb, 1e-15)
    Abal, B = gs.balance(Ac)
    S = gs.gschur(Abal)
    Vl = gs.geigvecs(S, left=True, normalize=False)
    gs.balancer_lmul_(B, Vl, inverse=True)
    lhs, rhs = Ac.conj().T @ Vl, Vl * np.conj(S.values)[None, :]
    assert np.allclose(lhs, rhs, rtol=1e-8, atol=1e-8 * np.abs(lhs).max())


def test_eigen_batched_badly_scaled(gs):
    rng = np.random.default_rng(23)
    n, batch = 16, 32
    A = _badly_scaled(rng, n, batch, True)
    w, V = gs.eigen_(A.copy(order="F"))
    for b in range(batch):
        # column-wise residual relative to ||A|| (unit-norm vectors): balancing keeps it at the rounding level of the
        # BALANCED problem; allow for the condition of the diagonal similarity
        R = A[:, :, b] @ V[:, :, b] - V[:, :, b] * w[:, b][None, :]
        assert np.all(np.isfinite(V[:, :, b]))
        np.testing.assert_allclose(np.linalg.norm(V[:, :, b], axis=0), 1.0, atol=1e-12)
        ref = np.linalg.eigvals(A[:, :, b])
        assert np.max(np.min(np.abs(w[:, b][:, None] - ref[None, :]), axis=1) / np.abs(ref).max()) < 1e-9


def test_triangularize(gs, O):
    """src/triang.jl:9-43 and test/real.jl:63-73: the complex triangular form of a real Schur decomposition"""
    rng = np.random.default_rng(29)
    n, batch = 24, 6
    A = np.asfortranarray(rng.random((n, n, batch)) - 0.5)
    S = gs.gschur(A)
    Sc = gs.triangularize(S)
    for b in range(batch):
        T, Z = Sc.T[:, :, b], Sc.Z[:, :, b]
        assert np.all(np.tril(T, -1) == 0)
        np.testing.assert_array_equal(np.diag(T), Sc.values[:, b])
        berr, oerr, _ = O.residuals(A[:, :, b].astype(complex), T, Z, 1)
        assert berr < 10 and oerr < 10, (berr, oerr)
        ref = S.values[:, b]
        assert np.max(np.min(np.abs(Sc.values[:, b][:, None] - ref[None, :]), axis=1)) < 1e-12 * np.abs(ref).max()
