"""CPU tests pinning the oracle's balance! (src/balance.jl:33-199) against what the reference's test/balance.jl holds and
against LAPACK's xGEBAL, the routine the reference says it translates (scipy.linalg.matrix_balance)."""
import numpy as np
import scipy.linalg as sl

import pytest


def unbal_classic():
    """test/balance.jl:27-40"""
    m = round(-2 * np.log2(np.finfo(float).eps) / 5)
    x = 2.0 ** m
    A = np.array([[1, 0, x ** -2], [1, 1, x ** -1], [x ** 2, x, 1]])
    rt2 = np.sqrt(2.0)
    lam = np.array([1 - rt2 + 1 / (4 * x), 1 - 1 / (2 * x), 1 + rt2 + 1 / (4 * x)])
    return A, lam


def test_classic_unbalanced_matrix(O):
    """test/balance.jl:56-62: the worst unbalanced eigenvalue error is much larger than the worst balanced one"""
    A, lam = unbal_classic()
    Ab, D, sp, ii, rc = O.balance(A)
    assert rc == 0 and ii[:2] == (1, 3)
    Tu = sl.schur(A, output="complex")[0]
    Tb = sl.schur(Ab, output="complex")[0]
    ru = np.max(np.abs(np.sort(np.diag(Tu).real) - lam))
    rb = np.max(np.abs(np.sort(np.diag(Tb).real) - lam))
    assert ru > 100 * rb
    assert np.all(np.log2(D) == np.round(np.log2(D)))        # powers of two


def test_permutation_sanity(O):
    """test/balance.jl:74-108: block form [T1 X Y; 0 C Z; 0 0 T2] after the permutations"""
    rng = np.random.default_rng(3)
    for n1 in range(3):
        for n2 in range(3):
            for n3 in range(3):
                n = n1 + n2 + n3
                if n == 0:
                    continue
                A = np.zeros((n, n))
                A[:n1, :n1] = np.triu(rng.random((n1, n1)))
                A[n1:n1 + n2, n1:n1 + n2] = rng.random((n2, n2))
                A[n1 + n2:, n1 + n2:] = np.triu(rng.random((n3, n3)))
                ic = rng.permutation(n)
                Pc = np.zeros((n, n))
                Pc[np.arange(n), ic] = 1
                C, D, sp, (ilo, ihi, trivial), rc = O.balance(Pc.T @ A @ Pc)
                assert rc == 0 and ilo <= ihi
                nsub = sum(1 for i in range(1, n + 1) for j in range(1, i) if (i < ilo or j > ihi) and C[i - 1, j - 1] != 0)
                assert nsub == 0


def test_against_lapack_gebal(O):
    rng = np.random.default_rng(0)
    for complex_ in (False, True):
        for n in (5, 12, 33):
            A = rng.random((n, n)) * np.exp(rng.normal(0, 8, (n, 1))) / np.exp(rng.normal(0, 8, (1, n)))
            if complex_:
                A = A * np.exp(2j * np.pi * rng.random((n, n)))
            Ab, D, sp, ii, rc = O.balance(A)
            B, (sc, perm) = sl.matrix_balance(A, permute=True, scale=True, separate=True)
            assert rc == 0
            np.testing.assert_array_equal(D, sc)
            np.testing.assert_array_equal(Ab, B)


def test_balancer_back_transformation(O):
    """lmul!(B, V) maps eigenvectors of the balanced matrix to eigenvectors of A (src/balance.jl:203-228)"""
    rng = np.random.default_rng(5)
    n = 9
    A = rng.random((n, n)) * np.exp(rng.normal(0, 6, (n, 1))) / np.exp(rng.normal(0, 6, (1, n)))
    A[3, :3] = 0
    A[3, 4:] = 0          # an isolated eigenvalue: exercises the permutation part
    Ab, D, sp, ii, rc = O.balance(A)
    w, V = np.linalg.eig(Ab)
    V = O.balance_apply(V.astype(complex), D, sp, ii)
    assert np.linalg.norm(A @ V - V * w[None, :]) / (np.linalg.norm(A) * np.linalg.norm(V)) < 1e-12
