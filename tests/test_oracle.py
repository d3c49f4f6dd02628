"""CPU tests of the oracle: pinned against every known-answer fixture the reference's tests hold for the hot path,
against the LAPACK Fortran originals, and against the committed golden file (regression pin)."""
import numpy as np
import pytest

from common import (MAGNS, ULP, checkblocks, csort, fnorm, godunov, match_eigs, reference_classes, structure_ok)


def test_golden_regression_bit_exact(O, golden):
    """The oracle reproduces its committed outputs bit for bit (same compiler flags: no contraction, -O2)."""
    for key in golden["names"]:
        key = str(key)
        A = golden[key + "__A"]
        kind = int(golden[key + "__meta"][0])
        T, Z, w, rc, st = O.gschur(A, kind)
        assert rc == 0
        assert np.array_equal(w, golden[key + "__w"]), key


def test_golden_vs_lapack_originals(O, golden):
    """Eigenvalues agree with xGEHD2+xLAHQR (the routines the reference says it translates) within a
    condition-scaled tolerance: |dlambda| <= 1e3 * ulp * ||A|| / s_i."""
    checked = 0
    for key in golden["names"]:
        key = str(key)
        A = golden[key + "__A"]
        w = golden[key + "__w"]
        wl = golden[key + "__wlapack"]
        if np.any(np.isnan(wl)):
            continue
        # condition numbers from the complex Schur form of A computed by the oracle, on a copy scaled to unit
        # max-norm (the reference's vectest does the same for tiny inputs, test/testfuncs.jl:9-17)
        sc = float(np.max(np.abs(A))) or 1.0
        Ac = (A / sc).astype(np.complex128)
        Tc, _, wc, rc, _ = O.gschur(Ac, 1)
        assert rc == 0
        s = O.eigvalscond(Tc, 1)
        s = np.where(np.isfinite(s) & (s > 0), s, 1e-300)   # defective (Jordan) blocks: no bound
        anorm = fnorm(Ac)
        tol = 1e3 * ULP * anorm / s
        d = match_eigs(wl / sc, wc, tol)
        assert np.all(d <= tol), (key, float(np.max(d / tol)))
        d2 = match_eigs(w / sc, wc, tol)
        assert np.all(d2 <= tol), (key, "real-vs-complex path", float(np.max(d2 / tol)))
        checked += 1
    assert checked > 60


@pytest.mark.parametrize("complex_", [False, True])
def test_reference_classes_acceptance(O, complex_):
    """schurtest of test/complex.jl:1-34 / test/real.jl:24-74 on the reference's matrix classes (n in {4, 32})."""
    kind = 1 if complex_ else 0
    for name, A, tol in reference_classes(complex_):
        T, Z, w, rc, st = O.gschur(A, kind)
        assert rc == 0, name
        ok, why = structure_ok(T, w, kind, tol)
        assert ok, (name, why)
        berr, oerr, anorm = O.residuals(A, T, Z, kind)
        assert berr < tol and oerr < tol, (name, berr, oerr)
        if name.startswith("normal"):
            off = fnorm(np.triu(T, 1 if complex_ else 2)) / (A.shape[0] * fnorm(A) * ULP)
            assert off < tol, (name, off)


def test_godunov_known_eigenvalues(O):
    """test/real.jl:142-157, test/complex.jl:139-154: eigenvalues {-4,-2,-1,0,1,2,4}, condition 7e16; needs > 70 bits,
    so it is run in double-double and in MPFR-256."""
    G, vals, econd = godunov()
    # double-double, real path
    Gd = O.dd_from_float(G)
    T, Z, w, rc, _ = O.gschur(Gd, 2)
    assert rc == 0
    berr, oerr, anorm = O.residuals(Gd, T, Z, 2)
    eps_dd = 2.0 ** -104
    delta = berr * 7 * anorm * eps_dd
    assert delta < 100 * eps_dd * anorm                       # δ < 100 eps ‖A‖
    wv = (w[0] + w[1]) + 1j * (w[2] + w[3])
    assert np.allclose(csort(wv), vals, atol=3 * delta * econd)
    # double-double, complex path
    Gc = np.zeros((4, 7, 7), order="F")
    Gc[0] = G
    T, Z, w, rc, _ = O.gschur(Gc, 3)
    assert rc == 0
    berr, _, _ = O.residuals(Gc, T, Z, 3)
    delta = berr * 7 * anorm * eps_dd
    wv = (w[0] + w[1]) + 1j * (w[2] + w[3])
    assert np.allclose(csort(wv), vals, atol=3 * delta * econd)
    # MPFR-256: essentially exact
    Tm, Zm, wm, rc = O.gschur_mp(G, 0)
    assert rc == 0
    wv = (wm[0] + wm[1]) + 1j * (wm[2] + wm[3])
    assert np.abs(csort(wv) - vals).max() < 1e-40 * econd


def test_gs2x2_nearly_degenerate(O):
    """test/real.jl:311-317"""
    rng = np.random.default_rng(5)
    for _ in range(20):
        l1, l2 = 1 + 2 * ULP, 1 - 2 * ULP
        B = np.diag([l1, l2]) + (ULP / 4) * rng.random((2, 2))
        abcd, csn, w1, w2 = O.gs2x2(B[0, 0], B[0, 1], B[1, 0], B[1, 1])
        assert w1.imag == 0 and w2.imag == 0
        assert abs(max(w1.real, w2.real) - l1) < 2 * ULP
        assert abs(min(w1.real, w2.real) - l2) < 2 * ULP
        assert abs(csn[0] ** 2 + csn[1] ** 2 - 1) < 4 * ULP


def test_gs2x2_matches_dlanv2(O):
    import tmg
    rng = np.random.default_rng(11)
    cases = [rng.standard_normal(4) for _ in range(200)]
    cases += [np.array([1.0, 0.0, 2.0, 3.0]), np.array([1.0, 2.0, 0.0, 3.0]), np.array([1.0, 2.0, -2.0, 1.0]),
              np.array([1.0, 1e-200, 1e200, 1.0]), np.array([2.0, 1.0, 1.0, 2.0])]
    for c in cases:
        abcd, csn, w1, w2 = O.gs2x2(*c)
        (la, lb, lc, ld), (r1r, r1i, r2r, r2i, cs, sn) = tmg.lapack_lanv2(*c)
        # same algorithm up to the safe-scaling refinements newer LAPACK added: compare to a few ulp
        np.testing.assert_allclose(abcd, [la, lb, lc, ld], rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose([w1.real, w1.imag, w2.real, w2.imag], [r1r, r1i, r2r, r2i], rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(csn, [cs, sn], rtol=1e-13, atol=1e-300)


def test_reflector_properties(O):
    """_reflector! (src/householder.jl:12-102): H*x = beta*e1, beta real, H unitary; tiny inputs take the rescaling loop."""
    rng = np.random.default_rng(3)
    for complex_ in (False, True):
        for n in (1, 2, 3, 10):
            # (complex subnormal input overflows 1/w inside _hypot3, src/util.jl:562-570, in the reference too;
            #  _scale! keeps the hot path away from it, so subnormals are exercised on the real variant only)
            for scale in ((1.0, 1e-300, 1e300) if complex_ else (1.0, 1e-300, 1e-310, 1e300)):
                x = rng.standard_normal(n) + (1j * rng.standard_normal(n) if complex_ else 0)
                x = x * scale
                y, tau = O.reflector(x)
                v = np.concatenate([[1.0], y[1:]])
                Hm = np.eye(n) - tau * np.outer(v, v.conj())
                beta = y[0]
                assert abs(np.imag(beta)) == 0
                r = Hm.conj().T @ x
                nx = float(np.max(np.abs(x))) * np.sqrt(n)
                assert abs(r[0] - beta) <= 8 * ULP * nx + 1e-323
                if n > 1:
                    assert np.abs(r[1:]).max() <= 8 * ULP * nx + 1e-323
                assert np.abs(Hm @ Hm.conj().T - np.eye(n)).max() < 16 * ULP
    # real: n == 1 or zero tail -> tau = 0, x untouched
    y, tau = O.reflector(np.array([3.0]))
    assert tau == 0 and y[0] == 3.0
    y, tau = O.reflector(np.array([3.0, 0.0, 0.0]))
    assert tau == 0 and y[0] == 3.0


def test_hessenberg_matches_gehd2(O):
    """hesstest (test/complex.jl:36-61, test/real.jl:76-99) + agreement with LAPACK xGEHD2 / sub-diagonal real."""
    import tmg
    rng = np.random.default_rng(1234)
    for complex_ in (False, True):
        for scale in (1.0, 100 * np.finfo(float).tiny, np.finfo(float).max / 100):
            if complex_ and scale > 1:
                continue
            A = rng.random((32, 32)) + (1j * rng.random((32, 32)) if complex_ else 0)
            A = np.asfortranarray(A * scale)
            kind = 1 if complex_ else 0
            F, tau, Q = O.hessenberg(A, kind)
            Hm = np.triu(F, -1)
            berr, oerr, _ = O.residuals(A, Hm, Q, kind)
            assert berr < 10 and oerr < 10
            assert np.all(np.imag(np.diag(Hm, -1)) == 0)
            Fl, taul = tmg.lapack_gehd2(A)
            if complex_:
                # zgehd2 leaves the last sub-diagonal complex; compare moduli of H and the leading taus
                np.testing.assert_allclose(np.abs(np.triu(F, -1)), np.abs(np.triu(Fl, -1)), rtol=0,
                                           atol=1e-12 * np.abs(A).max())
            else:
                np.testing.assert_allclose(np.triu(F, -1), np.triu(Fl, -1), rtol=0, atol=1e-12 * np.abs(A).max())
                np.testing.assert_allclose(np.tril(F, -2), np.tril(Fl, -2), rtol=0, atol=1e-12)   # reflector tails
                np.testing.assert_allclose(tau, taul, rtol=0, atol=1e-13)


def test_error_paths(O):
    """test/errors.jl:1-14: complex sub-diagonal -> ArgumentError (rc -2); NaN input -> UnconvergedException (rc 1)."""
    n = 5
    rng = np.random.default_rng(0)
    A = np.diag(np.full(n - 1, -1.0 + 1.0j), -1) + np.triu(rng.random((n, n)))
    _, _, _, rc = O.gschur_hess(A, 1, checksd=True)
    assert rc == -2
    B = rng.random((6, 6))
    B[2, 3] = np.nan
    _, _, _, rc, _ = O.gschur(B, 0)
    assert rc == 1
    _, _, _, rc, _ = O.gschur(B.astype(np.complex128), 1)
    assert rc == 1


def test_scaling_branches(O):
    """_scale! (src/util.jl:14-29) triggers for the ovfl*ulp / unfl/ulp magnitude classes and is undone exactly enough."""
    rng = np.random.default_rng(2)
    for mag in (MAGNS[1], MAGNS[2], 1e-200, 1e250):
        for kind in (0, 1):
            A = rng.random((12, 12)) + (1j * rng.random((12, 12)) if kind else 0)
            A = np.asfortranarray(A * mag)
            T, Z, w, rc, _ = O.gschur(A, kind)
            assert rc == 0
            berr, oerr, _ = O.residuals(A, T, Z, kind)
            assert berr < 10 and oerr < 10
            T2, _, w2, rc, _ = O.gschur(A, kind, scale=False)
            if rc == 0 and 1e-100 < mag < 1e100:
                assert np.array_equal(w, w2)


def test_dd_and_mp_agree(O):
    """The double-double instantiation agrees with the MPFR-256 one to double-double accuracy."""
    rng = np.random.default_rng(9)
    n = 12
    A = np.zeros((4, n, n), order="F")
    A[0] = rng.random((n, n))
    A[2] = rng.random((n, n))
    T, Z, w, rc, _ = O.gschur(A, 3)
    Tm, Zm, wm, rc2 = O.gschur_mp(A, 3)
    assert rc == 0 and rc2 == 0
    berr, oerr, _ = O.residuals(A, T, Z, 3)
    assert berr < 10 and oerr < 10
    berr, oerr, _ = O.residuals(A, Tm, Zm, 3)
    assert berr < 0.1 and oerr < 0.1              # MPFR result rounded to dd: only the final rounding is left
    # eigenvalues: differences in the low limbs only
    wd_hi = w[0] + 1j * w[2]
    wm_hi = wm[0] + 1j * wm[2]
    s = O.eigvalscond(np.asfortranarray(T), 3)
    order_d = np.lexsort((wd_hi.imag, wd_hi.real))
    order_m = np.lexsort((wm_hi.imag, wm_hi.real))
    dre = (w[0][order_d] - wm[0][order_m]) + (w[1][order_d] - wm[1][order_m])
    dim = (w[2][order_d] - wm[2][order_m]) + (w[3][order_d] - wm[3][order_m])
    err = np.hypot(dre, dim)
    tol = 100 * 2.0 ** -104 * np.linalg.norm(A[0] + 1j * A[2]) / s[order_d]
    assert np.all(err <= tol), float(np.max(err / tol))


def test_batched_threads(O):
    rng = np.random.default_rng(4)
    A = np.asfortranarray(rng.random((8, 8, 10)))
    T1, Z1, w1, i1 = O.gschur_batched(A.copy(order="F"), 0, nthreads=1)
    T2, Z2, w2, i2 = O.gschur_batched(A.copy(order="F"), 0, nthreads=4)
    assert np.array_equal(T1, T2) and np.array_equal(w1, w2) and not i1.any() and not i2.any()
