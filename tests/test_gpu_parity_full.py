"""GPU parity tests at BASELINE.json's FULL sizes and for the cases round 1 left to the oracle only (run with -m gpu).

  * config 3 (65 536 x 64x64 ComplexF64) and config 5 (4 096 x 96x96 Complex double-double): size-independent
    invariants on the WHOLE batch, plus 32 randomly drawn matrices checked in full against the oracle (MPFR-256
    residuals, eigenvalues against the oracle's / the MPFR-256 decomposition's within the eigvalscond-scaled bound);
  * the reference-held nearly-degenerate 2x2 of test/real.jl:311-317 through the CUDA path;
  * the double-double kinds over the reference's matrix classes, extreme magnitudes, Hessenberg, wantZ = false;
  * the in-process multi-slice host path on one GPU;
  * config 4: MPFR-256 residuals at n = 512, eigenvalues against LAPACK at n = 4096.

Tolerances are those of tests/test_gpu_parity.py (the reference's own: test/complex.jl:1-34, test/real.jl:24-74).
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from common import ULP, csort, fnorm, reference_classes
from test_gpu_parity import _bulk_properties, _check_one

pytestmark = pytest.mark.gpu

EPS_DD = 2.0 ** -104


def _pmap(fn, items):
    """The oracle is ctypes-bound C++ (the GIL is released inside): run the independent checks on all host cores."""
    with ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 2))) as ex:
        return list(ex.map(fn, items))


def test_cfg3_full_batch(gs, O):
    """BASELINE config 3 at full size: 65 536 random 64x64 ComplexF64 matrices on one GPU."""
    rng = np.random.default_rng(1234 + 3)
    n, batch = 64, 65536
    A = np.empty((n, n, batch), dtype=np.complex128, order="F")
    step = 8192
    for b0 in range(0, batch, step):        # generated in slices: the same values as one draw of the trailing slices
        A[:, :, b0:b0 + step] = np.asfortranarray(rng.random((n, n, step)) + 1j * rng.random((n, n, step)))
    S = gs.gschur(A)
    _bulk_properties(A, S, 1)
    assert np.array_equal(S.values, np.diagonal(S.T, axis1=0, axis2=1).T)          # values == diag(T), exactly
    pick = np.random.default_rng(7).choice(batch, size=32, replace=False)

    def one(b):
        _check_one(O, A[:, :, b], S.T[:, :, b], S.Z[:, :, b], S.values[:, b], 1, 10, f"cfg3[{b}]")
        return True

    assert all(_pmap(one, [int(b) for b in pick]))


def _dd_batch(rng, lead, n, batch):
    A = np.zeros((lead, n, n, batch), order="F")
    for part in range(0, lead, 2):
        hi = rng.random((n, n, batch))
        lo = (rng.random((n, n, batch)) - 0.5) * 2.0 ** -53 * hi
        s = hi + lo
        A[part] = s
        A[part + 1] = lo - (s - hi)
    return A


def _dd_eig_check(O, gs, Ab, Tb, wb, kind, name):
    """Eigenvalues of a double-double result against the MPFR-256 decomposition, limb by limb, within
    1e3 eps_dd ||A|| / s_i (s_i from the complex double-double Schur form)."""
    from scipy.optimize import linear_sum_assignment
    n = Ab.shape[-1]
    Tm, Zm, wm, rc = O.gschur_mp(Ab, kind)
    assert rc == 0, name
    if kind == gs.CDD:
        Tc = np.asfortranarray(Tb)
    else:
        Ac = np.zeros((4, n, n), order="F")
        Ac[0], Ac[1] = Ab[0], Ab[1]
        Tc, _, _, rc, _ = O.gschur(Ac, 3)
        assert rc == 0
    s = O.eigvalscond(Tc, 3)
    wg = wb[0] + 1j * wb[2]
    wr = wm[0] + 1j * wm[2]
    r, c = linear_sum_assignment(np.abs(wg[None, :] - wr[:, None]))
    dre = (wb[0][c] - wm[0][r]) + (wb[1][c] - wm[1][r])
    dim = (wb[2][c] - wm[2][r]) + (wb[3][c] - wm[3][r])
    err = np.hypot(dre, dim)
    if kind == gs.CDD:
        sc = s[c]
    else:
        wc = np.array([Tc[0, i, i] + 1j * Tc[2, i, i] for i in range(n)])
        rr, cc = linear_sum_assignment(np.abs(wc[None, :] - wg[c][:, None]))
        sc = s[cc]
    anorm = fnorm(Ab[0] + (1j * Ab[2] if kind == gs.CDD else 0))
    tol = 1e3 * EPS_DD * anorm / np.maximum(sc, 1e-300)
    ok = sc >= 1e-3            # first-order bound only for well-conditioned eigenvalues (see test_gpu_parity._eig_tol)
    assert np.all(err[ok] <= tol[ok]), (name, float(np.max(err[ok] / tol[ok])))


def test_cfg5_full_batch(gs, O):
    """BASELINE config 5 at full size: 4 096 random 96x96 Complex double-double matrices; 32 of them against the
    MPFR-256 (BigFloat(256) stand-in) residuals and decomposition.  Tolerance 20 with eps = 2^-104, see
    test_gpu_parity.test_double_double_vs_bigfloat."""
    rng = np.random.default_rng(1234 + 5)
    n, batch = 96, 4096
    A = _dd_batch(rng, 4, n, batch).view(gs.CDDArray)
    S = gs.gschur(A)
    assert not np.any(S.info)
    T = np.asarray(S.T)
    w = np.asarray(S.values)
    ii, jj = np.tril_indices(n, -1)
    assert not np.any(T[:, ii, jj, :])                                  # exactly upper triangular, every limb
    assert np.array_equal(w, np.diagonal(T, axis1=1, axis2=2).transpose(0, 2, 1))   # values == diag(T), every limb
    Ahi = np.asarray(A[0]) + 1j * np.asarray(A[2])
    whi = (w[0] + w[1]) + 1j * (w[2] + w[3])
    assert np.allclose(whi.sum(axis=0), np.trace(Ahi, axis1=0, axis2=1), rtol=0, atol=1e-11 * n)
    fa = np.sqrt((np.abs(Ahi) ** 2).sum(axis=(0, 1)))
    Thi = (T[0] + T[1]) + 1j * (T[2] + T[3])
    assert np.allclose(fa, np.sqrt((np.abs(Thi) ** 2).sum(axis=(0, 1))), rtol=1e-12)
    Z = np.asarray(S.Z)
    Zhi = (Z[0] + Z[1]) + 1j * (Z[2] + Z[3])
    assert np.allclose((np.abs(Zhi) ** 2).sum(axis=(0, 1)), n, rtol=1e-12)
    pick = np.random.default_rng(8).choice(batch, size=32, replace=False)

    def one(b):
        Ab = np.asfortranarray(np.asarray(A[..., b]))
        Tb, Zb, wb = np.asfortranarray(T[..., b]), np.asfortranarray(Z[..., b]), w[..., b]
        berr, oerr, _ = O.residuals(Ab, Tb, Zb, 3)
        assert berr < 20 and oerr < 20, (b, berr, oerr)
        _dd_eig_check(O, gs, Ab, Tb, wb, gs.CDD, f"cfg5[{b}]")
        return True

    assert all(_pmap(one, [int(b) for b in pick]))


def test_gs2x2_fixture_on_gpu(gs):
    """test/real.jl:311-317 ("tiny, almost degenerate") through the CUDA path: a 2x2 matrix is deflated at once by the
    device's 2x2 standardisation (_gs2x2!, src/GenericSchur.jl:716-803), so gschur returns exactly what the reference's
    fixture calls _gs2x2! for.  Real eigenvalues (imag == 0 exactly), each within 2 ulp of its target."""
    rng = np.random.default_rng(5)
    l1, l2 = 1 + 2 * ULP, 1 - 2 * ULP
    B = np.zeros((2, 2, 64), order="F")
    for b in range(64):
        B[:, :, b] = np.diag([l1, l2]) + (ULP / 4) * rng.random((2, 2))
    S = gs.gschur(B)
    assert not np.any(S.info)
    w = S.values
    assert np.all(w.imag == 0)
    assert np.all(np.abs(w.real.max(axis=0) - l1) < 2 * ULP)
    assert np.all(np.abs(w.real.min(axis=0) - l2) < 2 * ULP)
    assert np.all(S.T[1, 0, :] == 0)                                       # real pair: upper triangular
    assert np.array_equal(S.T[0, 0, :], w[0].real) and np.array_equal(S.T[1, 1, :], w[1].real)
    zz = np.einsum("ijb,ikb->jkb", S.Z, S.Z)
    assert np.abs(zz - np.eye(2)[:, :, None]).max() < 4 * ULP
    # a single matrix through the single-matrix entry as well
    S1 = gs.gschur(np.asfortranarray(B[:, :, 0]))
    assert np.array_equal(S1.values, w[:, 0])


def _to_dd(A, gs):
    """Float64 / ComplexF64 matrix -> double-double array with zero low limbs."""
    if np.iscomplexobj(A):
        X = np.zeros((4,) + A.shape, order="F")
        X[0], X[2] = A.real, A.imag
        return X.view(gs.CDDArray), gs.CDD
    X = np.zeros((2,) + A.shape, order="F")
    X[0] = A
    return X.view(gs.DDArray), gs.DD


@pytest.mark.parametrize("complex_", [False, True])
def test_double_double_reference_classes(gs, O, complex_):
    """The double-double kinds over the reference's matrix classes (test/complex.jl:221-389, test/real.jl:180-309:
    Jordan, latme / latmr at magnitudes 1, ovfl*ulp, unfl/ulp, diagonal, symmetric, triangular, unitary): residual
    ratios in MPFR-256 normalised by eps = 2^-104, tolerance 20; exact structure; eigenvalues against MPFR-256 for
    the well-conditioned ones."""
    for name, A, tol in reference_classes(complex_, sizes=(4, 24)):
        if "tiny" in name or "huge" in name:
            continue      # magnitudes outside the double-double range (floatmin = 2^-969) are covered by the scaling test
        X, kind = _to_dd(A, gs)
        S = gs.gschur(X, check=False)
        assert S.info == 0, name
        T, Z, w = np.asarray(S.T), np.asarray(S.Z), np.asarray(S.values)
        n = A.shape[0]
        if complex_:
            ii, jj = np.tril_indices(n, -1)
        else:
            ii, jj = np.tril_indices(n, -2)
        assert not np.any(T[:, ii, jj]), name
        berr, oerr, _ = O.residuals(np.asfortranarray(np.asarray(X)), np.asfortranarray(T), np.asfortranarray(Z), kind)
        assert berr < 20 and oerr < 20, (name, berr, oerr)
        if name.startswith(("rand", "normal", "sym", "diag")):
            _dd_eig_check(O, gs, np.asfortranarray(np.asarray(X)), T, w, kind, name)


def test_double_double_scaling_hessenberg_wantz(gs, O):
    """_scale! round trip at extreme magnitudes, hessenberg!, and the eigvals! path (wantZ = false) in double-double."""
    rng = np.random.default_rng(3)
    n = 12
    for kind, lead in ((gs.DD, 2), (gs.CDD, 4)):
        base = _dd_batch(rng, lead, n, 1)[..., 0]
        for mag in (2.0 ** -900, 2.0 ** -500, 1.0, 2.0 ** 700, 2.0 ** 960):
            A = np.asfortranarray(base * mag).view(gs.DDArray if kind == gs.DD else gs.CDDArray)   # power of two: exact in both limbs
            S = gs.gschur(A)
            assert S.info == 0
            berr, oerr, _ = O.residuals(np.asfortranarray(np.asarray(A)), np.asfortranarray(np.asarray(S.T)),
                                        np.asfortranarray(np.asarray(S.Z)), kind)
            assert berr < 20 and oerr < 20, (kind, mag, berr, oerr)
        A = np.asfortranarray(base).view(gs.DDArray if kind == gs.DD else gs.CDDArray)
        S1 = gs.gschur(A)
        S2 = gs.gschur(A, wantZ=False)
        assert S2.Z.shape == (0, 0)
        assert np.array_equal(np.asarray(S1.T), np.asarray(S2.T)) and np.array_equal(np.asarray(S1.values), np.asarray(S2.values))
        # hesstest (test/complex.jl:36-61, test/real.jl:76-99) in double-double: A = Q H Q', Q unitary, H Hessenberg,
        # real sub-diagonal for the complex kind
        Hs = gs.hessenberg(A)
        F = np.asarray(Hs.factors)
        Hm = F.copy()
        ii, jj = np.tril_indices(n, -2)
        Hm[:, ii, jj] = 0.0
        berr, oerr, _ = O.residuals(np.asfortranarray(np.asarray(A)), np.asfortranarray(Hm), np.asfortranarray(np.asarray(Hs.Q)), kind)
        assert berr < 20 and oerr < 20, (kind, berr, oerr)
        if kind == gs.CDD:
            sd = np.arange(n - 1)
            assert np.all(Hm[2, sd + 1, sd] == 0) and np.all(Hm[3, sd + 1, sd] == 0)
        Fo, tauo, Qo = O.hessenberg(np.asfortranarray(np.asarray(A)), kind)
        Fo_h = Fo.copy()
        Fo_h[:, ii, jj] = 0.0
        assert np.abs((Hm[0] - Fo_h[0]) + (Hm[1] - Fo_h[1])).max() < 1e-27      # same factorisation as the oracle's, to dd accuracy


def test_in_process_multi_slice(gs, O):
    """The host-pointer path with the batch split in several contiguous slices, each on its own host thread and
    three-stream pipeline — exercised on ONE GPU by naming it several times: bit-identical to the
    single-slice result, for an uneven split and both the three-stage (n <= 64) and the fused (n > 64) kernels."""
    rng = np.random.default_rng(78)
    for n, batch, devs in ((48, 4133, [0] * 5), (33, 1001, [0, 0, 0]), (72, 301, [0, 0])):
        A = np.asfortranarray(rng.random((n, n, batch)) + 1j * rng.random((n, n, batch)))
        S1 = gs.gschur(A, devices=[0])
        S2 = gs.gschur(A, devices=devs)
        assert not np.any(S2.info)
        assert np.array_equal(S1.T, S2.T) and np.array_equal(S1.Z, S2.Z) and np.array_equal(S1.values, S2.values)
        b = batch // 2
        _check_one(O, A[:, :, b], S2.T[:, :, b], S2.Z[:, :, b], S2.values[:, b], 1, 10, f"slice n{n}")


def test_three_stage_matches_fused(gs, O, monkeypatch):
    """The three-stage path (stage B logs the reflectors, stage C replays them on Z) against the fused stage B on the same
    inputs: both meet the acceptance ratios and their eigenvalues agree to rounding; a log that overflows (forced with a
    tiny pool) falls back to the fused kernel and still gives a valid decomposition."""
    from common import match_eigs
    rng = np.random.default_rng(41)
    for kind, n, batch in ((0, 64, 24), (1, 64, 24), (0, 19, 40), (1, 30, 40)):
        A = np.asfortranarray(rng.random((n, n, batch)) + (1j * rng.random((n, n, batch)) if kind else 0))
        monkeypatch.delenv("GSCHUR_QR", raising=False)
        S = gs.gschur(A)
        monkeypatch.setenv("GSCHUR_QR", "fused")
        S0 = gs.gschur(A)
        monkeypatch.delenv("GSCHUR_QR")
        monkeypatch.setenv("GSCHUR_LOG_TEST_TINY", "1")
        S1 = gs.gschur(A)
        monkeypatch.delenv("GSCHUR_LOG_TEST_TINY")
        for R in (S, S0, S1):
            assert not np.any(R.info)
        for b in range(0, batch, 7):
            for R, nm in ((S, "log"), (S1, "redo")):
                _check_one(O, A[:, :, b], R.T[:, :, b], R.Z[:, :, b], R.values[:, b], kind, 10, f"{nm} k{kind} n{n} b{b}")
            d = match_eigs(S.values[:, b], S0.values[:, b])
            assert np.max(d) <= 1e-11 * n * np.abs(A[:, :, b]).max()


def test_cfg4_parity(gs, O):
    """BASELINE config 4 parity beyond the Float64 residuals of test_gpu_parity: MPFR-256 residuals at n = 512 and the
    eigenvalues of the full n = 4096 run against LAPACK (dgeev through numpy), matched optimally."""
    rng = np.random.default_rng(1234 + 4)
    n = 512
    A0 = np.asfortranarray(rng.random((n, n)))
    S = gs.gschur(A0)
    assert S.info == 0
    berr, oerr, _ = O.residuals(A0, S.T, S.Z, 0)
    assert berr < 10 and oerr < 10, (berr, oerr)
    n = 4096
    A0 = np.asfortranarray(rng.random((n, n)))
    S = gs.gschur(A0, wantZ=False)
    assert S.info == 0
    ev = np.linalg.eigvals(A0)
    # nearest-neighbour matching in the complex plane; the map must be a bijection
    from scipy.spatial import cKDTree
    tree = cKDTree(np.c_[ev.real, ev.imag])
    dist, idx = tree.query(np.c_[S.values.real, S.values.imag])
    assert len(np.unique(idx)) == n
    # both solvers are backward stable to n ulp ||A||_2 ~ 1e-9 here; eigenvalue condition numbers of a random matrix of
    # this size reach O(1e2)
    assert dist.max() < 1e-6, float(dist.max())


def test_hessenberg_shares_the_schur_limits(gs, O):
    """Hessenberg-only requests run on the stage A kernel: the same size limits as gschur! (128 for the Float64 kinds, 96
    for the double-double kinds), factors and tau as hessenberg! returns them (src/hessenberg.jl:3-17)."""
    rng = np.random.default_rng(41)
    for kind, n in ((0, 128), (1, 100), (1, 128)):
        A = np.asfortranarray(rng.random((n, n, 2)) + (1j * rng.random((n, n, 2)) if kind else 0))
        Hs = gs.hessenberg(A)
        for b in range(2):
            F = Hs.factors[:, :, b]
            berr, oerr, _ = O.residuals(A[:, :, b], np.triu(F, -1), Hs.Q[:, :, b], kind)
            assert berr < 10 and oerr < 10, (kind, n, berr, oerr)
            Fo, tauo, _ = O.hessenberg(A[:, :, b], kind)
            np.testing.assert_allclose(np.triu(F, -1), np.triu(Fo, -1), rtol=0, atol=1e-10 * np.abs(A).max())
            np.testing.assert_allclose(np.tril(F, -2), np.tril(Fo, -2), rtol=0, atol=1e-10)
            np.testing.assert_allclose(Hs.tau[:, b], tauo, rtol=0, atol=1e-11)
    # one large Float64 matrix goes to the blocked reduction
    n = 300
    A = np.asfortranarray(rng.random((n, n)))
    Hs = gs.hessenberg(A)
    Hm = np.triu(Hs.factors, -1)
    assert np.linalg.norm(A - Hs.Q @ Hm @ Hs.Q.T) / (n * ULP * np.linalg.norm(A)) < 10
    assert np.linalg.norm(Hs.Q.T @ Hs.Q - np.eye(n)) / (n * ULP) < 10


def test_device_mode_counts_failures_without_info(gs):
    """GSCHUR_FLAG_DEVICE_PTRS with info == NULL: the return value still counts the matrices that hit the iteration cap."""
    import ctypes
    import torch
    from genericschur_jl_b200 import _lib
    L = _lib.lib()
    n, batch = 6, 5
    A = torch.rand((batch, n, n), dtype=torch.float64, device="cuda")
    A[1, 2, 3] = float("nan")
    A[3, 0, 0] = float("nan")
    Z = torch.empty_like(A)
    w = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
    rc = L.gschur_cuda_batched(0, n, batch, ctypes.c_void_p(A.data_ptr()), n, n * n, ctypes.c_void_p(Z.data_ptr()), n,
                               n * n, ctypes.c_void_p(w.data_ptr()), 1, 0, None, None, None, 0, _lib.FLAG_DEVICE_PTRS)
    assert rc == 2, rc


def test_pipeline_budget_and_release(gs, O, monkeypatch):
    """The host-pointer pipeline bounds its device buffers by a byte budget (more, smaller chunks: same result bit for bit)
    and gschur_cuda_release_workspace() frees what is cached; pageable and pinned callers get the same result."""
    rng = np.random.default_rng(43)
    n, batch = 24, 6000
    A = np.asfortranarray(rng.random((n, n, batch)))
    S0 = gs.gschur(A)
    monkeypatch.setenv("GSCHUR_PIPE_BUDGET_MB", "1")       # 1 MiB: ~56 matrices per chunk
    S1 = gs.gschur(A)
    monkeypatch.delenv("GSCHUR_PIPE_BUDGET_MB")
    assert np.array_equal(S0.T, S1.T) and np.array_equal(S0.Z, S1.Z) and np.array_equal(S0.values, S1.values)
    assert gs.release_workspace() == 0
    monkeypatch.setenv("GSCHUR_HOST_REGISTER", "1")
    S2 = gs.gschur(A)
    assert np.array_equal(S0.T, S2.T) and np.array_equal(S0.Z, S2.Z)
    for b in (0, batch - 1):
        _check_one(O, A[:, :, b], S1.T[:, :, b], S1.Z[:, :, b], S1.values[:, b], 0, 10, f"budget[{b}]")


def test_pageable_staging_matches_direct_path(gs, O, monkeypatch):
    """Ordinary (pageable) arrays go through the library's pinned staging and copy threads: same bytes out as the direct
    path — Schur with Z, eigenvalues only, Hessenberg factors, and with more chunks than staging buffers."""
    rng = np.random.default_rng(44)
    n, batch = 32, 5000
    A = np.asfortranarray(rng.random((n, n, batch)) + 1j * rng.random((n, n, batch)))
    monkeypatch.setenv("GSCHUR_HOST_STAGING", "0")
    S0 = gs.gschur(A)
    w0 = gs.eigvals(A)
    H0 = gs.hessenberg(A)
    monkeypatch.delenv("GSCHUR_HOST_STAGING")
    for budget in (None, "2"):                              # 2 MiB of staging: ~15 matrices per chunk, hundreds of chunks
        if budget:
            monkeypatch.setenv("GSCHUR_STAGE_BUDGET_MB", budget)
        S1 = gs.gschur(A)
        assert np.array_equal(S0.T, S1.T) and np.array_equal(S0.Z, S1.Z) and np.array_equal(S0.values, S1.values)
        assert np.array_equal(w0, gs.eigvals(A))
        H1 = gs.hessenberg(A)
        assert np.array_equal(H0.factors, H1.factors) and np.array_equal(H0.tau, H1.tau) and np.array_equal(H0.Q, H1.Q)
    for b in (0, batch - 1):
        _check_one(O, A[:, :, b], S1.T[:, :, b], S1.Z[:, :, b], S1.values[:, b], gs.C64, 10, f"staged[{b}]")
