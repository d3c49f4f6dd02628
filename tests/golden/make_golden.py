"""Generates tests/golden/cases.npz: the reference's test-matrix classes (inputs) together with what the CPU
oracle and the LAPACK Fortran originals (xGEHD2 + xLAHQR, reached through scipy's OpenBLAS) return for them.

The reference (Julia) cannot be executed in this environment, so these fixtures are NOT outputs of the Julia code:
they pin (a) the oracle against regressions bit-for-bit, (b) the oracle against the LAPACK routines the reference
was translated from, (c) the GPU path against both, on inputs that do not depend on the GPU box having scipy's
LAPACK test-matrix generators.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import ROOT, csort, godunov, reference_classes  # noqa: E402
import tmg  # noqa: E402

sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    out = {}
    names = []
    for complex_ in (False, True):
        kind = 1 if complex_ else 0
        for name, A, tol in reference_classes(complex_):
            key = ("c_" if complex_ else "r_") + name
            T, Z, w, rc, st = O.gschur(A, kind)
            assert rc == 0, key
            berr, oerr, anorm = O.residuals(A, T, Z, kind)
            # LAPACK originals on the same input: gehd2 -> lahqr (no balancing, no scaling: compare only when the
            # oracle itself did not need _scale!, i.e. always here because lahqr is scale-free for these magnitudes)
            # xLAHQR has no scaling of its own (xGEES scales before calling it): feed it A / max|a_ij| when the
            # magnitude is extreme and scale the eigenvalues back.
            amax = float(np.max(np.abs(A)))
            sc = amax if (amax > 0 and not 1e-100 < amax < 1e100) else 1.0
            F, tau = tmg.lapack_gehd2(A / sc)
            try:
                _, _, wl = tmg.lapack_lahqr(np.triu(F, -1), wantz=False)
                wl = wl * sc
            except AssertionError:
                wl = np.full(A.shape[0], np.nan + 0j)
            out[key + "__A"] = A
            out[key + "__w"] = w
            out[key + "__wlapack"] = wl
            out[key + "__meta"] = np.array([kind, tol, berr, oerr, st[0], st[1], st[3]], dtype=np.float64)
            names.append(key)
    # cfg1 of BASELINE.json: one 64x64 random ComplexF64 matrix
    rng = np.random.default_rng(1234 + 1)
    A = np.asfortranarray(rng.random((64, 64)) + 1j * rng.random((64, 64)))
    T, Z, w, rc, st = O.gschur(A, 1)
    berr, oerr, _ = O.residuals(A, T, Z, 1)
    F, tau = tmg.lapack_gehd2(A)
    _, _, wl = tmg.lapack_lahqr(np.triu(F, -1), wantz=False)
    out["c_cfg1_n64__A"], out["c_cfg1_n64__w"], out["c_cfg1_n64__wlapack"] = A, w, wl
    out["c_cfg1_n64__meta"] = np.array([1, 10, berr, oerr, st[0], st[1], st[3]], dtype=np.float64)
    names.append("c_cfg1_n64")
    # Godunov in MPFR-256: the only known-answer eigenvalue fixture of the reference (test/testfuncs.jl:127-142)
    G, vals, econd = godunov()
    Tm, Zm, wm, rc = O.gschur_mp(G, 0)
    assert rc == 0
    out["godunov__A"] = G
    out["godunov__w_mp_dd"] = wm
    out["godunov__vals"] = vals
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "cases.npz"), **out)
    print("wrote", len(names), "cases;", os.path.getsize(os.path.join(HERE, "cases.npz")), "bytes")


if __name__ == "__main__":
    main()
