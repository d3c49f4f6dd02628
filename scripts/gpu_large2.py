import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
gs = load_package()
from genericschur_jl_b200 import _lib
L = _lib.lib()
rng = np.random.default_rng(0)
vp = ctypes.c_void_p
def run(n, wantZ=True):
    A0 = rng.random((n, n)); A = np.asfortranarray(A0.copy()); Z = np.zeros((n, n), order='F'); w = np.zeros(n, dtype=np.complex128)
    info = ctypes.c_int(0); st = (ctypes.c_longlong * 3)()
    t = time.time()
    rc = L.gschur_cuda_large(n, A.ctypes.data_as(vp), n, Z.ctypes.data_as(vp) if wantZ else None, n, w.ctypes.data_as(vp), 1, ctypes.byref(info), st, 0)
    dt = time.time() - t
    msg = L.gschur_cuda_large_last_error().decode()
    if rc != 0:
        print(f"n={n} rc={rc} info={info.value} {msg} time={dt:.2f}s stats={list(st)}", flush=True); return
    T = A
    low = np.abs(np.tril(T, -2)).max()
    res = np.linalg.norm(A0 - Z @ T @ Z.T) / (n * np.linalg.norm(A0) * 2.2e-16) if wantZ else -1
    orth = np.linalg.norm(Z.T @ Z - np.eye(n)) / (n * 2.2e-16) if wantZ else -1
    ev = np.linalg.eigvals(A0)
    from scipy.optimize import linear_sum_assignment
    if n <= 1500:
        D = np.abs(w[None, :] - ev[:, None]); r, c = linear_sum_assignment(D); ed = D[r, c].max()
    else:
        ed = abs(np.sort(w.real).sum() - np.sort(ev.real).sum())
    # standard form check of 2x2 blocks
    sub = np.diag(T, -1); bad = 0
    for j in np.nonzero(sub)[0]:
        if not (T[j, j] == T[j+1, j+1] and T[j, j+1] * T[j+1, j] < 0): bad += 1
    print(f"n={n} rc={rc} time={dt:.2f}s stats={list(st)} lower={low} backward={res:.3f} orth={orth:.3f} eigdiff={ed:.2e} nonstd_blocks={bad} trace_err={abs(w.sum()-np.trace(A0)):.2e}", flush=True)
for n in (150, 200, 300, 512, 1024):
    run(n)
run(2048)
run(4096)
