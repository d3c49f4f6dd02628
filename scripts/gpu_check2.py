import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package, load_oracle
gs = load_package(); O = load_oracle()
rng = np.random.default_rng(7)
def mk(kind, n, batch):
    if kind == gs.F64: return np.asfortranarray(rng.random((n, n, batch)))
    if kind == gs.C64: return np.asfortranarray(rng.random((n, n, batch)) + 1j * rng.random((n, n, batch)))
    lead = 2 if kind == gs.DD else 4
    A = np.zeros((lead, n, n, batch), order="F")
    for p in range(0, lead, 2):
        hi = rng.random((n, n, batch)); lo = (rng.random((n, n, batch)) - 0.5) * 2.0**-53 * hi
        s = hi + lo; A[p] = s; A[p+1] = lo - (s - hi)
    return A.view(gs.DDArray if kind == gs.DD else gs.CDDArray)
def check(kind, n, batch, nchk=2, wantZ=True):
    A = mk(kind, n, batch)
    t = time.time()
    try:
        S = gs.gschur(A, check=False, wantZ=wantZ)
    except Exception as e:
        print("FAIL", kind, n, batch, repr(e), flush=True); return
    dt = time.time() - t
    bad = int(np.count_nonzero(S.info)); worst = (0, 0)
    if wantZ:
        for b in np.linspace(0, batch - 1, nchk).astype(int):
            be, oe, _ = O.residuals(np.asarray(A[..., b]), np.asarray(S.T[..., b]), np.asarray(S.Z[..., b]), kind)
            worst = (max(worst[0], be), max(worst[1], oe))
    print(f"kind={kind} n={n} batch={batch} wantZ={wantZ} time={dt*1e3:.1f}ms unconverged={bad} backward={worst[0]:.3f} orth={worst[1]:.3f} stats={S.stats[:, 0]}", flush=True)
for kind, n, batch in [(0, 100, 4), (0, 128, 4), (1, 96, 4), (1, 128, 3), (2, 16, 8), (2, 80, 4), (3, 40, 4), (3, 96, 4)]:
    check(kind, n, batch)
check(1, 128, 2, wantZ=False); check(3, 96, 2, wantZ=False)
t=time.time(); check(3, 96, 592, nchk=1); 
