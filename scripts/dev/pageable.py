"""e2e on pageable host arrays: the library's pinned staging against the driver's own staging (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from __graft_entry__ import load_package
gs = load_package()
n, batch = 64, 65536
rng = np.random.default_rng(0)
A0 = np.asfortranarray(rng.random((n, n, batch)) + 1j * rng.random((n, n, batch)))
Z = np.zeros_like(A0)
ref = None
for stg in ("1", "1", "1", "0", "0"):
    os.environ["GSCHUR_HOST_STAGING"] = stg
    A = A0.copy(order="F")
    t0 = time.perf_counter()
    S = gs.gschur_(A, Z=Z)
    dt = time.perf_counter() - t0
    sig = (float(np.abs(A).sum()), float(np.abs(Z).sum()))
    if ref is None:
        ref = sig
    print(f"GSCHUR_HOST_STAGING={stg}: {1e3*dt:.0f} ms -> {batch/dt:.0f} matrices/s  same={sig == ref}", flush=True)
