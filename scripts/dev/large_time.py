"""Large-matrix path: time split of one 4096 x 4096 Float64 Schur decomposition (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from __graft_entry__ import load_package
gs = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(7)
A = np.asfortranarray(rng.random((n, n)))
os.environ["GSCHUR_LQ_TIMING"] = "1"
for it in range(2):
    t0 = time.perf_counter()
    try:
        S = gs.gschur(A, check=False)
    except Exception as exc:
        print("exc", str(exc)[:80])
    dt = time.perf_counter() - t0
    print(f"n={n}: {1e3*dt:.0f} ms  launches so far {gs.launch_count()}  stats (sweeps, windows, small blocks) {getattr(S, 'stats', None)}", flush=True)
