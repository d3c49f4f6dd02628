"""Small runs for compute-sanitizer (racecheck / memcheck): one launch of each batched path (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from __graft_entry__ import load_package
gs = load_package()
rng = np.random.default_rng(5)
which = sys.argv[1] if len(sys.argv) > 1 else "cdd"
if which == "cdd":
    n, batch = 96, 2
    A = np.zeros((4, n, n, batch), order="F"); A[0] = rng.random((n, n, batch)); A[2] = rng.random((n, n, batch))
    S = gs.gschur(np.asfortranarray(A).view(gs.CDDArray))
elif which == "c64":
    n, batch = 64, 8
    S = gs.gschur(np.asfortranarray(rng.random((n, n, batch)) + 1j * rng.random((n, n, batch))))
else:
    n, batch = 64, 8
    S = gs.gschur(np.asfortranarray(rng.random((n, n, batch))))
print(which, "ok", None if S is None else S.values.shape)
