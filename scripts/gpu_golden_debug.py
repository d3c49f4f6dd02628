import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from common import *
import importlib.util
spec = importlib.util.spec_from_file_location('tg', os.path.join(ROOT, 'tests/test_gpu_parity.py')); tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)
gs = load_package(); O = load_oracle()
g = np.load(os.path.join(ROOT, 'tests/golden/cases.npz'))
for key in g['names']:
    key = str(key); A = np.asfortranarray(g[key+'__A']); kind = int(g[key+'__meta'][0])
    S = gs.gschur(A)
    sc = float(np.max(np.abs(A))) or 1.0
    Ac = np.asfortranarray((A / sc).astype(np.complex128))
    Tc, _, wc, rc, _ = O.gschur(Ac, 1)
    s = O.eigvalscond(Tc, 1)
    d = match_eigs(S.values / sc, wc, None)
    bound = ULP * fnorm(Ac) / np.where(np.isfinite(s) & (s > 0), s, 1e-300)
    ratio = d / bound
    i = int(np.argmax(np.where(s >= 1e-3, ratio, 0)))
    j = int(np.argmax(np.where((s >= 1e-6) & (s < 1e-3), ratio, 0)))
    print(f"{key:28s} max ratio (s>=1e-3): {ratio[i]:10.2f} at s={s[i]:.2e};  (1e-6<=s<1e-3): {ratio[j]:10.2f} at s={s[j]:.2e}  min s={np.nanmin(s):.1e}")
print("---- replicate test_golden_fixtures")
for key in g['names']:
    key = str(key); A = np.asfortranarray(g[key+'__A']); kind = int(g[key+'__meta'][0])
    S = gs.gschur(A)
    sc, wc, etol = tg._eig_tol(O, A)
    for nm in ('__w', '__wlapack'):
        ref = g[key+nm]
        if np.any(np.isnan(ref)): continue
        d = match_eigs(S.values / sc, ref / sc, np.where(np.isfinite(etol), etol, 1e300))
        bad = d > 2*etol
        if bad.any():
            i = int(np.argmax(np.where(bad, d/etol, 0)))
            print("FAIL", key, nm, "d", d[i], "etol", etol[i], "ref", ref[i]/sc, "nfinite", int(np.isfinite(etol).sum()))
print("---- dd orth")
rng = np.random.default_rng(77)
for n in (24, 48):
    A = np.zeros((4,n,n,2), order='F'); A[0]=rng.random((n,n,2)); A[2]=rng.random((n,n,2))
    S = gs.gschur(A.view(gs.CDDArray)); print('cdd', n, O.residuals(np.asfortranarray(A[...,0]), np.asarray(S.T[...,0]), np.asarray(S.Z[...,0]), 3)[:2])
    A = np.zeros((2,n,n,2), order='F'); A[0]=rng.random((n,n,2))
    S = gs.gschur(A.view(gs.DDArray)); print('dd ', n, O.residuals(np.asfortranarray(A[...,0]), np.asarray(S.T[...,0]), np.asarray(S.Z[...,0]), 2)[:2])
