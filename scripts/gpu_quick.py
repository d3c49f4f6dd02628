"""Quick GPU probe: correctness of the fast path on a few shapes + device-resident timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package, load_oracle
gs = load_package(); O = load_oracle()
rng = np.random.default_rng(7)
def check(kind, n, batch):
    A = np.asfortranarray(rng.random((n, n, batch)) + (1j * rng.random((n, n, batch)) if kind else 0))
    S = gs.gschur(A, check=False)
    worst = [0, 0, 0]
    for b in range(min(batch, 4)):
        be, oe, _ = O.residuals(A[..., b], S.T[..., b], S.Z[..., b], kind)
        _, _, wr, rc, st = O.gschur(A[..., b], kind)
        ed = np.abs(np.sort_complex(S.values[:, b]) - np.sort_complex(wr)).max()
        worst = [max(worst[0], be), max(worst[1], oe), max(worst[2], ed)]
    print(f"kind={kind} n={n} batch={batch} unconverged={int(np.count_nonzero(S.info))} backward={worst[0]:.3f} orth={worst[1]:.3f} eigdiff={worst[2]:.2e} stats={S.stats[:,0]} oracle={st}", flush=True)
for kind, n, batch in [(0, 3, 4), (0, 8, 8), (0, 32, 64), (0, 33, 8), (0, 64, 32), (1, 2, 4), (1, 8, 8), (1, 32, 32), (1, 47, 8), (1, 64, 32)]:
    check(kind, n, batch)
import torch
def bench(kind, n, batch, reps=3):
    dt = torch.float64 if kind == gs.F64 else torch.complex128
    A0 = torch.rand((batch, n, n), dtype=dt, device="cuda")
    Z = torch.empty_like(A0); w = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for r in range(reps):
        A = A0.clone()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); gs.gschur_device_(kind, n, batch, A.data_ptr(), Z.data_ptr(), w.data_ptr(), info.data_ptr(), stream=st); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print(f"bench kind={kind} n={n} batch={batch}: {best:.2f} ms -> {batch/best*1e3:.0f} matrices/s, unconverged={int((info!=0).sum())}", flush=True)
bench(gs.F64, 32, 16384)
bench(gs.C64, 64, 8192)
bench(gs.F64, 64, 16384)
