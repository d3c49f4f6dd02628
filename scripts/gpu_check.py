"""Ad-hoc GPU correctness/timing probe (development aid; the real tests live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package, load_oracle
gs = load_package(); O = load_oracle()
print("devices", gs.device_count(), "max n", [gs.max_batched_n(k) for k in range(4)])
rng = np.random.default_rng(7)

def check(kind, n, batch, nchk=3):
    if kind == gs.F64:
        A = np.asfortranarray(rng.random((n, n, batch)))
    elif kind == gs.C64:
        A = np.asfortranarray(rng.random((n, n, batch)) + 1j * rng.random((n, n, batch)))
    elif kind == gs.DD:
        A = np.zeros((2, n, n, batch), order="F"); A[0] = rng.random((n, n, batch)); A = A.view(gs.DDArray)
    else:
        A = np.zeros((4, n, n, batch), order="F"); A[0] = rng.random((n, n, batch)); A[2] = rng.random((n, n, batch)); A = A.view(gs.CDDArray)
    t = time.time()
    try:
        S = gs.gschur(A, check=False)
    except Exception as e:
        print("FAIL", kind, n, batch, repr(e)); return
    dt = time.time() - t
    bad = int(np.count_nonzero(S.info))
    worst = (0, 0, 0)
    for b in np.linspace(0, batch - 1, nchk).astype(int):
        be, oe, _ = O.residuals(np.asarray(A[..., b]), np.asarray(S.T[..., b]), np.asarray(S.Z[..., b]), kind)
        Tb = np.asarray(S.T[..., b])
        if kind in (gs.F64, gs.C64):
            _, _, wr, rc, st = O.gschur(np.asarray(A[..., b]), kind)
            ed = np.abs(np.sort_complex(S.values[:, b]) - np.sort_complex(wr)).max()
            low = np.abs(np.tril(Tb, -1 if kind == gs.C64 else -2)).max()
        else:
            ed = 0; low = 0; st = None
        worst = (max(worst[0], be), max(worst[1], oe), max(worst[2], ed))
    print(f"kind={kind} n={n} batch={batch} time={dt*1e3:.1f}ms unconverged={bad} backward={worst[0]:.3f} orth={worst[1]:.3f} eigdiff={worst[2]:.2e} lower={low} gpu_stats={S.stats[:, 0]} oracle_stats={st}")

for kind, n, batch in [(gs.F64, 4, 8), (gs.F64, 32, 256), (gs.C64, 4, 8), (gs.C64, 32, 64), (gs.C64, 64, 64), (gs.F64, 64, 64),
                       (gs.F64, 7, 3), (gs.C64, 33, 5), (gs.F64, 100, 4), (gs.DD, 16, 8), (gs.CDD, 16, 8), (gs.CDD, 40, 4)]:
    check(kind, n, batch)

# timing, device resident
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
    print(f"bench kind={kind} n={n} batch={batch}: {best:.2f} ms -> {batch/best*1e3:.0f} matrices/s, unconverged={int((info!=0).sum())}")
bench(gs.F64, 32, 16384)
bench(gs.C64, 64, 4096)
bench(gs.F64, 64, 8192)
