import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from __graft_entry__ import load_package
gs = load_package()
for kind, n, batch in ((gs.CDD, 96, 592), (gs.DD, 96, 592), (gs.CDD, 64, 592), (gs.CDD, 32, 2048)):
    lead = 4 if kind == gs.CDD else 2
    hi = torch.rand((batch, n, n, lead), dtype=torch.float64, device="cuda")
    if lead == 4:
        hi[..., 1] = 0; hi[..., 3] = 0
    else:
        hi[..., 1] = 0
    A0 = hi.contiguous(); A = torch.empty_like(A0); Z = torch.empty_like(A0)
    w = torch.empty((batch, n, 4), dtype=torch.float64, device="cuda"); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    for r in range(2):
        A.copy_(A0); torch.cuda.synchronize(); t0 = time.perf_counter()
        gs.gschur_device_(kind, n, batch, A.data_ptr(), Z.data_ptr(), w.data_ptr(), info.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"kind={kind} n={n} batch={batch}: {1e3*dt:.1f} ms -> {batch/dt:.0f} matrices/s, unconverged={int((info!=0).sum())}", flush=True)
