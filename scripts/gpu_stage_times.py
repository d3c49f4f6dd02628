"""Device-resident timings with the stage A / stage B split, a few shapes (development aid).
python scripts/gpu_stage_times.py [check]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package, load_oracle
gs = load_package()
from importlib import import_module
L = import_module("genericschur_jl_b200._lib").lib()
def bench(kind, n, batch, reps=3):
    dt = torch.float64 if kind == gs.F64 else torch.complex128
    torch.manual_seed(1)
    A0 = torch.rand((batch, n, n), dtype=dt, device="cuda")
    Z = torch.empty_like(A0); w = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    best = (1e9, 0, 0)
    L.gschur_cuda_stage_timing(1, None, None)
    for r in range(reps):
        A = A0.clone()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); gs.gschur_device_(kind, n, batch, A.data_ptr(), Z.data_ptr(), w.data_ptr(), info.data_ptr(), stream=st); e1.record()
        torch.cuda.synchronize()
        a = ctypes.c_float(); b = ctypes.c_float()
        L.gschur_cuda_stage_timing(-1, ctypes.byref(a), ctypes.byref(b))
        t = e0.elapsed_time(e1)
        if t < best[0]: best = (t, a.value, b.value)
    L.gschur_cuda_stage_timing(0, None, None)
    print(f"kind={kind} n={n} batch={batch}: {best[0]:.2f} ms (A {best[1]:.2f} + B {best[2]:.2f}) -> {batch/best[0]*1e3:.0f} matrices/s, unconverged={int((info!=0).sum())}", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "check":
    O = load_oracle()
    rng = np.random.default_rng(7)
    for kind, n, batch in [(0, 3, 4), (0, 32, 16), (0, 33, 8), (0, 64, 8), (1, 2, 4), (1, 32, 8), (1, 47, 8), (1, 64, 8)]:
        A = np.asfortranarray(rng.random((n, n, batch)) + (1j * rng.random((n, n, batch)) if kind else 0))
        S = gs.gschur(A, check=False)
        worst = [0, 0]
        for b in range(min(batch, 3)):
            be, oe, _ = O.residuals(A[..., b], S.T[..., b], S.Z[..., b], kind)
            worst = [max(worst[0], be), max(worst[1], oe)]
        print(f"check kind={kind} n={n}: unconverged={int(np.count_nonzero(S.info))} backward={worst[0]:.3f} orth={worst[1]:.3f}", flush=True)
bench(gs.F64, 32, 16384)
bench(gs.C64, 64, 16384)
bench(gs.F64, 64, 16384)
bench(gs.C64, 32, 16384)
