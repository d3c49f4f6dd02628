"""Occupancy / Z sweep of the stage-B kernel (development aid): run once per value of GSCHUR_QR_CTAS_PER_SM.
usage: python scripts/probe_occ.py kind n batch"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
gs = load_package()
from importlib import import_module
_L = import_module(gs.__name__ + "._lib").lib()
kind = int(sys.argv[1]); n = int(sys.argv[2]); batch = int(sys.argv[3])
dt = torch.float64 if kind == 0 else torch.complex128
torch.manual_seed(3)
A0 = torch.rand((batch, n, n), dtype=dt, device="cuda")
Z = torch.empty_like(A0); w = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
info = torch.zeros(batch, dtype=torch.int32, device="cuda")
stats = torch.zeros((batch, 4), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for wantZ in (True, False):
    best = (1e9, 0, 0)
    for r in range(3):
        A = A0.clone()
        _L.gschur_cuda_stage_timing(1, None, None)
        gs.gschur_device_(kind, n, batch, A.data_ptr(), Z.data_ptr() if wantZ else 0, w.data_ptr(), info.data_ptr(), stats.data_ptr(), stream=st)
        a, b = ctypes.c_float(0), ctypes.c_float(0)
        _L.gschur_cuda_stage_timing(0, ctypes.byref(a), ctypes.byref(b))
        torch.cuda.synchronize()
        if b.value < best[0]: best = (b.value, a.value)
    steps = float(stats[:, 1].double().mean())
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    cyc = best[0] * 1e-3 * 1.965e9 * sms / (batch * steps)
    print(f"ctas/sm={os.environ.get('GSCHUR_QR_CTAS_PER_SM','max')} kind={kind} n={n} batch={batch} wantZ={wantZ}: stageA {best[1]:.2f} ms, stageB {best[0]:.2f} ms -> {batch/best[0]*1e3:.0f} mat/s (B only), steps/matrix {steps:.0f}, SM-cycles per step {cyc:.0f}, unconverged={int((info!=0).sum())}", flush=True)
    if os.environ.get("GS_PROF"):
        s = stats.double().mean(0)
        print(f"profile (cycles/64 per matrix): total {s[0]:.0f}  wait-on-Z {s[2]:.0f}  step loops {s[3]:.0f}  steps {s[1]:.0f} -> loop cycles/step {64*s[3]/s[1]:.0f}, total cycles/step {64*s[0]/s[1]:.0f}")
