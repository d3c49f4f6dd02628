"""Run one device-resident batch (for ncu captures): python scripts/prof_one.py kind n batch"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
gs = load_package()
kind, n, batch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dt = torch.float64 if kind == 0 else torch.complex128
torch.manual_seed(0)
A0 = torch.rand((batch, n, n), dtype=dt, device="cuda")
Z = torch.empty_like(A0); w = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
info = torch.zeros(batch, dtype=torch.int32, device="cuda")
for r in range(2):
    A = A0.clone()
    gs.gschur_device_(kind, n, batch, A.data_ptr(), Z.data_ptr(), w.data_ptr(), info.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
print("unconverged", int((info != 0).sum()))
