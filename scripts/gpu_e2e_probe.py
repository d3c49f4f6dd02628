import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
gs = load_package()
n, batch = 64, 32768
x = torch.empty((batch, n, n), dtype=torch.complex128).pin_memory(); x.real.uniform_(); 
d = torch.empty_like(x, device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t=time.time(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); h2d=time.time()-t
    torch.cuda.synchronize(); t=time.time(); x.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h=time.time()-t
gb = x.numel()*16/1e9
print(f"pinned H2D {gb/h2d:.1f} GB/s, D2H {gb/d2h:.1f} GB/s")
A0 = torch.rand((batch, n, n), dtype=torch.complex128)
Ah = torch.empty((batch, n, n), dtype=torch.complex128).pin_memory(); Zh = torch.empty_like(Ah).pin_memory()
for wantZ in (True, False):
    for rep in range(3):
        Ah.copy_(A0); torch.cuda.synchronize(); t=time.time()
        S = gs.gschur_(Ah.numpy().T, Z=Zh.numpy().T if wantZ else None, wantZ=wantZ, devices=[0])
        dt=time.time()-t
    print(f"e2e wantZ={wantZ}: {dt*1e3:.0f} ms -> {batch/dt:.0f} matrices/s")
