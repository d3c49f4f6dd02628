import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from __graft_entry__ import load_package
gs = load_package()
from genericschur_jl_b200 import _lib
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A0 = torch.rand((n, n), dtype=torch.float64, device='cuda'); A = torch.empty_like(A0); Q = torch.empty_like(A0)
for withq in (1, 0):
    for r in range(3):
        A.copy_(A0); torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = L.gschur_cuda_hessenberg_large(n, ctypes.c_void_p(A.data_ptr()), n, None, ctypes.c_void_p(Q.data_ptr()) if withq else None, n, 1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"n={n} withQ={withq}: rc={rc} {1e3*dt:.1f} ms", flush=True)
