import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
gs = load_package()
from genericschur_jl_b200 import _lib
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.manual_seed(0)
A = torch.rand((n, n), dtype=torch.float64, device='cuda'); Z = torch.empty_like(A); w = torch.empty((n,), dtype=torch.complex128, device='cuda')
info = ctypes.c_int(0); st = (ctypes.c_longlong * 3)()
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
A0 = A.clone()
for r in range(reps):
    A.copy_(A0)
    torch.cuda.synchronize(); t = time.time()
    rc = L.gschur_cuda_large(n, ctypes.c_void_p(A.data_ptr()), n, ctypes.c_void_p(Z.data_ptr()), n, ctypes.c_void_p(w.data_ptr()), 1, ctypes.byref(info), st, 1)
    torch.cuda.synchronize(); print(rc, info.value, list(st), time.time() - t, flush=True)
