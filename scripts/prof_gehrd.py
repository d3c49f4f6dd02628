import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
gs = load_package()
from genericschur_jl_b200 import _lib
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A = torch.rand((n, n), dtype=torch.float64, device='cuda'); Q = torch.empty_like(A)
rc = L.gschur_cuda_hessenberg_large(n, ctypes.c_void_p(A.data_ptr()), n, None, ctypes.c_void_p(Q.data_ptr()), n, 1)
torch.cuda.synchronize(); print(rc)
