import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
gs = load_package()
from genericschur_jl_b200 import _lib
L = _lib.lib()
M=N=K=4096
A = torch.rand((K,M), dtype=torch.float64, device='cuda'); B = torch.rand((N,K), dtype=torch.float64, device='cuda'); C = torch.zeros((N,M), dtype=torch.float64, device='cuda')
for _ in range(3):
    L.gschur_cuda_dgemm(0,0,M,N,K,1.0,ctypes.c_void_p(A.data_ptr()),M,ctypes.c_void_p(B.data_ptr()),K,0.0,ctypes.c_void_p(C.data_ptr()),M)
torch.cuda.synchronize()
