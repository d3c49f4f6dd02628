import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
gs = load_package()
from genericschur_jl_b200 import _lib
L = _lib.lib()
rng = np.random.default_rng(0)
# --- dgemm
for (ta, tb, M, N, K) in [(0,0,100,70,50), (1,0,33,129,200), (0,1,257,64,32), (1,1,65,65,17), (0,0,4096,96,96), (1,0,32,4000,4000)]:
    A = torch.rand((K, M) if not ta else (M, K), dtype=torch.float64, device='cuda')   # torch row-major (K,M) == col-major (M,K)
    B = torch.rand((N, K) if not tb else (K, N), dtype=torch.float64, device='cuda')
    C = torch.rand((N, M), dtype=torch.float64, device='cuda'); C0 = C.clone()
    lda = A.shape[1]; ldb = B.shape[1]
    rc = L.gschur_cuda_dgemm(ta, tb, M, N, K, 1.5, ctypes.c_void_p(A.data_ptr()), lda, ctypes.c_void_p(B.data_ptr()), ldb, 0.5, ctypes.c_void_p(C.data_ptr()), M)
    Am = A.T if not ta else A      # col-major (M,K) matrix as torch (M,K): A stored (K,M) row-major -> .T
    opA = (A.T if not ta else A.T.T)
    # build mathematically: col-major X with shape (r,c) is torch tensor of shape (c,r) transposed
    Amat = A.T            # (M,K) if not ta else (K,M)
    Bmat = B.T            # (K,N) if not tb else (N,K)
    opA = Amat if not ta else Amat.T
    opB = Bmat if not tb else Bmat.T
    ref = 1.5 * opA @ opB + 0.5 * C0.T
    err = (C.T - ref).abs().max().item()
    print("dgemm", ta, tb, M, N, K, "rc", rc, "maxerr", err)
# timing big gemm
M=N=K=4096
A = torch.rand((K,M), dtype=torch.float64, device='cuda'); B = torch.rand((N,K), dtype=torch.float64, device='cuda'); C = torch.zeros((N,M), dtype=torch.float64, device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t=time.time(); L.gschur_cuda_dgemm(0,0,M,N,K,1.0,ctypes.c_void_p(A.data_ptr()),M,ctypes.c_void_p(B.data_ptr()),K,0.0,ctypes.c_void_p(C.data_ptr()),M); dt=time.time()-t
print("dgemm 4096^3: %.1f ms -> %.2f TFLOP/s" % (dt*1e3, 2*M*N*K/dt/1e12))
# --- hessenberg large
for n in (64, 200, 512, 1024, 4096):
    A0 = rng.random((n, n)); A = np.asfortranarray(A0.copy()); Q = np.zeros((n, n), order='F'); tau = np.zeros(n)
    t = time.time()
    rc = L.gschur_cuda_hessenberg_large(n, A.ctypes.data_as(ctypes.c_void_p), n, tau.ctypes.data_as(ctypes.c_void_p), Q.ctypes.data_as(ctypes.c_void_p), n, 0)
    dt = time.time() - t
    H = np.triu(A, -1)
    res = np.linalg.norm(A0 - Q @ H @ Q.T) / (n * np.linalg.norm(A0) * 2.2e-16)
    orth = np.linalg.norm(Q.T @ Q - np.eye(n)) / (n * 2.2e-16)
    print(f"gehrd n={n} rc={rc} {L.gschur_cuda_large_last_error().decode()} time={dt*1e3:.1f}ms backward={res:.3f} orth={orth:.3f}", flush=True)
# device-resident timing n=4096
n = 4096
A = torch.rand((n, n), dtype=torch.float64, device='cuda'); Q = torch.empty_like(A)
for _ in range(2):
    A1 = A.clone(); torch.cuda.synchronize(); t = time.time()
    rc = L.gschur_cuda_hessenberg_large(n, ctypes.c_void_p(A1.data_ptr()), n, None, ctypes.c_void_p(Q.data_ptr()), n, 1)
    torch.cuda.synchronize(); dt = time.time() - t
print("gehrd+Q 4096 device-resident: %.1f ms" % (dt*1e3))
