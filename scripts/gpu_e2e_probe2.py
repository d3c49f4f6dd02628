"""PCIe bandwidth from pinned memory and repeated end-to-end calls through the host API (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
gs = load_package()
n, batch = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
Ah = torch.empty((batch, n, n), dtype=torch.complex128).pin_memory()
Zh = torch.empty((batch, n, n), dtype=torch.complex128).pin_memory()
g = torch.Generator().manual_seed(3)
A0 = torch.rand((batch, n, n, 2), dtype=torch.float64, generator=g)
A0c = torch.view_as_complex(A0)
d = torch.empty((batch, n, n), dtype=torch.complex128, device="cuda")
for name, fn in (("H2D 4.3GB", lambda: d.copy_(Ah, non_blocking=True)), ("D2H 4.3GB", lambda: Zh.copy_(d, non_blocking=True))):
    for r in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name}: {dt*1e3:.1f} ms  {Ah.numel()*16/dt/1e9:.1f} GB/s", flush=True)
del d
torch.cuda.empty_cache()
Ah_np = Ah.numpy().T; Zh_np = Zh.numpy().T
for it in range(int(os.environ.get("E2E_CALLS", "5"))):
    Ah.copy_(A0c)
    t0 = time.perf_counter()
    S = gs.gschur_(Ah_np, Z=Zh_np, devices=[0])
    dt = time.perf_counter() - t0
    print(f"e2e call {it}: {dt*1e3:.1f} ms -> {batch/dt:.0f} matrices/s", flush=True)
