"""Three-stage path (stage A -> stage B with reflector log -> stage C replay) against the fused path, plus stage timings.
Development aid.  python scripts/gpu_qr3.py [check] [bench]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from __graft_entry__ import load_oracle, load_package

gs = load_package()
from importlib import import_module

L = import_module("genericschur_jl_b200._lib").lib()


def bench(kind, n, batch, reps=3, wantZ=True):
    dt = torch.float64 if kind == gs.F64 else torch.complex128
    torch.manual_seed(1)
    A0 = torch.rand((batch, n, n), dtype=dt, device="cuda")
    Z = torch.empty_like(A0)
    w = torch.empty((batch, n), dtype=torch.complex128, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    best = (1e9, 0, 0, 0)
    L.gschur_cuda_stage_timing3(1, None, None, None)
    for r in range(reps):
        A = A0.clone()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        gs.gschur_device_(kind, n, batch, A.data_ptr(), Z.data_ptr() if wantZ else 0, w.data_ptr(), info.data_ptr(), stream=st)
        e1.record()
        torch.cuda.synchronize()
        a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        L.gschur_cuda_stage_timing3(-1, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        t = e0.elapsed_time(e1)
        if t < best[0]:
            best = (t, a.value, b.value, c.value)
    L.gschur_cuda_stage_timing3(0, None, None, None)
    print(f"[{os.environ.get('GSCHUR_QR', 'log')}] kind={kind} n={n} batch={batch} wantZ={wantZ}: {best[0]:.2f} ms "
          f"(A {best[1]:.2f} + B {best[2]:.2f} + C {best[3]:.2f}) -> {batch / best[0] * 1e3:.0f} matrices/s, "
          f"unconverged={int((info != 0).sum())}", flush=True)


def check():
    O = load_oracle()
    rng = np.random.default_rng(7)
    for kind, n, batch in [(0, 3, 4), (0, 5, 300), (0, 32, 16), (0, 33, 8), (0, 64, 40), (1, 2, 4), (1, 7, 300), (1, 32, 8),
                           (1, 47, 8), (1, 64, 40)]:
        A = np.asfortranarray(rng.random((n, n, batch)) + (1j * rng.random((n, n, batch)) if kind else 0))
        os.environ.pop("GSCHUR_QR", None)
        S = gs.gschur(A, check=False)
        os.environ["GSCHUR_QR"] = "fused"
        S0 = gs.gschur(A, check=False)
        os.environ.pop("GSCHUR_QR", None)
        same_T = np.array_equal(S.T, S0.T) and np.array_equal(S.values, S0.values) and np.array_equal(S.stats, S0.stats)
        dz = float(np.abs(S.Z - S0.Z).max())
        worst = [0, 0]
        for b in range(min(batch, 3)):
            be, oe, _ = O.residuals(A[..., b], S.T[..., b], S.Z[..., b], kind)
            worst = [max(worst[0], be), max(worst[1], oe)]
        print(f"check kind={kind} n={n} batch={batch}: unconverged={int(np.count_nonzero(S.info))} T/w/stats identical to fused: "
              f"{same_T}; max|Z - Z_fused| = {dz:.2e}; backward={worst[0]:.3f} orth={worst[1]:.3f}", flush=True)
    # tiny pool: force log overflow -> redo by the fused kernel
    os.environ["GSCHUR_LOG_TEST_TINY"] = "1"
    A = np.asfortranarray(rng.random((24, 24, 50)) + 1j * rng.random((24, 24, 50)))
    S = gs.gschur(A, check=False)
    os.environ.pop("GSCHUR_LOG_TEST_TINY", None)
    S0 = gs.gschur(A, check=False)
    print("overflow/redo: identical T:", np.array_equal(S.T, S0.T), " max|dZ| =", float(np.abs(S.Z - S0.Z).max()),
          " unconverged =", int(np.count_nonzero(S.info)), flush=True)


if __name__ == "__main__":
    if "check" in sys.argv:
        check()
    if "bench" in sys.argv:
        for sel in ("log", "fused"):
            if sel == "fused":
                os.environ["GSCHUR_QR"] = "fused"
            else:
                os.environ.pop("GSCHUR_QR", None)
            bench(gs.C64, 64, 16384)
            bench(gs.F64, 64, 16384)
            bench(gs.F64, 32, 16384)
            bench(gs.C64, 32, 16384)
            bench(gs.C64, 64, 16384, wantZ=False)
        os.environ.pop("GSCHUR_QR", None)
