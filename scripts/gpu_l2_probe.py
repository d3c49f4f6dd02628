"""L2 streaming bandwidth of the current B200 (the ceiling of stage B's Z stream) beside the
FP64 FMA peak.  Run on a GPU box: python scripts/gpu_l2_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package

gs = load_package()
g, ms = gs.measure_l2_bandwidth()
t, ms2 = gs.measure_fp64_peak()
print(json.dumps({"l2_stream_gbs": g, "l2_ms": ms, "fp64_tflops": t, "fp64_ms": ms2}))
