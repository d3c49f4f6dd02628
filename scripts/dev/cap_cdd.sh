cat > /tmp/dd_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from __graft_entry__ import load_package
gs = load_package()
n, batch = 96, 148
hi = torch.rand((batch, n, n, 2), dtype=torch.float64, device="cuda")
A = torch.stack([hi[..., 0], torch.zeros_like(hi[..., 0]), hi[..., 1], torch.zeros_like(hi[..., 0])], dim=-1).contiguous()
Z = torch.empty_like(A); w = torch.empty((batch, n, 4), dtype=torch.float64, device="cuda"); info = torch.zeros(batch, dtype=torch.int32, device="cuda")
gs.gschur_device_(gs.CDD, n, batch, A.data_ptr(), Z.data_ptr(), w.data_ptr(), info.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print("unconverged", int((info != 0).sum()))
PY
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:gschur_qr_kernel -c 1 -o gpurun_out/r02n_cdd96 python /tmp/dd_one.py > gpurun_out/ncu_r02n_cdd.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02n_cdd96.ncu-rep | grep -E "time_duration|stalled|issue_active|pipe_fp64_cycles|registers"
