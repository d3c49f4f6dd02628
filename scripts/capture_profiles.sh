#!/bin/bash
# Profile captures of one round on ONE B200 (run through gpurun from the repository root):
#   bash scripts/capture_profiles.sh r02k
# writes gpurun_out/<tag>_*.summary.txt (ncu --set full, one launch each of stages A / B / C for 64x64 ComplexF64 and
# Float64, and of the complex double-double QR kernel at 96x96), the launch lists of `bench.py` and of the large-matrix
# path (first 700 launches of the Hessenberg reduction, the QR kernels of three sweeps), and refreshes profiles/kernel_traffic.json.  The binary captures are summarised on the box and deleted (together
# they exceed what travels back).  Needs the development harness binaries (scripts/dev/qr3_dev.cu, see its header).
set -u
TAG=${1:-r02}
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {   # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  $NCU -k regex:$rx -c 1 -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${TAG}_$name.ncu-rep > gpurun_out/${TAG}_$name.summary.txt 2>&1
}
cap stageB_c64 gschur_chain_kernel ./scripts/dev/qr3_dev 64 8880 1
cap stageB_f64 gschur_chain_kernel ./scripts/dev/qr3_dev_real 64 8880 1
python scripts/ncu_traffic.py stageB_c64_n64 gpurun_out/${TAG}_stageB_c64.ncu-rep 8880 > /dev/null
python scripts/ncu_traffic.py stageB_f64_n64 gpurun_out/${TAG}_stageB_f64.ncu-rep 8880 > /dev/null
cp profiles/kernel_traffic.json gpurun_out/kernel_traffic.json
cap stageC_c64 gschur_zreg_kernel ./scripts/dev/qr3_dev 64 8880 1
cap stageC_f64 gschur_zreg_kernel ./scripts/dev/qr3_dev_real 64 8880 1
cap stageA_c64 gehrd ./scripts/dev/qr3_dev 64 8880 1
cap stageA_f64 gehrd ./scripts/dev/qr3_dev_real 64 8880 1
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
cap stageB_cdd96 gschur_qr_kernel python /tmp/dd_one.py
rm -f gpurun_out/${TAG}_*.ncu-rep
# launch lists: the bench command (headline workload only) and the large-matrix path (Hessenberg + 3 QR sweeps at n = 4096)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_cfg3.csv \
    python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_under_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_bench_cfg3.csv > gpurun_out/${TAG}_launches_bench_cfg3.shares.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_gehrd4096_first700.csv \
    python scripts/prof_gehrd.py 4096 > /dev/null 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_gehrd4096_first700.csv > gpurun_out/${TAG}_launches_gehrd4096_first700.shares.txt
GSCHUR_LQ_MAXSWEEPS=3 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lq_ -c 1200 --csv \
    --log-file gpurun_out/${TAG}_launches_lqr4096_3sweeps.csv python scripts/dev/large_time.py > /dev/null 2>&1
python scripts/launch_shares.py gpurun_out/${TAG}_launches_lqr4096_3sweeps.csv > gpurun_out/${TAG}_launches_lqr4096_3sweeps.shares.txt
head -14 gpurun_out/${TAG}_launches_bench_cfg3.shares.txt gpurun_out/${TAG}_launches_gehrd4096_first700.shares.txt gpurun_out/${TAG}_launches_lqr4096_3sweeps.shares.txt
