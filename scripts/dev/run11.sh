# round-2 profile captures (one GPU)
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:gschur_chain_kernel -c 1 -o gpurun_out/r02g_stageB_c64 ./scripts/dev/qr3_dev 64 8880 1 > gpurun_out/ncu_g_b_c64.log 2>&1
$NCU -k regex:gschur_chain_kernel -c 1 -o gpurun_out/r02g_stageB_f64 ./scripts/dev/qr3_dev_real 64 8880 1 > gpurun_out/ncu_g_b_f64.log 2>&1
$NCU -k regex:gschur_zreg_kernel -c 1 -o gpurun_out/r02g_stageC_c64 ./scripts/dev/qr3_dev 64 8880 1 > gpurun_out/ncu_g_c_c64.log 2>&1
$NCU -k regex:gschur_zreg_kernel -c 1 -o gpurun_out/r02g_stageC_f64 ./scripts/dev/qr3_dev_real 64 8880 1 > gpurun_out/ncu_g_c_f64.log 2>&1
$NCU -k regex:gehrd -c 1 -o gpurun_out/r02g_stageA_c64 ./scripts/dev/qr3_dev 64 8880 1 > gpurun_out/ncu_g_a_c64.log 2>&1
$NCU -k regex:gehrd -c 1 -o gpurun_out/r02g_stageA_f64 ./scripts/dev/qr3_dev_real 64 8880 1 > gpurun_out/ncu_g_a_f64.log 2>&1
# the double-double kernel (config 5 shape, 148 matrices)
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
$NCU -k regex:gschur_qr_kernel -c 1 -o gpurun_out/r02g_stageB_cdd96 python /tmp/dd_one.py > gpurun_out/ncu_g_cdd.log 2>&1
tail -2 gpurun_out/ncu_g_cdd.log
# summaries on the box (the captures together exceed what travels back), then drop the big files
for f in r02g_stageB_c64 r02g_stageB_f64 r02g_stageC_c64 r02g_stageC_f64 r02g_stageA_c64 r02g_stageA_f64 r02g_stageB_cdd96; do
  python scripts/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.summary.txt 2>&1
done
mkdir -p profiles; cp profiles/kernel_traffic.json gpurun_out/kernel_traffic.json 2>/dev/null
python scripts/ncu_traffic.py stageB_c64_n64 gpurun_out/r02g_stageB_c64.ncu-rep 8880 > /dev/null
python scripts/ncu_traffic.py stageB_f64_n64 gpurun_out/r02g_stageB_f64.ncu-rep 8880 > /dev/null
cp profiles/kernel_traffic.json gpurun_out/kernel_traffic.json
rm -f gpurun_out/r02g_stageC_*.ncu-rep gpurun_out/r02g_stageA_*.ncu-rep gpurun_out/r02g_stageB_f64.ncu-rep gpurun_out/r02g_stageB_cdd96.ncu-rep
# launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_under_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r02g_launches_bench_cfg3.csv | head -12
