#!/bin/bash
# One gpurun call: ncu --set full captures of the batched kernels + the bench launch list (results under gpurun_out/).
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:gschur_qr_kernel --launch-skip 1 -c 1 -o gpurun_out/r01b_stageB_cfg3 python scripts/prof_one.py 1 64 2960 > gpurun_out/ncu_b3.log 2>&1
$NCU -k regex:gehrd_q_kernel --launch-skip 1 -c 1 -o gpurun_out/r01b_stageA_cfg3 python scripts/prof_one.py 1 64 2960 > gpurun_out/ncu_a3.log 2>&1
$NCU -k regex:gschur_qr_kernel --launch-skip 1 -c 1 -o gpurun_out/r01b_stageB_cfg2 python scripts/prof_one.py 0 32 16384 > gpurun_out/ncu_b2.log 2>&1
$NCU -k regex:gschur_qr_kernel --launch-skip 1 -c 1 -o gpurun_out/r01b_stageB_f64n64 python scripts/prof_one.py 0 64 4096 > gpurun_out/ncu_b64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 1 --no-others --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/ncu_b3.log gpurun_out/bench_under_ncu.log
