(timeout 1200 python -m pytest tests/test_gpu_parity_full.py -x -q -m gpu > gpurun_out/r02_parity_full.log 2>&1; echo rc=$? >> gpurun_out/r02_parity_full.log)
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:gschur_chain_kernel -c 1 -o gpurun_out/r02a_chain32_real ./scripts/dev/qr3_dev_real 64 2960 1 > gpurun_out/ncu_chain32_real.log 2>&1
GSCHUR_CHAIN=16 $NCU -k regex:gschur_chain_kernel -c 1 -o gpurun_out/r02a_chain16_real ./scripts/dev/qr3_dev_real 64 2960 1 > gpurun_out/ncu_chain16_real.log 2>&1
$NCU -k regex:gehrd -c 1 -o gpurun_out/r02a_stageA_real ./scripts/dev/qr3_dev_real 64 2960 1 > gpurun_out/ncu_stageA_real.log 2>&1
tail -5 gpurun_out/r02_parity_full.log
