GSCHUR_CHAIN=1 ./scripts/dev/qr3_dev_real_prof 64 16384 1 | tail -2
GSCHUR_CHAIN=32 ./scripts/dev/qr3_dev_real_prof 64 16384 1 | tail -2
NCU="ncu --set full --clock-control none --import-source on -f"
GSCHUR_CHAIN=1 $NCU -k regex:gschur_chain_kernel -c 1 -o gpurun_out/r02c_own_real ./scripts/dev/qr3_dev_real 64 8880 1 > gpurun_out/ncu_own_real.log 2>&1
