NCU="ncu --set full --clock-control none --import-source on -f"
GSCHUR_QR_CTAS_PER_SM=1 GSCHUR_CHAIN=2 $NCU -k regex:gschur_own2_kernel -c 1 -o gpurun_out/r02c_own2_c64_lone ./scripts/dev/qr3_dev 64 592 1 > gpurun_out/ncu_own2_c64_lone.log 2>&1
tail -3 gpurun_out/ncu_own2_c64_lone.log
