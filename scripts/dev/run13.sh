ncu --set full --clock-control none --import-source on -f -k regex:gehrd_reg -c 1 -o gpurun_out/r02h_stageA_reg ./scripts/dev/qr3_dev_real 64 4736 1 > gpurun_out/ncu_h_a.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02h_stageA_reg.ncu-rep
