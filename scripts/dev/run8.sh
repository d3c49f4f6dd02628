python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -x -q -m gpu -k "hessenberg or device_mode or pipeline_budget or empty or strided" > gpurun_out/r02e_api_tests.log 2>&1; tail -15 gpurun_out/r02e_api_tests.log
python bench.py --steps 3 --warmup 3 --e2e-steps 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -5 gpurun_out/r02e_bench.err; head -c 1500 gpurun_out/r02e_bench.json
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:gschur_chain_kernel -c 1 -o gpurun_out/r02e_stageB_c64 ./scripts/dev/qr3_dev 64 8880 1 > gpurun_out/ncu_stageB_c64.log 2>&1
