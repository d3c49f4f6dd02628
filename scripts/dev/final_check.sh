python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02r_pytest_gpu.log; tail -2 gpurun_out/r02r_pytest_gpu.log
python bench.py > gpurun_out/r02r_bench_1gpu.json 2> gpurun_out/r02r_bench_1gpu.err; tail -c 300 gpurun_out/r02r_bench_1gpu.json
python bench.py --impl reference > gpurun_out/r02r_bench_reference.json 2>/dev/null; tail -c 200 gpurun_out/r02r_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1; tail -2 gpurun_out/r02r_smoke.log
