python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02s_pytest_gpu.log; tail -3 gpurun_out/r02s_pytest_gpu.log
bash scripts/capture_profiles.sh r02s > gpurun_out/r02s_capture.log 2>&1; tail -30 gpurun_out/r02s_capture.log
python bench.py > gpurun_out/r02s_bench_1gpu.json 2> gpurun_out/r02s_bench_1gpu.err; tail -c 600 gpurun_out/r02s_bench_1gpu.json
python bench.py --impl reference > gpurun_out/r02s_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r02s_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; tail -2 gpurun_out/r02s_smoke.log
