python -m pytest tests -x -q -m gpu -k "large or cfg4" > gpurun_out/r02f_large_tests.log 2>&1; tail -5 gpurun_out/r02f_large_tests.log
python scripts/prof_gehrd.py 4096 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02f_launches_gehrd4096_first700.csv python scripts/prof_gehrd.py 4096 > /dev/null 2>&1
python scripts/launch_shares.py gpurun_out/r02f_launches_gehrd4096_first700.csv | head -12
