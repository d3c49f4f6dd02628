python scripts/dev/e2e_streams.py 2>&1 | tail -9
python scripts/dev/e2e_streams.py real 2>&1 | tail -9
python -m pytest tests -x -q -m gpu -k "pageable or pipeline" 2>&1 | tail -3
