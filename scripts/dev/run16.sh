nproc; free -g | head -2
GSCHUR_PIPE_TRACE=1 python scripts/dev/pageable.py 2>&1 | tail -12
for t in 4 16; do echo "threads $t"; GSCHUR_COPY_THREADS=$t python scripts/dev/pageable.py 2>&1 | grep STAGING=1 | tail -2; done
python -m pytest tests -x -q -m gpu -k "full or api or host or pipeline or hess" 2>&1 | tail -5
