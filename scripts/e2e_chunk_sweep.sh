# end-to-end time of the host-pointer pipeline against the number of chunks (development aid)
for c in ${CHUNKS:-4 6 8 12}; do echo CHUNKS $c; GSCHUR_PIPE_CHUNKS=$c E2E_CALLS=6 python scripts/gpu_e2e_probe2.py 2>&1 | grep "e2e call [1-5]"; done
