for b in qr3_dev qr3_dev_prof qr3_dev_real qr3_dev_real_prof; do echo "== $b"; ./scripts/dev/$b 64 8880 3 2>&1 | tail -4; done
echo "== lone warp (148 matrices)"; ./scripts/dev/qr3_dev_prof 64 148 3 | tail -1; ./scripts/dev/qr3_dev_real_prof 64 148 3 | tail -1
