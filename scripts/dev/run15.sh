echo "== f64 overlap"; GSCHUR_OVERLAP=1 timeout 30 ./scripts/dev/qr3_dev_real 64 65536 3 2>&1 | grep -E "rc=|matrix|stats"
echo "== f64 overlap 2"; GSCHUR_OVERLAP=1 GSCHUR_OVERLAP_CTAS=2 timeout 30 ./scripts/dev/qr3_dev_real 64 65536 3 2>&1 | grep -E "rc="
echo "== c64 overlap, B 5/SM"; GSCHUR_OVERLAP=1 GSCHUR_QR_CTAS_PER_SM=5 timeout 30 ./scripts/dev/qr3_dev 64 65536 3 2>&1 | grep -E "rc=|matrix|stats"
