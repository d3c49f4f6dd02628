echo "== small"; for n in 2 3 5 7 23 33; do GSCHUR_CHAIN=2 timeout 60 ./scripts/dev/qr3_dev $n 300 1 | tail -3; done
for c in 1 2; do
echo "== CHAIN=$c c64 n=64"; GSCHUR_CHAIN=$c timeout 120 ./scripts/dev/qr3_dev 64 16384 2 | tail -7
echo "== CHAIN=$c c64 n=32"; GSCHUR_CHAIN=$c timeout 120 ./scripts/dev/qr3_dev 32 16384 2 | tail -3
done
GSCHUR_CHAIN=2 timeout 120 ./scripts/dev/qr3_dev_prof 64 2960 1 | tail -1
GSCHUR_QR_CTAS_PER_SM=1 GSCHUR_CHAIN=2 timeout 120 ./scripts/dev/qr3_dev_prof 64 592 1 | tail -1
