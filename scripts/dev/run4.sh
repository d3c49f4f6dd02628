for c in 32 1; do
echo "== CHAIN=$c c64 n=64"; GSCHUR_CHAIN=$c timeout 120 ./scripts/dev/qr3_dev 64 16384 2 | tail -7
echo "== CHAIN=$c c64 n=32"; GSCHUR_CHAIN=$c timeout 120 ./scripts/dev/qr3_dev 32 16384 2 | tail -7
done
echo "== odd"; for n in 2 3 4 5 7 23 33 47 63; do GSCHUR_CHAIN=1 timeout 60 ./scripts/dev/qr3_dev $n 500 1 | tail -6; done
