for g in v1 ""; do
echo "== GSCHUR_GEHRD=$g f64 n=64"; GSCHUR_GEHRD=$g timeout 120 ./scripts/dev/qr3_dev_real 64 16384 2 | tail -7
echo "== GSCHUR_GEHRD=$g f64 n=32"; GSCHUR_GEHRD=$g timeout 120 ./scripts/dev/qr3_dev_real 32 16384 2 | tail -3
done
for n in 2 3 5 17 33 47 63; do timeout 60 ./scripts/dev/qr3_dev_real $n 300 1 | tail -6; done
