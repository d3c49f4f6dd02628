echo "== c64 n=64"; ./scripts/dev/qr3_dev 64 16384 2 | tail -6
echo "== f64 n=64"; ./scripts/dev/qr3_dev_real 64 16384 2 | tail -6
echo "== c64 n=32"; ./scripts/dev/qr3_dev 32 16384 2 | tail -6
echo "== f64 n=32"; ./scripts/dev/qr3_dev_real 32 16384 2| tail -6
echo "== odd sizes"; ./scripts/dev/qr3_dev 47 999 1; ./scripts/dev/qr3_dev_real 47 999 1; ./scripts/dev/qr3_dev 33 500 1; ./scripts/dev/qr3_dev_real 5 100 1; ./scripts/dev/qr3_dev 2 100 1;  ./scripts/dev/qr3_dev_real 64 100 1; ./scripts/dev/qr3_dev_real 63 100 1
