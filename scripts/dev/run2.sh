for z in 0 1; do
echo "== ZREG=$z c64 n=64"; GSCHUR_ZREG=$z ./scripts/dev/qr3_dev 64 16384 2 | grep -v "^rc=0.*total [0-9]\{3,\}\." 
echo "== ZREG=$z f64 n=64"; GSCHUR_ZREG=$z ./scripts/dev/qr3_dev_real 64 16384 2
echo "== ZREG=$z c64 n=32"; GSCHUR_ZREG=$z ./scripts/dev/qr3_dev 32 16384 2
echo "== ZREG=$z f64 n=32"; GSCHUR_ZREG=$z ./scripts/dev/qr3_dev_real 32 16384 2
done
echo "== odd sizes"; ./scripts/dev/qr3_dev 47 999 1; ./scripts/dev/qr3_dev_real 47 999 1; ./scripts/dev/qr3_dev 33 500 1; ./scripts/dev/qr3_dev_real 5 100 1; ./scripts/dev/qr3_dev 2 100 1
