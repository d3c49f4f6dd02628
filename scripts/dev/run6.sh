for c in 1 32; do for o in 1 2 4 6; do echo "== c64 CHAIN=$c CTAS=$o"; GSCHUR_QR_CTAS_PER_SM=$o GSCHUR_CHAIN=$c ./scripts/dev/qr3_dev_prof 64 2960 1 | egrep "rc=|profile"; done; done
for c in 1 32; do for o in 1 4 8 12; do echo "== f64 CHAIN=$c CTAS=$o"; GSCHUR_QR_CTAS_PER_SM=$o GSCHUR_CHAIN=$c ./scripts/dev/qr3_dev_real_prof 64 2960 1 | egrep "rc=|profile"; done; done
