export SC_RES_TIMEOUT_S=3
for fm in 512 2048 4096 8192; do for rm in 32768 65536 131072; do
echo "== fine_max $fm res_max $rm"; SC_RES_FINE_MAX_PAIRS=$fm SC_RES_MAX_PAIRS=$rm timeout 100 python tools/res_prof.py 24 2>&1 | grep "proof [45]"; SC_RES_FINE_MAX_PAIRS=$fm SC_RES_MAX_PAIRS=$rm timeout 100 python tools/res_prof.py 20 2>&1 | grep "proof [45]"; done; done
