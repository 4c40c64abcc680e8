run() { env "$@" DR_CACHE_DIR=/tmp/c_$RANDOM timeout 120 python tools/nbody_sweep.py 2>&1 | tail -1; }
run DR_SK_R=4 DR_SK_TH=128
run DR_SK_R=4 DR_SK_TH=256
run DR_SK_R=8 DR_SK_TH=128
run DR_SK_R=8 DR_SK_TH=64
run DR_SK_R=2 DR_SK_TH=256
run DR_SK_R=6 DR_SK_TH=128
run DR_SK_R=4 DR_SK_TH=64
