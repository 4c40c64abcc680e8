run() { env "$@" DR_CACHE_DIR=/tmp/c_$RANDOM timeout 120 python tools/heat_sweep.py 2>&1 | tail -1; }
(DR_ST_THREADS=992 DR_CACHE_DIR=/tmp/c_t timeout 300 python -m pytest tests -m gpu -q -x -k "heat or views or stencil" 2>&1 | tail -2)
run DR_ST_THREADS=992
run DR_ST_THREADS=992 DR_ST_NS=3
run DR_ST_THREADS=992 DR_ST_NS=5
run DR_ST_THREADS=992 DR_ST_TH=64 DR_ST_NS=3
run DR_ST_THREADS=992 DR_ST_TH=48 DR_ST_NS=4
run DR_ST_THREADS=992 DR_ST_TH=16 DR_ST_NS=6
run DR_ST_THREADS=744
