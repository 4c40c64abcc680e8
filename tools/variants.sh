(timeout 300 python -m pytest tests -m gpu -q --tb=short --timeout 120 -x -k "heat or views") 2>&1 | tail -2 | cut -c1-200
run() { env "$@" DR_CACHE_DIR=/tmp/c_$RANDOM timeout 120 python tools/heat_sweep.py 2>&1 | tail -1; }
run DR_ST_TH=64
run DR_ST_TH=32
run DR_ST_TH=32 DR_ST_NS=4
run DR_ST_TW=248 DR_ST_THREADS=248 DR_ST_TH=32
run DR_ST_TW=248 DR_ST_THREADS=248 DR_ST_TH=64
run DR_ST_TW=248 DR_ST_THREADS=496 DR_ST_TH=64
run DR_ST_TW=248 DR_ST_THREADS=496 DR_ST_TH=32 DR_ST_NS=4
