run() { env "$@" DR_CACHE_DIR=/tmp/c_$RANDOM timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$*', d['ms_per_step'], round(d['roofline']['frac'],4))"; }
run DR_STAGES=2
run DR_STAGES=4
run DR_STAGES=4 DR_MINBLOCKS=5
run DR_STAGES=3 DR_MINBLOCKS=3
run DR_STAGED=0 DR_UNROLL=2 DR_MINBLOCKS=3
