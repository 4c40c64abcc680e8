run() { for k in 10 100; do env "$@" DR_CACHE_DIR=/tmp/c_$1 python bench.py --steps $k --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', d['steps'], round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; done; }
run DR_THREADS=1024 DR_VPL=2
run DR_THREADS=512 DR_VPL=4
run DR_THREADS=512 DR_VPL=2
run DR_THREADS=768 DR_VPL=2
