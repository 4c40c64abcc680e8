run() { env "$@" python bench.py --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['steps'], round(d['ms_per_step'],3), d['clocks'])"; }
run A=1 -- 2>/dev/null
for k in 10 20 50 100 200; do python bench.py --steps $k --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['steps'], round(d['ms_per_step'],3), d['clocks'])"; done
for t in 640 768; do DR_THREADS=$t DR_CACHE_DIR=/tmp/c_$t python bench.py --steps 100 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('threads $t', d['steps'], round(d['ms_per_step'],3), d['clocks'])"; done
