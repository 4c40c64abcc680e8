mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q --tb=short --timeout 120 -x -s) > gpurun_out/pytest.log 2>&1; grep -E "^exp|^log|passed|failed|Error" gpurun_out/pytest.log | cut -c1-200 | head -20
run() { env "$@" DR_CACHE_DIR=/tmp/c_$RANDOM timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$*', d['ms_per_step'], round(d['roofline']['frac'],4))"; }
run DR_X=0
run DR_PREFETCH=1
run DR_UNROLL=2
run DR_MINBLOCKS=5
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 5 --launch-count 1 -o gpurun_out/bs_r5 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bs.log 2>&1; tail -1 gpurun_out/ncu_bs.log
