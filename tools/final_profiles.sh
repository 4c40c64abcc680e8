#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): tests, smoke, bench lines, the ncu launch
# list of the bench command and one `--set full` capture per kernel family.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/final_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -2 > $O/final_smoke.log
python bench.py > $O/final_bench.json 2> $O/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2>> $O/final_bench.err
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --others > $O/final_bench_others.json 2>> $O/final_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $O/final_launch_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:dr_flat -s 2 -c 1 -o $O/bs_r13 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/ncu1.log 2>&1
$NCU -k regex:dr_stencil -s 3 -c 1 -o $O/stencil_r3 python tools/heat_sweep.py > $O/ncu2.log 2>&1
$NCU -k regex:dr_mm_skinny -s 1 -c 1 -o $O/skinny_r1 python tools/nbody_sweep.py > $O/ncu3.log 2>&1
$NCU -k regex:dr_flat -s 20 -c 1 -o $O/axpy_r1 python tools/host_overhead.py > $O/ncu4.log 2>&1
$NCU -k regex:dr_flat -s 4 -c 1 -o $O/l2_r1 python tools/scale_others.py > $O/ncu5.log 2>&1
tail -2 $O/final_pytest.log; cat $O/final_smoke.log; cut -c1-300 $O/final_bench.json
